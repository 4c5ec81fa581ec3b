// Instance tables on the device: everything get_inst_info_dict
// (loader/postproc.py:12-98) and tiatoolbox's HoVerNet.get_instance_info (infer/wsi.py:150)
// compute per instance — bounding box (misc/utils.py:82-91), cv2.moments m00 / m10 / m01 of the
// box crop, the majority type with the "0 loses to the runner-up" rule (:60-68) and
// cv2.findContours(RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0] (:27-31, contour_core.h) — in four
// passes over an int32 label image instead of one Python iteration + three OpenCV calls per
// instance. The x2 nearest-neighbour cv2.resize of the label / type maps that precedes the call
// in tile mode (infer/tile.py:196-201) is the `up` argument: the kernels address the upsampled
// image directly, it is never materialised.
//
// HBM/latency-bound integer work: pass 1 is one read of the label (+ type) image with atomics
// aggregated over row runs; the border following is one warp per instance.
#include <climits>
#include <cstdio>

#include "capi_internal.cuh"
#include "contour_core.h"

using cerb::fail;

namespace {

// Type values are histogrammed in quarter units: class ids are integers (at most 7 classes in the
// reference's heads), but the WSI gland path hands over the type channel after cv2.resize(fx=0.5)
// (infer/wsi.py:773-790), i.e. 2x2 means of integers; np.unique then counts those values as they are.
constexpr int kTypes = 64;  // values 0, 0.25, ..., 15.75
constexpr int kThreads = 256;

struct Workspace {
  // per label (capacity cap_labels)
  unsigned long long *cnt = nullptr, *sx = nullptr, *sy = nullptr;
  int *rmin = nullptr, *rmax = nullptr, *cmin = nullptr, *cmax = nullptr;
  unsigned int* hist = nullptr;  // [label][kTypes]
  // per instance (dense, ascending id)
  int *ids = nullptr, *box = nullptr, *type = nullptr, *start = nullptr, *npts = nullptr;
  long long *mom = nullptr, *off = nullptr;
  size_t cap_labels = 0;
  int* mark = nullptr;  // [H*up, W*up]
  size_t cap_mark = 0;
  int* lab_stage = nullptr;
  float* type_stage = nullptr;
  size_t cap_lab_stage = 0, cap_type_stage = 0;
  int* xy = nullptr;
  size_t cap_xy = 0;
  int* scalars = nullptr;  // [0] max label, [1] flags (1 zero pixel, 2 negative label, 4 bad type), [2] n_inst
  long long* total = nullptr;
  // results of the last cerb_inst_info call
  int n_inst = 0;
  long long n_points = 0;
  std::vector<void*> owned;
  ~Workspace() {
    for (void* p : owned) cudaFree(p);
  }
};

Workspace* ws_for(cerb_ctx* ctx) {
  if (ctx->instinfo_ws == nullptr) {
    ctx->instinfo_ws = new Workspace();
    ctx->instinfo_ws_free = [](void* w) { delete static_cast<Workspace*>(w); };
  }
  return static_cast<Workspace*>(ctx->instinfo_ws);
}

template <typename T>
cudaError_t grow(Workspace* ws, T*& p, size_t& cap, size_t need, bool update_cap = true) {
  if (p != nullptr && need <= cap) return cudaSuccess;
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, (need ? need : 1) * sizeof(T));
  if (e != cudaSuccess) return e;
  if (p != nullptr) {
    for (size_t i = 0; i < ws->owned.size(); ++i)
      if (ws->owned[i] == p) {
        ws->owned.erase(ws->owned.begin() + i);
        break;
      }
    cudaFree(p);  // synchronises the device: nothing queued still reads the old buffer
  }
  ws->owned.push_back(q);
  p = static_cast<T*>(q);
  if (update_cap) cap = need;
  return cudaSuccess;
}

// ---- pass 0: largest label, background / negative flags -----------------------------------------
__global__ void k_label_range(const int* __restrict__ lab, size_t n, int* scalars) {
  int mx = 0, fl = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int v = lab[i];
    mx = max(mx, v);
    fl |= (v == 0 ? 1 : 0) | (v < 0 ? 2 : 0);
  }
  for (int o = 16; o; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    fl |= __shfl_xor_sync(0xffffffffu, fl, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (mx > 0) atomicMax(&scalars[0], mx);
    if (fl) atomicOr(&scalars[1], fl);
  }
}

__global__ void k_init_labels(unsigned long long* cnt, unsigned long long* sx,
                              unsigned long long* sy, int* rmin, int* rmax, int* cmin, int* cmax,
                              unsigned int* hist, int n_labels) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_labels) return;
  cnt[i] = sx[i] = sy[i] = 0ull;
  rmin[i] = cmin[i] = INT_MAX;
  rmax[i] = cmax[i] = -1;
  for (int t = 0; t < kTypes; ++t) hist[(size_t)i * kTypes + t] = 0u;
}

// ---- pass 1: per-label pixel count, coordinate sums, box, type histogram -------------------------
// A thread walks kSeg consecutive pixels of one row and issues its atomics once per run of equal
// labels (and once per run of equal types inside it).
constexpr int kSeg = 8;
__global__ void k_stats(const int* __restrict__ lab, const float* __restrict__ type, int H, int W,
                        unsigned long long* cnt, unsigned long long* sx, unsigned long long* sy,
                        int* rmin, int* rmax, int* cmin, int* cmax, unsigned int* hist,
                        int* scalars) {
  const int segs_per_row = (W + kSeg - 1) / kSeg;
  const size_t total = (size_t)H * segs_per_row;
  for (size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x; s < total;
       s += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(s / segs_per_row);
    const int x0 = (int)(s % segs_per_row) * kSeg;
    const int x1 = min(W, x0 + kSeg);
    int cur = 0, run_x0 = 0, tcur = -1, tcount = 0;
    unsigned long long sumx = 0;
    auto flush_type = [&]() {
      if (tcur >= 0 && tcount) atomicAdd(&hist[(size_t)cur * kTypes + tcur], (unsigned)tcount);
      tcount = 0;
    };
    auto flush = [&](int x_end) {  // run [run_x0, x_end) of label cur
      if (cur > 0) {
        atomicAdd(&cnt[cur], (unsigned long long)(x_end - run_x0));
        atomicAdd(&sx[cur], sumx);
        atomicAdd(&sy[cur], (unsigned long long)y * (x_end - run_x0));
        atomicMin(&rmin[cur], y);
        atomicMax(&rmax[cur], y);
        atomicMin(&cmin[cur], run_x0);
        atomicMax(&cmax[cur], x_end - 1);
        if (type) flush_type();
      }
    };
    for (int x = x0; x < x1; ++x) {
      const int v = lab[(size_t)y * W + x];
      if (v != cur) {
        flush(x);
        cur = v;
        run_x0 = x;
        sumx = 0;
        tcur = -1;
        tcount = 0;
      }
      if (v > 0) {
        sumx += (unsigned long long)x;
        if (type) {
          const float tf = type[(size_t)y * W + x];
          const float tq = tf * 4.f;
          int t = (int)tq;
          if (!((float)t == tq) || t < 0 || t >= kTypes) {
            atomicOr(&scalars[1], 4);
            t = 0;
          }
          if (t != tcur) {
            flush_type();
            tcur = t;
          }
          ++tcount;
        }
      }
    }
    flush(x1);
  }
}

// ---- single-block exclusive scan helper -----------------------------------------------------------
__device__ __forceinline__ long long block_exscan(long long v, long long* warp_sums, long long& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    long long w = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
    long long wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    warp_sums[lane] = wi - w;          // exclusive prefix of the warp sums
    if (lane == 31) warp_sums[32] = wi;  // block total
  }
  __syncthreads();
  const long long res = warp_sums[warp] + incl - v;
  total = warp_sums[32];
  __syncthreads();
  return res;
}

// ---- pass 2: dense instance table in ascending id order -------------------------------------------
// Type rule of loader/postproc.py:60-68: most frequent type, the smaller id on equal counts
// (np.unique order + stable sort); if that is 0 and another type occurs, the runner-up.
__global__ void __launch_bounds__(1024) k_compact(const unsigned long long* cnt, const unsigned long long* sx,
                          const unsigned long long* sy, const int* rmin, const int* rmax,
                          const int* cmin, const int* cmax, const unsigned int* hist, int n_labels,
                          int up, int has_type, int* ids, int* box, long long* mom, int* type,
                          int* scalars) {
  __shared__ long long warp_sums[33];
  long long base = 0;
  for (int l0 = 0; l0 < n_labels; l0 += blockDim.x) {
    const int l = l0 + threadIdx.x;
    const bool live = l > 0 && l < n_labels && cnt[l] > 0;
    long long tot;
    const long long pos = base + block_exscan(live ? 1 : 0, warp_sums, tot);
    if (live) {
      const long long u = up, c = (long long)cnt[l];
      const long long r0 = (long long)rmin[l] * u, c0 = (long long)cmin[l] * u;
      const long long cu = c * u * u;
      // sum over the up x up copies of every source pixel
      const long long sX = u * u * u * (long long)sx[l] + u * u * (u - 1) / 2 * c;
      const long long sY = u * u * u * (long long)sy[l] + u * u * (u - 1) / 2 * c;
      ids[pos] = l;
      box[pos * 4 + 0] = (int)r0;
      box[pos * 4 + 1] = (int)c0;
      box[pos * 4 + 2] = (rmax[l] + 1) * up;
      box[pos * 4 + 3] = (cmax[l] + 1) * up;
      mom[pos * 3 + 0] = cu;
      mom[pos * 3 + 1] = sX - c0 * cu;
      mom[pos * 3 + 2] = sY - r0 * cu;
      int best = -1;
      unsigned int best_n = 0;
      if (has_type) {
        const unsigned int* h = hist + (size_t)l * kTypes;
        for (int t = 0; t < kTypes; ++t)
          if (h[t] > best_n) {
            best = t;
            best_n = h[t];
          }
        if (best == 0) {
          int second = -1;
          unsigned int second_n = 0;
          for (int t = 1; t < kTypes; ++t)
            if (h[t] > second_n) {
              second = t;
              second_n = h[t];
            }
          if (second > 0) {
            best = second;
            best_n = second_n;
          }
        }
      }
      type[pos * 2 + 0] = best;
      type[pos * 2 + 1] = (int)(best_n * (unsigned)(up * up));
    }
    base += tot;
  }
  if (threadIdx.x == 0) scalars[2] = (int)base;
}

// ---- pass 3: one warp per instance runs the OpenCV scan over its box -------------------------------
__global__ void k_contour_scan(const int* __restrict__ lab, int* mark, int W, int up,
                               const int* __restrict__ ids, const int* __restrict__ box, int n_inst,
                               int* start, int* npts) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_inst) return;
  cc_view v;
  v.lab = lab;
  v.mark = mark;
  v.W = W;
  v.up = up;
  v.r0 = box[warp * 4 + 0];
  v.c0 = box[warp * 4 + 1];
  v.h = box[warp * 4 + 2] - v.r0;
  v.w = box[warp * 4 + 3] - v.c0;
  v.id = ids[warp];
  const cc_border b = cc_scan_warp(v, lane);
  if (lane == 0) {
    start[warp * 2 + 0] = b.y;
    start[warp * 2 + 1] = b.x;
    npts[warp] = b.npts;
  }
}

__global__ void __launch_bounds__(1024) k_offsets(const int* npts, int n, long long* off, long long* total_out) {
  __shared__ long long warp_sums[33];
  long long base = 0;
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    long long tot;
    const long long e = block_exscan(i < n ? npts[i] : 0, warp_sums, tot);
    if (i < n) off[i] = base + e;
    base += tot;
  }
  if (threadIdx.x == 0) {
    off[n] = base;
    *total_out = base;
  }
}

// ---- pass 4: re-follow the chosen border and keep the CHAIN_APPROX_SIMPLE points -------------------
__global__ void k_contour_write(const int* __restrict__ lab, int W, int up,
                                const int* __restrict__ ids, const int* __restrict__ box,
                                const int* __restrict__ start, const long long* __restrict__ off,
                                int n_inst, int* xy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inst || start[i * 2] < 0) return;
  cc_view v;
  v.lab = lab;
  v.mark = nullptr;
  v.W = W;
  v.up = up;
  v.r0 = box[i * 4 + 0];
  v.c0 = box[i * 4 + 1];
  v.h = box[i * 4 + 2] - v.r0;
  v.w = box[i * 4 + 3] - v.c0;
  v.id = ids[i];
  cc_trace<false>(v, start[i * 2], start[i * 2 + 1], 0, 0, xy + off[i] * 2, v.c0, v.r0);
}

}  // namespace

extern "C" int cerb_inst_info(cerb_ctx* ctx, const int32_t* labels, int H, int W,
                              const float* type_map, int up, int flags, int32_t* n_inst_out,
                              int64_t* n_points_out, int32_t* any_background_out) {
  if (!ctx || !labels || H <= 0 || W <= 0 || up < 1 || up > 8 ||
      (long long)H * up > INT_MAX / 2 || (long long)W * up > INT_MAX / 2)
    return fail(CERB_ERR_ARG, "cerb_inst_info: bad arguments");
  CERB_CUDA(cudaSetDevice(ctx->device));
  Workspace* ws = ws_for(ctx);
  cudaStream_t st = ctx->stream;
  const size_t px = (size_t)H * W;
  size_t dummy = 0;
  if (!ws->scalars) {
    CERB_CUDA(grow(ws, ws->scalars, dummy, 4, false));
    CERB_CUDA(grow(ws, ws->total, dummy, 1, false));
  }
  const int* lab = labels;
  const float* typ = type_map;
  if (!(flags & 1)) {  // host inputs
    CERB_CUDA(grow(ws, ws->lab_stage, ws->cap_lab_stage, px));
    CERB_CUDA(cudaMemcpyAsync(ws->lab_stage, labels, px * sizeof(int), cudaMemcpyHostToDevice, st));
    lab = ws->lab_stage;
    if (type_map) {
      CERB_CUDA(grow(ws, ws->type_stage, ws->cap_type_stage, px));
      CERB_CUDA(cudaMemcpyAsync(ws->type_stage, type_map, px * sizeof(float),
                                cudaMemcpyHostToDevice, st));
      typ = ws->type_stage;
    }
  }
  const size_t mark_px = px * up * up;
  CERB_CUDA(grow(ws, ws->mark, ws->cap_mark, mark_px));
  CERB_CUDA(cudaMemsetAsync(ws->mark, 0, mark_px * sizeof(int), st));
  CERB_CUDA(cudaMemsetAsync(ws->scalars, 0, 4 * sizeof(int), st));
  const int grid = 148 * 8;
  k_label_range<<<grid, kThreads, 0, st>>>(lab, px, ws->scalars);
  int sc[4] = {0, 0, 0, 0};
  CERB_CUDA(cudaMemcpyAsync(sc, ws->scalars, sizeof(sc), cudaMemcpyDeviceToHost, st));
  CERB_CUDA(cudaStreamSynchronize(st));
  ctx->launches += 1;
  if (sc[1] & 2) return fail(CERB_ERR_ARG, "cerb_inst_info: negative instance id");
  if ((size_t)sc[0] > px * 4 + 1024)
    return fail(CERB_ERR_ARG, "cerb_inst_info: instance id %d is not a label of this image", sc[0]);
  const int n_labels = sc[0] + 1;
  if ((size_t)n_labels > ws->cap_labels || !ws->cnt) {
    const size_t need = (size_t)n_labels + n_labels / 2 + 1024;
    size_t c;
#define GROWL(field, mul)  \
  c = ws->cap_labels;      \
  CERB_CUDA(grow(ws, ws->field, c, need * (mul), false));
    GROWL(cnt, 1) GROWL(sx, 1) GROWL(sy, 1) GROWL(rmin, 1) GROWL(rmax, 1) GROWL(cmin, 1)
    GROWL(cmax, 1) GROWL(hist, kTypes) GROWL(ids, 1) GROWL(box, 4) GROWL(type, 2) GROWL(start, 2)
    GROWL(npts, 1) GROWL(mom, 3) GROWL(off, 2)
#undef GROWL
    ws->cap_labels = need;
  }
  k_init_labels<<<(n_labels + kThreads - 1) / kThreads, kThreads, 0, st>>>(
      ws->cnt, ws->sx, ws->sy, ws->rmin, ws->rmax, ws->cmin, ws->cmax, ws->hist, n_labels);
  k_stats<<<grid, kThreads, 0, st>>>(lab, typ, H, W, ws->cnt, ws->sx, ws->sy, ws->rmin, ws->rmax,
                                     ws->cmin, ws->cmax, ws->hist, ws->scalars);
  k_compact<<<1, 1024, 0, st>>>(ws->cnt, ws->sx, ws->sy, ws->rmin, ws->rmax, ws->cmin, ws->cmax,
                                ws->hist, n_labels, up, typ ? 1 : 0, ws->ids, ws->box, ws->mom,
                                ws->type, ws->scalars);
  CERB_CUDA(cudaMemcpyAsync(sc, ws->scalars, sizeof(sc), cudaMemcpyDeviceToHost, st));
  CERB_CUDA(cudaStreamSynchronize(st));
  ctx->launches += 3;
  if (sc[1] & 4)
    return fail(CERB_ERR_ARG, "cerb_inst_info: type map holds a value that is not a multiple of 0.25 in [0, %d)",
                kTypes / 4);
  const int n = sc[2];
  long long total = 0;
  if (n > 0) {
    k_contour_scan<<<(n * 32 + kThreads - 1) / kThreads, kThreads, 0, st>>>(
        lab, ws->mark, W, up, ws->ids, ws->box, n, ws->start, ws->npts);
    k_offsets<<<1, 1024, 0, st>>>(ws->npts, n, ws->off, ws->total);
    CERB_CUDA(cudaMemcpyAsync(&total, ws->total, sizeof(total), cudaMemcpyDeviceToHost, st));
    CERB_CUDA(cudaStreamSynchronize(st));
    CERB_CUDA(grow(ws, ws->xy, ws->cap_xy, (size_t)total * 2 + 2));
    k_contour_write<<<(n + 127) / 128, 128, 0, st>>>(lab, W, up, ws->ids, ws->box, ws->start,
                                                     ws->off, n, ws->xy);
    ctx->launches += 3;
  }
  CERB_CUDA(cudaGetLastError());
  ws->n_inst = n;
  ws->n_points = total;
  if (n_inst_out) *n_inst_out = n;
  if (n_points_out) *n_points_out = total;
  if (any_background_out) *any_background_out = sc[1] & 1;
  return CERB_OK;
}

extern "C" int cerb_inst_info_read(cerb_ctx* ctx, int32_t* ids, int32_t* box, int64_t* moments,
                                   int32_t* type, int64_t* contour_off, int32_t* contour_xy) {
  if (!ctx || !ctx->instinfo_ws) return fail(CERB_ERR_ARG, "cerb_inst_info_read: no result");
  CERB_CUDA(cudaSetDevice(ctx->device));
  Workspace* ws = static_cast<Workspace*>(ctx->instinfo_ws);
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)ws->n_inst;
  if (n) {
    if (ids) CERB_CUDA(cudaMemcpyAsync(ids, ws->ids, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (box) CERB_CUDA(cudaMemcpyAsync(box, ws->box, n * 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (moments)
      CERB_CUDA(cudaMemcpyAsync(moments, ws->mom, n * 3 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    if (type) CERB_CUDA(cudaMemcpyAsync(type, ws->type, n * 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (contour_off)
      CERB_CUDA(cudaMemcpyAsync(contour_off, ws->off, (n + 1) * sizeof(long long),
                                cudaMemcpyDeviceToHost, st));
    if (contour_xy && ws->n_points)
      CERB_CUDA(cudaMemcpyAsync(contour_xy, ws->xy, (size_t)ws->n_points * 2 * sizeof(int),
                                cudaMemcpyDeviceToHost, st));
  } else if (contour_off) {
    contour_off[0] = 0;
  }
  CERB_CUDA(cudaStreamSynchronize(st));
  return CERB_OK;
}
