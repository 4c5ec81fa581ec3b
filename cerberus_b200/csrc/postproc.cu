// On-device instance post-processing (integer / bit work, HBM- and latency-bound).
//
// Replaces loader/postproc.py:268-407 (PostProcInstErodedContourMap) of the reference, which
// runs scipy.ndimage / scikit-image / OpenCV on the CPU:
//   __proc_nuclei :352-381  threshold, erode(cross), CC + size filters, fill holes, CC,
//                           marker-controlled watershed (skimage 0.19 heap, restated in
//                           SURVEY.md Appendix D)
//   __proc_gland  :270-309  / __proc_lumen :312-350
//                           threshold, size filter, CC, per-instance crop-limited dilation
//                           with an even-sized ellipse + crop-limited hole filling, painted in
//                           ascending id order (= per-pixel max id)
// All label maps are int32 and bit-exact with the reference given identical float inputs.
// Everything is batched over images (gridDim.y / blockIdx of the per-image kernels); there is
// no host synchronisation inside a call apart from the final copy-out.
#include <cuda_runtime.h>

#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "capi_internal.cuh"

using namespace cerb;

namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------ union-find CC
// 4-connectivity; the representative of a component is its smallest linear index, which is
// also its first pixel in raster order (scipy.ndimage.label numbering is by that pixel).
__device__ __forceinline__ int uf_find(const int* L, int x) {
  int p = L[x];
  while (p != x) {
    x = p;
    p = L[x];
  }
  return x;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  bool done;
  do {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a < b) {
      const int old = atomicMin(&L[b], a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const int old = atomicMin(&L[a], b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// fg: [n][hw] bytes (0/1). L: [n][hw].
// Initial label = first pixel of the horizontal run inside the warp's 32-pixel segment (found
// with one ballot), so a run is a depth-1 tree before any atomic is issued.
__global__ void k_cc_init(const uint8_t* __restrict__ fg, int* __restrict__ L, int* __restrict__ size,
                          int hw, int W) {
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  const int lane = threadIdx.x & 31;
  const int hw_up = (hw + 31) & ~31;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw_up; p += gridDim.x * blockDim.x) {
    const bool inb = p < hw;
    const bool f = inb && fg[base + p];
    const bool prev_f = __shfl_up_sync(0xffffffffu, f ? 1 : 0, 1) != 0;
    const bool cont = f && lane > 0 && prev_f && (p % W) != 0;  // continues the previous lane's run
    const unsigned breaks = ~__ballot_sync(0xffffffffu, cont);
    if (inb) {
      const int start = 31 - __clz(breaks & (0xffffffffu >> (31 - lane)));
      L[base + p] = f ? p - (lane - start) : -1;
      size[base + p] = 0;
    }
  }
}

// Unions that the run initialisation has not already implied: the horizontal link across a
// 32-pixel segment boundary, and the vertical link unless the left and upper-left neighbours
// are both foreground (then the link was made one pixel to the left).
__global__ void k_cc_merge(const uint8_t* __restrict__ fg, int* __restrict__ L, int H, int W) {
  const int hw = H * W;
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  const uint8_t* f = fg + base;
  int* l = L + base;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    if (!f[p]) continue;
    const int x = p % W;
    const bool left = x > 0 && f[p - 1];
    if (left && (p & 31) == 0) uf_union(l, p, p - 1);
    if (p >= W && f[p - W]) {
      if (!(left && f[p - W - 1])) uf_union(l, p, p - W);
    }
  }
}

__global__ void k_cc_flatten_count(const uint8_t* __restrict__ fg, int* __restrict__ L,
                                   int* __restrict__ size, int hw) {
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  const int lane = threadIdx.x & 31;
  const int hw_up = (hw + 31) & ~31;
  // Neighbouring pixels mostly share their root: one atomicAdd per (warp, root) instead of one
  // per pixel (a 40 K-pixel background component otherwise serialises 40 K atomics on one word).
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw_up; p += gridDim.x * blockDim.x) {
    const bool f = p < hw && fg[base + p];
    int r = -1;
    if (f) {
      r = uf_find(L + base, p);
      L[base + p] = r;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, r);
    if (f && lane == __ffs(peers) - 1) atomicAdd(&size[base + r], __popc(peers));
  }
}

// skimage.morphology.remove_small_objects: drop components with size < min_size (strict).
__global__ void k_filter_small(uint8_t* __restrict__ fg, const int* __restrict__ L,
                               const int* __restrict__ size, int hw, int min_size) {
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    if (fg[base + p] && size[base + L[base + p]] < min_size) fg[base + p] = 0;
  }
}

// One block per image: rank[p] = 1 + number of surviving roots before p (raster order) for
// root pixels; count[img] = number of components.
__global__ void k_rank_roots(const uint8_t* __restrict__ fg, const int* __restrict__ L,
                             int* __restrict__ rank, int* __restrict__ count,
                             int* __restrict__ has_bg, int hw) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const size_t base = static_cast<size_t>(blockIdx.x) * hw;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  int any_bg = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // four consecutive pixels per thread and iteration: a quarter of the block-wide scans
  for (int start = 0; start < hw; start += 4 * blockDim.x) {
    const int p0 = start + 4 * threadIdx.x;
    int flags[4];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = p0 + i;
      const bool in = p < hw;
      const bool f = in && fg[base + p];
      flags[i] = (f && L[base + p] == p) ? 1 : 0;
      if (in && !f) any_bg = 1;
      cnt += flags[i];
    }
    int v = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int w = lane < nwarps ? warp_sums[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_sums[lane] = w;  // inclusive
    }
    __syncthreads();
    int before = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + v - cnt;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (flags[i]) rank[base + p0 + i] = ++before;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_sums[nwarps - 1];
    __syncthreads();
  }
  any_bg = __syncthreads_or(any_bg);
  if (threadIdx.x == 0) {
    count[blockIdx.x] = carry;
    if (has_bg != nullptr) has_bg[blockIdx.x] = any_bg;
  }
}

__global__ void k_apply_rank(const uint8_t* __restrict__ fg, const int* __restrict__ L,
                             const int* __restrict__ rank, int* __restrict__ lab, int hw) {
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    lab[base + p] = fg[base + p] ? rank[base + L[base + p]] : 0;
  }
}

// ------------------------------------------------------------------ nuclei
// loader/postproc.py:358-363,370-371
__global__ void k_nuc_threshold(const float* __restrict__ canvas, int C, int ch0,
                                uint8_t* __restrict__ msk, uint8_t* __restrict__ mrk,
                                float* __restrict__ val, int* __restrict__ any_fg, int hw) {
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  int any = 0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    const float inner = canvas[(base + p) * C + ch0];
    const float cnt = canvas[(base + p) * C + ch0 + 1];
    const float raw = __fadd_rn(inner, cnt);
    const uint8_t m = raw > 0.5f;
    msk[base + p] = m;
    mrk[base + p] = inner > 0.5f;
    val[base + p] = -inner;
    any |= m;
  }
  if (__syncthreads_or(any) && threadIdx.x == 0) atomicOr(&any_fg[blockIdx.y], 1);
}

// cv2.erode with the 3x3 MORPH_ELLIPSE element (= cross); pixels outside the image do not
// constrain the result (OpenCV's default morphology border value).
__global__ void k_erode_cross(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W) {
  const int hw = H * W;
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  const uint8_t* f = in + base;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    const int x = p % W, y = p / W;
    uint8_t v = f[p];
    if (x > 0) v &= f[p - 1];
    if (x < W - 1) v &= f[p + 1];
    if (y > 0) v &= f[p - W];
    if (y < H - 1) v &= f[p + W];
    out[base + p] = v;
  }
}

__global__ void k_invert(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int hw) {
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x)
    out[base + p] = !in[base + p];
}

// scipy.ndimage.binary_fill_holes: background components (4-conn) that do not touch the
// image border are holes. `size` is reused as the per-root "touches the border" flag.
__global__ void k_mark_border(const uint8_t* __restrict__ bg, const int* __restrict__ L,
                              int* __restrict__ outer, int H, int W) {
  const int hw = H * W;
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  const int nb = 2 * (H + W);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += gridDim.x * blockDim.x) {
    int p;
    if (i < W) p = i;
    else if (i < 2 * W) p = (H - 1) * W + (i - W);
    else if (i < 2 * W + H) p = (i - 2 * W) * W;
    else p = (i - 2 * W - H) * W + (W - 1);
    if (bg[base + p]) outer[base + L[base + p]] = -1;
  }
}

__global__ void k_fill_holes(uint8_t* __restrict__ fg, const uint8_t* __restrict__ bg,
                             const int* __restrict__ L, const int* __restrict__ outer, int hw) {
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    if (bg[base + p] && outer[base + L[base + p]] != -1) fg[base + p] = 1;
  }
}

// markers * mask (skimage watershed _validate_inputs); also the initial output map.
__global__ void k_mask_markers(int* __restrict__ lab, const uint8_t* __restrict__ msk, int hw) {
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x)
    if (!msk[base + p]) lab[base + p] = 0;
}

// ---- marker ties that provably cannot matter -----------------------------------------------------
// Two marker entries a, e of one component with bit-equal values leave the reference's queue in an
// order fixed by its global heap layout. ws_tie_harmless proves, from the flood itself, that the
// order is irrelevant. Setting: a and e surfaced back to back - no other entry left the queue in
// between ("clean"), so every pixel a pushed has a value >= the tie value. a_mask / e_mask: which
// of the neighbours -W, -1, +1, +W (bits 0..3) each of them labelled and pushed when it surfaced.
// Harmless iff (1) no pixel pushed by a is 4-adjacent to e (a pixel next to both would get the
// label of whichever surfaces first), (2) no pixel pushed by e has a value below the tie value (in
// the other order it would surface - and flood - before a), (3) no pixel pushed by a shares its
// value with a pixel pushed by e (their ages swap between the two orders; ages only break ties
// between equal values). Then both orders push the same pixels with the same labels, the age
// counter ends at the same value, and only the ages of those <= 8 entries are exchanged between
// entries of different value: every later event is identical. The argument is pairwise, so a
// group of k tied entries that are pairwise harmless is harmless in all k! orders.
__device__ __forceinline__ uint32_t ws_valkey(float v) {
  uint32_t u = __float_as_uint(v);
  if ((u << 1) == 0) u = 0;  // -0.0 == +0.0
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ int ws_nb(int p, int i, int W) {
  return i == 0 ? p - W : i == 1 ? p - 1 : i == 2 ? p + 1 : p + W;
}
__device__ __noinline__ bool ws_tie_harmless(const float* __restrict__ v, int W, int a,
                                             unsigned a_mask, int e, unsigned e_mask) {
  const int ey = e / W, ex = e % W;
  const uint32_t tie = ws_valkey(v[e]);
  for (int i = 0; i < 4; ++i) {
    if (!(a_mask & (1u << i))) continue;
    const int q = ws_nb(a, i, W);
    const int dy = q / W - ey, dx = q % W - ex;
    if (dy * dy + dx * dx == 1) return false;  // (1)
  }
  for (int j = 0; j < 4; ++j) {
    if (!(e_mask & (1u << j))) continue;
    const uint32_t kr = ws_valkey(v[ws_nb(e, j, W)]);
    if (kr < tie) return false;  // (2)
    for (int i = 0; i < 4; ++i)
      if ((a_mask & (1u << i)) && ws_valkey(v[ws_nb(a, i, W)]) == kr) return false;  // (3)
  }
  return true;
}

// ---- exact emulation of skimage 0.19 watershed_raveled (connectivity 1, no compactness,
// no watershed line). Ordering is (value, age) with a global push counter, ties between
// equal keys are decided by the array-heap layout, so the heap itself is reproduced:
// push = append + sift-up, pop = move last to root + sift-down preferring the left child.
// One image per block; the queue is sequential by definition, lane 0 drives it. The top
// kTopLevels of the heap live in shared memory, deeper levels in global memory (L1/L2).
constexpr int kHeapTop = 2047;  // 11 levels

struct HeapRef {
  float* gv;
  int* ga;
  int* gi;
  float* sv;
  int* sa;
  int* si;
};

__device__ __forceinline__ void heap_get(const HeapRef& h, int i, float& v, int& a, int& idx) {
  if (i < kHeapTop) {
    v = h.sv[i]; a = h.sa[i]; idx = h.si[i];
  } else {
    v = h.gv[i]; a = h.ga[i]; idx = h.gi[i];
  }
}
__device__ __forceinline__ void heap_set(const HeapRef& h, int i, float v, int a, int idx) {
  if (i < kHeapTop) {
    h.sv[i] = v; h.sa[i] = a; h.si[i] = idx;
  } else {
    h.gv[i] = v; h.ga[i] = a; h.gi[i] = idx;
  }
}
__device__ __forceinline__ bool ws_smaller(float va, int aa, float vb, int ab) {
  return (va != vb) ? (va < vb) : (aa < ab);
}

__device__ __forceinline__ void heap_push(const HeapRef& h, int& n, float v, int a, int idx) {
  int c = n++;
  while (c > 0) {
    const int parent = (c + 1) / 2 - 1;
    float pv; int pa, pi;
    heap_get(h, parent, pv, pa, pi);
    if (!ws_smaller(v, a, pv, pa)) break;
    heap_set(h, c, pv, pa, pi);
    c = parent;
  }
  heap_set(h, c, v, a, idx);
}

__device__ __forceinline__ void heap_pop(const HeapRef& h, int& n, float& tv, int& ta, int& ti) {
  heap_get(h, 0, tv, ta, ti);
  --n;
  if (n == 0) return;
  float xv; int xa, xi;
  heap_get(h, n, xv, xa, xi);
  int i = 0;
  for (;;) {
    const int l = 2 * i + 1;
    if (l >= n) break;
    const int r = l + 1;
    float lv; int la, li;
    heap_get(h, l, lv, la, li);
    int s = i;
    float sv = xv; int sa = xa, si = xi;
    if (ws_smaller(lv, la, xv, xa)) { s = l; sv = lv; sa = la; si = li; }
    if (r < n) {
      float rv; int ra, ri;
      heap_get(h, r, rv, ra, ri);
      if (ws_smaller(rv, ra, sv, sa)) { s = r; sv = rv; sa = ra; si = ri; }
    }
    if (s == i) break;
    heap_set(h, i, sv, sa, si);
    i = s;
  }
  heap_set(h, i, xv, xa, xi);
}

__global__ void __launch_bounds__(32, 1)
k_watershed(const float* __restrict__ val, const uint8_t* __restrict__ msk, int* __restrict__ out,
            float* __restrict__ heap_v, int* __restrict__ heap_a, int* __restrict__ heap_i, int H,
            int W, const int* __restrict__ run_flag) {
  __shared__ float sv[kHeapTop];
  __shared__ int sa[kHeapTop];
  __shared__ int si[kHeapTop];
  if (threadIdx.x != 0) return;
  if (run_flag != nullptr && run_flag[blockIdx.x] == 0) return;  // the component-parallel path succeeded
  const int hw = H * W;
  const size_t base = static_cast<size_t>(blockIdx.x) * hw;
  const float* v = val + base;
  const uint8_t* m = msk + base;
  int* o = out + base;
  HeapRef h{heap_v + base, heap_a + base, heap_i + base, sv, sa, si};
  int n = 0;
  for (int p = 0; p < hw; ++p)
    if (o[p] != 0) heap_push(h, n, v[p], 0, p);
  int age = 1;
  while (n > 0) {
    float ev; int ea, ei;
    heap_pop(h, n, ev, ea, ei);
    const int lab = o[ei];
    const int x = ei % W;
    // neighbour order of _offsets_to_raveled_neighbors (connectivity 1): -W, -1, +1, +W
    int q = ei - W;
    if (q >= 0 && m[q] && o[q] == 0) { ++age; o[q] = lab; heap_push(h, n, v[q], age, q); }
    q = ei - 1;
    if (x > 0 && m[q] && o[q] == 0) { ++age; o[q] = lab; heap_push(h, n, v[q], age, q); }
    q = ei + 1;
    if (x < W - 1 && m[q] && o[q] == 0) { ++age; o[q] = lab; heap_push(h, n, v[q], age, q); }
    q = ei + W;
    if (q < hw && m[q] && o[q] == 0) { ++age; o[q] = lab; heap_push(h, n, v[q], age, q); }
  }
}

// ---- fast path: images of <= 65536 pixels keep the label map (u16) and the first kWsHeapSmem
// heap nodes in shared memory. A heap node is ONE 64-bit word [value:32 | age:16 | pixel:16]:
// the value is mapped to an order-preserving unsigned, ages fit 16 bits (at most 65535
// non-marker pushes), and the comparison looks at the upper 48 bits only, so equal
// (value, age) pairs still compare "not smaller" exactly like the reference. Same algorithm,
// same heap layout; the sift-down reads children and grandchildren together (two levels per
// shared-memory round trip) and the neighbours' values are fetched before the sift-down.
constexpr int kWsThreads = 128;
constexpr int kWsHeapSmem = 10500;
constexpr uint16_t kWsOutside = 0xFFFF;  // not in the mask

__device__ __forceinline__ uint64_t ws_entry(float v, uint32_t age, int pix) {
  uint32_t u = __float_as_uint(v);
  if ((u << 1) == 0) u = 0;  // -0.0 == +0.0 for the reference's float compare
  u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;
  const uint32_t a16 = age == 0 ? 0u : age - 1u;  // ages start at 2 (watershed_raveled)
  return (static_cast<uint64_t>(u) << 32) | (static_cast<uint64_t>(a16) << 16) |
         static_cast<uint32_t>(pix);
}

// Warp-cooperative version of the same heap (same array layout after every operation). One
// GPU thread retires roughly one dependent instruction every 4-6 cycles, so the sequential
// sift loops above cost ~2000 cycles per queue operation. Here warp 0 owns the heap:
//  * push: the ancestor chain of the new leaf is known in advance ((c+1) >> d), so lane d loads
//    ancestor d+1, every lane compares with the new key at once, a ballot gives the number of
//    moves, the moves are independent stores;
//  * pop: the hole follows the "smaller child" path, which does not depend on the element being
//    re-inserted; one byte per node caches "the right child is strictly smaller", so the path
//    is a chase over bytes, after which lanes load the path entries in parallel, a ballot finds
//    where the re-inserted element stops, and the moves are independent stores.
// The cached bytes of the touched nodes are recomputed in parallel from the moved entries and
// their (untouched) siblings. All arguments are warp-uniform; every lane must call.
struct WarpHeap {
  uint64_t* s;
  uint64_t* g;
  uint8_t* sb;
  uint8_t* gb;
  int lane;
  __device__ __forceinline__ uint64_t get(int i) const { return i < kWsHeapSmem ? s[i] : g[i]; }
  __device__ __forceinline__ void set(int i, uint64_t e) const {
    if (i < kWsHeapSmem) s[i] = e; else g[i] = e;
  }
  __device__ __forceinline__ int getb(int i) const { return i < kWsHeapSmem ? sb[i] : gb[i]; }
  __device__ __forceinline__ void setb(int i, int b) const {
    if (i < kWsHeapSmem) sb[i] = static_cast<uint8_t>(b); else gb[i] = static_cast<uint8_t>(b);
  }

  __device__ __forceinline__ void push(int& n, uint64_t e) const {
    const int c = n;
    n = c + 1;
    const int dep = 31 - __clz(c + 1);  // depth of the new leaf; ancestors d = 1..dep
    const uint64_t k = e >> 16;
    const int d = lane;
    const int a_d = ((c + 1) >> d) - 1;        // ancestor d levels up (d = 0: the leaf slot)
    const int a_p = ((c + 1) >> (d + 1)) - 1;  // its parent
    const bool has_parent = d < dep;
    uint64_t P = 0;
    if (has_parent) P = get(a_p);
    const unsigned mv = __ballot_sync(0xffffffffu, has_parent && (k < (P >> 16)));
    const int t = __ffs(~mv) - 1;  // number of moves (the predicate is monotone along the chain)
    // sibling of a_d (child of a_p on the chain); needed by lanes d <= t with a parent
    uint64_t S = 0;
    bool sib_exists = false;
    const bool upd = has_parent && d <= t;
    const bool is_left = (a_d & 1) != 0;
    if (upd) {
      const int sib = is_left ? a_d + 1 : a_d - 1;
      sib_exists = sib < n;
      if (sib_exists) S = get(sib);
    }
    __syncwarp();
    if (d < t) set(a_d, P);
    if (d == t) set(a_d, e);
    if (upd) {
      const uint64_t nc = (d < t ? P : e) >> 16;  // new content of slot a_d
      const int bit = is_left ? (sib_exists && ((S >> 16) < nc)) : (nc < (S >> 16));
      setb(a_p, bit);
    }
    __syncwarp();
  }

  // Removes the root (the caller has already read it).
  __device__ __forceinline__ void remove_top(int& n) const {
    n = n - 1;
    if (n == 0) return;
    const uint64_t x = get(n);
    const uint64_t xk = x >> 16;
    if (lane == 0 && (n & 1) == 0) setb((n - 1) >> 1, 0);  // the last node was a right child
    __syncwarp();
    int c = 0, my_c = 0, my_child = 0, D = 0;
    for (;;) {
      const int l = 2 * c + 1;
      if (l >= n) break;
      const int ch = l + getb(c);
      if (lane == D) { my_c = c; my_child = ch; }
      c = ch;
      ++D;
    }
    const int d = lane;
    uint64_t E = 0;
    if (d < D) E = get(my_child);
    const unsigned mv = __ballot_sync(0xffffffffu, d < D && ((E >> 16) < xk));
    const int t = __ffs(~mv) - 1;
    const uint64_t E_next = __shfl_down_sync(0xffffffffu, E, 1);
    uint64_t S = 0;
    bool sib_exists = false;
    const bool is_left = (my_child & 1) != 0;
    if (d < t) {
      const int sib = is_left ? my_child + 1 : my_child - 1;
      sib_exists = sib < n;
      if (sib_exists) S = get(sib);
    }
    __syncwarp();
    if (d < t) set(my_c, E);
    if (t < D) {
      if (d == t) set(my_c, x);
    } else if (d == 0) {
      set(c, x);  // the hole reached a leaf
    }
    if (d < t) {
      const uint64_t nc = (d + 1 < t ? E_next : x) >> 16;  // new content of slot my_child
      const int bit = is_left ? (sib_exists && ((S >> 16) < nc)) : (nc < (S >> 16));
      setb(my_c, bit);
    }
    __syncwarp();
  }
};

__global__ void __launch_bounds__(kWsThreads, 1)
k_watershed_smem(const float* __restrict__ val, const uint8_t* __restrict__ msk,
                 int* __restrict__ out, uint64_t* __restrict__ heap_k, uint8_t* __restrict__ heap_b,
                 int* __restrict__ list_idx, float* __restrict__ list_val,
                 const int* __restrict__ run_flag, int H, int W) {
  extern __shared__ __align__(16) uint8_t ws_smem[];
  if (run_flag != nullptr && run_flag[blockIdx.x] == 0) return;  // the fast path handled this image
  __shared__ int s_warp[kWsThreads / 32];
  __shared__ int s_carry;
  const int hw = H * W;
  uint64_t* sh = reinterpret_cast<uint64_t*>(ws_smem);
  uint16_t* lab16 = reinterpret_cast<uint16_t*>(ws_smem + sizeof(uint64_t) * kWsHeapSmem);
  uint8_t* sbits = ws_smem + sizeof(uint64_t) * kWsHeapSmem + 2u * ((hw + 7) & ~7);
  const size_t base = static_cast<size_t>(blockIdx.x) * hw;
  const float* v = val + base;
  const uint8_t* m = msk + base;
  int* o = out + base;
  int* lidx = list_idx + base;
  float* lval = list_val + base;
  // parallel: stage labels, compact the marker pixels in raveled order
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int start = 0; start < hw; start += kWsThreads) {
    const int p = start + threadIdx.x;
    int flag = 0;
    if (p < hw) {
      const int l = o[p];  // markers * mask (k_mask_markers)
      lab16[p] = m[p] ? static_cast<uint16_t>(l) : kWsOutside;
      flag = l != 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = s_carry;
    for (int w2 = 0; w2 < warp; ++w2) before += s_warp[w2];
    if (flag) {
      const int pos = before + __popc(bal & ((1u << lane) - 1u));
      lidx[pos] = p;
      lval[pos] = v[p];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w2 = 0; w2 < kWsThreads / 32; ++w2) t += s_warp[w2];
      s_carry += t;
    }
    __syncthreads();
  }
  const int n_markers = s_carry;
  __threadfence_block();
  __syncthreads();
  if (warp == 0) {
    WarpHeap h{sh, heap_k + base, sbits, heap_b + base, lane};
    int n = 0;
    for (int j0 = 0; j0 < n_markers; j0 += 32) {
      const int j = j0 + lane;
      uint64_t mine = 0;
      if (j < n_markers) mine = ws_entry(lval[j], 0u, lidx[j]);
      const int cnt = min(32, n_markers - j0);
      for (int k = 0; k < cnt; ++k) h.push(n, __shfl_sync(0xffffffffu, mine, k));
    }
    uint32_t age = 1;
    while (n > 0) {
      const int ei = static_cast<int>(h.get(0) & 0xFFFFu);
      const int x = ei % W;
      const uint16_t lab = lab16[ei];
      // neighbour order of _offsets_to_raveled_neighbors (connectivity 1): -W, -1, +1, +W;
      // lane k < 4 examines neighbour k and fetches its value while the root is removed
      int q = ei;
      bool ok = false;
      if (lane == 0) { q = ei - W; ok = q >= 0; }
      if (lane == 1) { q = ei - 1; ok = x > 0; }
      if (lane == 2) { q = ei + 1; ok = x < W - 1; }
      if (lane == 3) { q = ei + W; ok = q < hw; }
      ok = ok && lab16[q] == 0;
      float vq = 0.f;
      if (ok) vq = __ldg(v + q);
      const unsigned todo = __ballot_sync(0xffffffffu, ok);
      h.remove_top(n);
      for (int k = 0; k < 4; ++k) {
        if (!((todo >> k) & 1u)) continue;
        ++age;
        const int qk = __shfl_sync(0xffffffffu, q, k);
        const float vk = __shfl_sync(0xffffffffu, vq, k);
        if (lane == 0) lab16[qk] = lab;
        h.push(n, ws_entry(vk, age, qk));  // push ends with __syncwarp
      }
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < hw; p += kWsThreads) {
    const uint16_t l = lab16[p];
    o[p] = l == kWsOutside ? 0 : static_cast<int>(l);
  }
}

// ---- component-parallel fast path. Inside one 4-connected component of the mask the order of
// queue events depends only on that component's own entries: non-marker entries carry unique,
// monotonically assigned ages, so the pop order restricted to a component is a total order on
// its own keys - unless two of its marker pixels have bit-equal values (age 0 ties are resolved
// by the global heap layout). Marker pixels without an unlabeled in-mask neighbour push nothing
// and are inert. So: one thread per mask component floods it with a private heap holding only
// its boundary markers; markers leave the queue in non-decreasing value order, so a tie shows
// up as two consecutive marker pops with equal values. Most ties provably cannot matter
// (ws_tie_harmless) and the flood goes on; otherwise the tile is handed, untouched, to the exact
// whole-tile emulation above (k_watershed_smem). Results are identical either way.
constexpr int kWcThreads = 256;
constexpr int kWcPool = 9000;     // heap entries (8 bytes) in shared memory
constexpr int kWcMaxComp = 1024;

__global__ void __launch_bounds__(kWcThreads, 1)
k_watershed_comp(const float* __restrict__ val, const uint8_t* __restrict__ msk,
                 const int* __restrict__ Lroot, const int* __restrict__ csize,
                 int* __restrict__ out, int* __restrict__ aux_map, int* __restrict__ slow_flag, int H,
                 int W, unsigned long long* __restrict__ stat, int force_slow_first) {
  extern __shared__ __align__(16) uint8_t ws_smem[];
  __shared__ int s_warp_cnt[kWcThreads / 32], s_warp_sz[kWcThreads / 32];
  __shared__ int s_ncomp, s_total, s_unsafe;
  const int hw = H * W;
  uint64_t* pool = reinterpret_cast<uint64_t*>(ws_smem);
  uint16_t* lab16 = reinterpret_cast<uint16_t*>(ws_smem + sizeof(uint64_t) * kWcPool);
  int* c_off = reinterpret_cast<int*>(ws_smem + sizeof(uint64_t) * kWcPool + 2u * ((hw + 7) & ~7));
  int* c_cnt = c_off + kWcMaxComp;
  // marker label of the component's boundary markers, or -1 once two different labels were seen.
  // A component whose boundary markers all carry ONE label is flooded with that label whatever the
  // pop order, so a tie between its markers cannot change the result and needs no fallback.
  int* c_lab = c_cnt + kWcMaxComp;
  const size_t base = static_cast<size_t>(blockIdx.x) * hw;
  const float* v = val + base;
  const uint8_t* m = msk + base;
  const int* L = Lroot + base;
  const int* sz = csize + base;
  int* o = out + base;
  int* amap = aux_map + base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { s_ncomp = 0; s_total = 0; s_unsafe = 0; }
  // A: stage the label map; clear the per-root "has a marker" flag
  for (int p = tid; p < hw; p += kWcThreads) {
    const bool in = m[p] != 0;
    lab16[p] = in ? static_cast<uint16_t>(o[p]) : kWsOutside;
    if (in && L[p] == p) amap[p] = 0;
  }
  __syncthreads();
  for (int p = tid; p < hw; p += kWcThreads) {
    const uint16_t l = lab16[p];
    if (l != 0 && l != kWsOutside) amap[L[p]] = 1;
  }
  __syncthreads();
  // B: number the components that own markers (raster order) and carve the heap pool
  for (int start = 0; start < hw; start += kWcThreads) {
    const int p = start + tid;
    int flag = 0, size = 0;
    if (p < hw && m[p] && L[p] == p) {
      if (amap[p]) { flag = 1; size = sz[p]; } else amap[p] = -1;
    }
    int c = flag, a = size;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int tc = __shfl_up_sync(0xffffffffu, c, d), ta = __shfl_up_sync(0xffffffffu, a, d);
      if (lane >= d) { c += tc; a += ta; }
    }
    if (lane == 31) { s_warp_cnt[warp] = c; s_warp_sz[warp] = a; }
    __syncthreads();
    int bc = s_ncomp, ba = s_total;
    for (int w2 = 0; w2 < warp; ++w2) { bc += s_warp_cnt[w2]; ba += s_warp_sz[w2]; }
    if (flag) {
      const int k = bc + c - 1;
      if (k < kWcMaxComp) { c_off[k] = ba + a - size; c_cnt[k] = 0; c_lab[k] = 0; }
      amap[p] = k;
    }
    __syncthreads();
    if (tid == 0) {
      int tc = 0, ta = 0;
      for (int w2 = 0; w2 < kWcThreads / 32; ++w2) { tc += s_warp_cnt[w2]; ta += s_warp_sz[w2]; }
      s_ncomp += tc;
      s_total += ta;
    }
    __syncthreads();
  }
  const int ncomp = s_ncomp;
  const bool fits = ncomp <= kWcMaxComp && s_total <= kWcPool;
  if (fits) {
    // C: boundary markers go straight into their component's heap slice (arbitrary order)
    for (int p = tid; p < hw; p += kWcThreads) {
      const uint16_t l = lab16[p];
      if (l == 0 || l == kWsOutside) continue;
      const int x = p % W;
      const bool b = (p >= W && lab16[p - W] == 0) || (x > 0 && lab16[p - 1] == 0) ||
                     (x < W - 1 && lab16[p + 1] == 0) || (p + W < hw && lab16[p + W] == 0);
      if (!b) continue;
      const int k = amap[L[p]];
      const int slot = c_off[k] + atomicAdd(&c_cnt[k], 1);
      pool[slot] = ws_entry(v[p], 0u, p);
      const int prev = atomicCAS(&c_lab[k], 0, static_cast<int>(l));
      if (prev != 0 && prev != static_cast<int>(l)) c_lab[k] = -1;
    }
    __syncthreads();
    // D: one thread per component
    for (int k = tid; k < ncomp; k += kWcThreads) {
      uint64_t* hp = pool + c_off[k];
      int n = c_cnt[k];
      // The queue is a 4-ARY heap: within a component every (value, age) key is unique except
      // between tied marker entries, whose relative order is either proven irrelevant or sends the
      // tile to the exact emulation - so ANY correct priority queue pops in the reference's order,
      // and a 4-ary sift-down reads four children at once (one shared-memory latency per level,
      // half as many levels as the binary heap the exact kernels must emulate).
      for (int j = 1; j < n; ++j) {  // in-place build by successive pushes
        const uint64_t e = hp[j];
        const uint64_t ek = e >> 16;
        int c = j;
        while (c > 0) {
          const int parent = (c - 1) >> 2;
          const uint64_t pe = hp[parent];
          if (!(ek < (pe >> 16))) break;
          hp[c] = pe;
          c = parent;
        }
        hp[c] = e;
      }
      uint32_t age = 1;
      uint64_t last_marker = ~0ull;
      bool tie = false;
      // current group of tied marker entries, see ws_tie_harmless
      int g_pix0 = -1, g_pix1 = -1, g_n = 0;
      unsigned g_mask0 = 0, g_mask1 = 0;
      bool g_clean = false;
      while (n > 0) {
        const uint64_t top = hp[0];
        const int ei = static_cast<int>(top & 0xFFFFu);
        const bool is_marker = ((top >> 16) & 0xFFFFu) == 0;  // age 0
        bool tie_now = false;
        if (is_marker) {
          if ((top >> 32) == last_marker && c_lab[k] == -1) {
            if (!g_clean || g_n >= 3) { tie = true; break; }
            tie_now = true;
          } else {
            g_n = 0;
            g_clean = true;
          }
          last_marker = top >> 32;
        } else {
          g_clean = false;
        }
        const int x = ei % W;
        const uint16_t lab = lab16[ei];
        const int q0 = ei - W, q1 = ei - 1, q2 = ei + 1, q3 = ei + W;
        const bool c0 = q0 >= 0 && lab16[q0] == 0;
        const bool c1 = x > 0 && lab16[q1] == 0;
        const bool c2 = x < W - 1 && lab16[q2] == 0;
        const bool c3 = q3 < hw && lab16[q3] == 0;
        if (is_marker) {
          const unsigned e_mask = (c0 ? 1u : 0u) | (c1 ? 2u : 0u) | (c2 ? 4u : 0u) | (c3 ? 8u : 0u);
          if (tie_now && (!ws_tie_harmless(v, W, g_pix0, g_mask0, ei, e_mask) ||
                          (g_n == 2 && !ws_tie_harmless(v, W, g_pix1, g_mask1, ei, e_mask)))) {
            tie = true;
            break;
          }
          if (g_n == 0) { g_pix0 = ei; g_mask0 = e_mask; }
          else if (g_n == 1) { g_pix1 = ei; g_mask1 = e_mask; }
          ++g_n;
        }
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
        if (c0) v0 = __ldg(v + q0);
        if (c1) v1 = __ldg(v + q1);
        if (c2) v2 = __ldg(v + q2);
        if (c3) v3 = __ldg(v + q3);
        // remove the root
        --n;
        if (n > 0) {
          const uint64_t xe = hp[n];
          const uint64_t xk = xe >> 16;
          int i = 0;
          for (;;) {
            const int c0 = 4 * i + 1;
            if (c0 >= n) break;
            const uint64_t e0 = hp[c0];
            const uint64_t e1 = (c0 + 1 < n) ? hp[c0 + 1] : ~0ull;
            const uint64_t e2 = (c0 + 2 < n) ? hp[c0 + 2] : ~0ull;
            const uint64_t e3 = (c0 + 3 < n) ? hp[c0 + 3] : ~0ull;
            int sidx = c0;
            uint64_t se = e0;
            if ((e1 >> 16) < (se >> 16)) { sidx = c0 + 1; se = e1; }
            if ((e2 >> 16) < (se >> 16)) { sidx = c0 + 2; se = e2; }
            if ((e3 >> 16) < (se >> 16)) { sidx = c0 + 3; se = e3; }
            if (!((se >> 16) < xk)) break;
            hp[i] = se;
            i = sidx;
          }
          hp[i] = xe;
        }
#define CERB_WC_PUSH(cond, vv, qq)                                  \
        if (cond) {                                                     \
          ++age;                                                        \
          lab16[qq] = lab;                                              \
          /* the rows above / below will be read when qq is popped */   \
          if (qq >= W) asm volatile("prefetch.global.L1 [%0];" ::"l"(v + qq - W));       \
          if (qq + W < hw) asm volatile("prefetch.global.L1 [%0];" ::"l"(v + qq + W));   \
          const uint64_t e = ws_entry(vv, age, qq);                     \
          const uint64_t ek = e >> 16;                                  \
          int c = n++;                                                  \
          while (c > 0) {                                               \
            const int parent = (c - 1) >> 2;                            \
            const uint64_t pe = hp[parent];                             \
            if (!(ek < (pe >> 16))) break;                              \
            hp[c] = pe;                                                 \
            c = parent;                                                 \
          }                                                             \
          hp[c] = e;                                                    \
        }
        CERB_WC_PUSH(c0, v0, q0)
        CERB_WC_PUSH(c1, v1, q1)
        CERB_WC_PUSH(c2, v2, q2)
        CERB_WC_PUSH(c3, v3, q3)
#undef CERB_WC_PUSH
      }
      if (tie) s_unsafe = 1;
    }
  }
  __syncthreads();
  if (!fits || s_unsafe || (force_slow_first && blockIdx.x == 0)) {
    if (tid == 0) {
      slow_flag[blockIdx.x] = 1;  // `out` still holds markers * mask
      if (stat != nullptr) { atomicAdd(stat + 0, 1ull); atomicAdd(stat + (fits ? 2 : 3), 1ull); }
    }
    return;
  }
  if (tid == 0) {
    slow_flag[blockIdx.x] = 0;
    if (stat != nullptr) atomicAdd(stat + 0, 1ull);
  }
  for (int p = tid; p < hw; p += kWcThreads) {
    const uint16_t l = lab16[p];
    o[p] = l == kWsOutside ? 0 : static_cast<int>(l);
  }
}

// ---- component-parallel path for LARGE images (WSI post-processing tiles, infer/wsi.py:137-150:
// up to 4032 x 4032 pixels). Same independence argument as k_watershed_comp, but nothing fits in
// shared memory: the label map stays int32 in global memory and every mask component floods with
// a private binary heap carved out of one global pool (a component of s pixels pushes at most s
// entries, so the pool needs at most hw entries). One thread per component: tens of thousands of
// nuclei per tile keep the GPU busy where the whole-image emulation runs on a single lane.
// A tie between two marker entries of one component (bit-equal values, age 0) is decided by the
// GLOBAL heap layout in the reference; it is detected (consecutive marker pops with equal values)
// and the tile is then redone, from the saved markers, by the exact whole-image kernel.
__device__ __forceinline__ uint64_t wsg_key(float v, uint32_t age) {
  uint32_t u = __float_as_uint(v);
  if ((u << 1) == 0) u = 0;  // -0.0 == +0.0
  u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;
  return (static_cast<uint64_t>(u) << 32) | age;
}

// ctl[0] = pool top, ctl[1] = number of components, ctl[2] = "redo the image with the exact
// whole-image kernel", ctl[3] = components that saw a marker tie (per image: 4 ints)
constexpr int kWsgTied = -2;  // clab value of a component waiting for k_wsg_certify
__global__ void k_wsg_alloc(const uint8_t* __restrict__ msk, const int* __restrict__ L,
                            const int* __restrict__ size, int* __restrict__ off, int* __restrict__ cnt,
                            int* __restrict__ clab, int* __restrict__ roots, int* __restrict__ ctl,
                            int hw) {
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  int* c = ctl + 4 * blockIdx.y;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    if (!msk[base + p] || L[base + p] != p) continue;
    off[base + p] = atomicAdd(&c[0], size[base + p]);
    cnt[base + p] = 0;
    clab[base + p] = 0;
    roots[base + atomicAdd(&c[1], 1)] = p;
  }
}

__global__ void k_wsg_seed(const float* __restrict__ val, const uint8_t* __restrict__ msk,
                           const int* __restrict__ L, const int* __restrict__ lab,
                           const int* __restrict__ off, int* __restrict__ cnt, int* __restrict__ clab,
                           unsigned long long* __restrict__ hk, int* __restrict__ hi, int H, int W,
                           const int* __restrict__ ctl, int tied_only) {
  const int hw = H * W;
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  const uint8_t* m = msk + base;
  const int* l = lab + base;
  if (tied_only && ctl[4 * blockIdx.y + 3] == 0) return;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    if (!m[p] || l[p] == 0) continue;
    if (tied_only && clab[base + L[base + p]] != kWsgTied) continue;
    const int x = p % W;
    const bool b = (p >= W && m[p - W] && l[p - W] == 0) || (x > 0 && m[p - 1] && l[p - 1] == 0) ||
                   (x < W - 1 && m[p + 1] && l[p + 1] == 0) || (p + W < hw && m[p + W] && l[p + W] == 0);
    if (!b) continue;  // a marker pixel without an unlabeled in-mask neighbour pushes nothing
    const int root = L[base + p];
    const int slot = off[base + root] + atomicAdd(&cnt[base + root], 1);
    hk[base + slot] = wsg_key(val[base + p], 0u);
    hi[base + slot] = p;
    if (tied_only) continue;
    // one label per component -> ties are harmless (see k_watershed_comp)
    const int prev = atomicCAS(&clab[base + root], 0, l[p]);
    if (prev != 0 && prev != l[p]) clab[base + root] = -1;
  }
}

// Floods one mask component from the n entries in its heap slice (k, ix). CERT = false: the
// regular pass - marker entries carry age 0 and the function returns false as soon as two marker
// entries with bit-equal values surface in a component whose markers carry different labels.
// CERT = true (k_wsg_certify): the age field of the marker entries holds an explicit rank, pushed
// entries get ages above every rank, every labelled pixel is appended to `log`.
#ifndef CERB_WSG_ARITY
#define CERB_WSG_ARITY 4
#endif
constexpr int kWsgArity = CERB_WSG_ARITY;  // children per node of the global-memory flood queues

template <bool CERT>
__device__ __forceinline__ bool wsg_flood_component(const float* __restrict__ v,
                                                    const uint8_t* __restrict__ m, int* __restrict__ o,
                                                    int W, int hw, unsigned long long* __restrict__ k,
                                                    int* __restrict__ ix, int n, bool multi,
                                                    uint32_t age, int* __restrict__ log, int& nlog) {
  // 4-ary heap (see k_watershed_comp: any correct priority queue gives the reference's order here)
  for (int j = 1; j < n; ++j) {  // in-place heap build by successive pushes
    const unsigned long long ek = k[j];
    const int ei = ix[j];
    int cidx = j;
    while (cidx > 0) {
      const int parent = (cidx - 1) / kWsgArity;
      if (!(ek < k[parent])) break;
      k[cidx] = k[parent];
      ix[cidx] = ix[parent];
      cidx = parent;
    }
    k[cidx] = ek;
    ix[cidx] = ei;
  }
  unsigned long long last_marker = ~0ull;
  // current group of tied marker entries (first two members) and whether only its members left
  // the queue since it started: see ws_tie_harmless
  int g_pix0 = -1, g_pix1 = -1, g_n = 0;
  unsigned g_mask0 = 0, g_mask1 = 0;
  bool g_clean = false, tie_now = false;
  while (n > 0) {
    const unsigned long long top = k[0];
    const int ei = ix[0];
    if (!CERT) {
      tie_now = false;
      if ((top & 0xFFFFFFFFull) == 0) {  // a marker entry
        if ((top >> 32) == last_marker && multi) {
          if (!g_clean || g_n >= 3) return false;
          tie_now = true;
        } else {
          g_n = 0;
          g_clean = true;
        }
        last_marker = top >> 32;
      } else {
        g_clean = false;
      }
    }
    unsigned e_mask = 0;
    --n;
    if (n > 0) {  // move the last entry to the root and sift down
      const unsigned long long xk = k[n];
      const int xi = ix[n];
      int i = 0;
      for (;;) {
        const int c0 = kWsgArity * i + 1;
        if (c0 >= n) break;
        int sidx = c0;
        unsigned long long sk = k[c0];
#pragma unroll
        for (int d = 1; d < kWsgArity; ++d) {
          const unsigned long long kd = (c0 + d < n) ? k[c0 + d] : ~0ull;
          if (kd < sk) { sidx = c0 + d; sk = kd; }
        }
        if (!(sk < xk)) break;
        k[i] = sk;
        ix[i] = ix[sidx];
        i = sidx;
      }
      k[i] = xk;
      ix[i] = xi;
    }
    const int lb = o[ei];
    const int x = ei % W;
#define CERB_WSG_PUSH(cond, qq, bit)                                     \
    if (cond) {                                                          \
      const int q = (qq);                                                \
      if (m[q] && o[q] == 0) {                                           \
        ++age;                                                           \
        o[q] = lb;                                                       \
        e_mask |= (bit);                                                 \
        if (CERT) log[nlog++] = q;                                       \
        const unsigned long long ek = wsg_key(v[q], age);                \
        int cidx = n++;                                                  \
        while (cidx > 0) {                                               \
          const int parent = (cidx - 1) / kWsgArity;                     \
          if (!(ek < k[parent])) break;                                  \
          k[cidx] = k[parent];                                           \
          ix[cidx] = ix[parent];                                         \
          cidx = parent;                                                 \
        }                                                                \
        k[cidx] = ek;                                                    \
        ix[cidx] = q;                                                    \
      }                                                                  \
    }
    CERB_WSG_PUSH(ei >= W, ei - W, 1u)
    CERB_WSG_PUSH(x > 0, ei - 1, 2u)
    CERB_WSG_PUSH(x < W - 1, ei + 1, 4u)
    CERB_WSG_PUSH(ei + W < hw, ei + W, 8u)
#undef CERB_WSG_PUSH
    if (!CERT && (top & 0xFFFFFFFFull) == 0) {
      if (tie_now) {  // pairwise against the earlier members of the group
        if (!ws_tie_harmless(v, W, g_pix0, g_mask0, ei, e_mask)) return false;
        if (g_n == 2 && !ws_tie_harmless(v, W, g_pix1, g_mask1, ei, e_mask)) return false;
      }
      if (g_n == 0) { g_pix0 = ei; g_mask0 = e_mask; }
      else if (g_n == 1) { g_pix1 = ei; g_mask1 = e_mask; }
      ++g_n;
    }
  }
  return true;
}

__global__ void k_wsg_flood(const float* __restrict__ val, const uint8_t* __restrict__ msk,
                            int* __restrict__ lab, const int* __restrict__ off,
                            const int* __restrict__ cnt, int* __restrict__ clab,
                            const int* __restrict__ roots, unsigned long long* __restrict__ hk,
                            int* __restrict__ hi, int* __restrict__ ctl, int H, int W) {
  const int hw = H * W;
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  int* c = ctl + 4 * blockIdx.y;
  const int n_roots = c[1];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_roots; t += gridDim.x * blockDim.x) {
    const int root = roots[base + t];
    const int n = cnt[base + root];
    if (n == 0) continue;
    int nlog = 0;
    if (!wsg_flood_component<false>(val + base, msk + base, lab + base, W, hw,
                                    hk + base + off[base + root], hi + base + off[base + root], n,
                                    clab[base + root] == -1, 1u, nullptr, nlog)) {
      clab[base + root] = kWsgTied;  // abandoned half-flooded: k_wsg_certify redoes it
      atomicAdd(&c[3], 1);
    }
  }
}

// ---- marker ties without the whole-image emulation ---------------------------------------------
// Entries with bit-equal (value, age 0) keys leave the reference's heap in an order fixed by its
// global array layout; everything else in a component follows the keys alone. So the true run of
// a tied component is "key order, with SOME order inside each group of tied marker entries". If
// every such order yields the same labels, that labelling is the reference's, whatever its heap
// did. k_wsg_certify floods a tied component once per combination of permutations of its tie
// groups (marker ranks in the age field) and compares the label sets; only when two orders
// disagree - or there are more than kTieMaxVariants of them (plateaus) - is the image handed to the
// exact whole-image kernel.
constexpr int kTieMaxVariants = 24;
constexpr int kTieMaxGroups = 4;
constexpr int kTieMaxSeeds = 1 << 20;

// markers back, flood undone, seed counters cleared for the tied components
__global__ void k_wsg_tied_reset(const uint8_t* __restrict__ msk, const int* __restrict__ L,
                                 const int* __restrict__ clab, const int* __restrict__ saved,
                                 int* __restrict__ lab, int* __restrict__ cnt,
                                 const int* __restrict__ ctl, int hw) {
  if (ctl[4 * blockIdx.y + 3] == 0) return;
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    if (!msk[base + p]) continue;
    const int root = L[base + p];
    if (clab[base + root] != kWsgTied) continue;
    lab[base + p] = saved[base + p];
    if (p == root) cnt[base + root] = 0;
  }
}

__global__ void k_wsg_certify(const float* __restrict__ val, const uint8_t* __restrict__ msk,
                              int* __restrict__ lab, const int* __restrict__ off,
                              const int* __restrict__ cnt, const int* __restrict__ clab,
                              const int* __restrict__ roots, unsigned long long* __restrict__ hk,
                              int* __restrict__ hi, int* __restrict__ ctl, int* __restrict__ saved,
                              int* __restrict__ logbuf, int* __restrict__ seedbuf, int H, int W) {
  const int hw = H * W;
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  int* c = ctl + 4 * blockIdx.y;
  if (c[3] == 0) return;
  const float* v = val + base;
  int* o = lab + base;
  int* sv = saved + base;
  const int n_roots = c[1];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_roots; t += gridDim.x * blockDim.x) {
    const int root = roots[base + t];
    if (clab[base + root] != kWsgTied) continue;
    const int n = cnt[base + root];
    unsigned long long* k = hk + base + off[base + root];
    int* ix = hi + base + off[base + root];
    int* lg = logbuf + base + off[base + root];
    int* sd = seedbuf + base + off[base + root];
    if (n > kTieMaxSeeds) { atomicExch(&c[2], 1); continue; }
    // seeds sorted by (value, pixel): tie groups become runs. In-place heapsort of 64-bit keys
    // (value bits : pixel) in the component's heap slice - a cluster can have thousands of seeds.
    for (int j = 0; j < n; ++j) k[j] = wsg_key(v[ix[j]], 0u) | static_cast<unsigned>(ix[j]);
    for (int start = n / 2 - 1, end = n; ; ) {
      unsigned long long x;
      if (start >= 0) {
        x = k[start];  // heapify phase
      } else {
        if (--end <= 0) break;
        x = k[end];  // extraction phase: the maximum moves behind the shrinking heap
        k[end] = k[0];
      }
      int i = start >= 0 ? start : 0;
      for (;;) {  // sift x down a max-heap of `end` entries
        int ch = 2 * i + 1;
        if (ch >= end) break;
        if (ch + 1 < end && k[ch + 1] > k[ch]) ++ch;
        if (!(k[ch] > x)) break;
        k[i] = k[ch];
        i = ch;
      }
      k[i] = x;
      if (start >= 0) --start;
    }
    for (int j = 0; j < n; ++j) sd[j] = static_cast<int>(k[j] & 0xFFFFFFFFull);
    int g_start[kTieMaxGroups], g_len[kTieMaxGroups], n_groups = 0;
    long long variants = 1;
    for (int j = 0; j < n && variants > 0;) {
      int e = j + 1;
      while (e < n && wsg_key(v[sd[e]], 0u) == wsg_key(v[sd[j]], 0u)) ++e;
      if (e - j > 1) {
        if (n_groups == kTieMaxGroups) { variants = -1; break; }
        g_start[n_groups] = j;
        g_len[n_groups] = e - j;
        ++n_groups;
        for (int f = 2; f <= e - j && variants > 0; ++f) {
          variants *= f;
          if (variants > kTieMaxVariants) variants = -1;
        }
      }
      j = e;
    }
    if (variants < 0) { atomicExch(&c[2], 1); continue; }
    bool same = true;
    for (int var = 0; var < static_cast<int>(variants) && same; ++var) {
      for (int j = 0; j < n; ++j) {
        k[j] = wsg_key(v[sd[j]], static_cast<uint32_t>(j));
        ix[j] = sd[j];
      }
      int code = var;
      for (int g = 0; g < n_groups; ++g) {  // var -> one permutation per group (Lehmer code)
        int fact = 1;
        for (int f = 2; f <= g_len[g]; ++f) fact *= f;
        int idx = code % fact;
        code /= fact;
        unsigned used = 0;
        for (int i = 0; i < g_len[g]; ++i) {
          fact /= (g_len[g] - i);
          int sel = idx / fact;
          idx %= fact;
          int r = 0;
          for (;; ++r) {
            if (used & (1u << r)) continue;
            if (sel-- == 0) break;
          }
          used |= 1u << r;
          k[g_start[g] + i] = wsg_key(v[sd[g_start[g] + i]], static_cast<uint32_t>(g_start[g] + r));
        }
      }
      int nlog = 0;
      wsg_flood_component<true>(v, msk + base, o, W, hw, k, ix, n, true, static_cast<uint32_t>(n), lg,
                                nlog);
      if (var == 0) {
        for (int j = 0; j < nlog; ++j) sv[lg[j]] = -o[lg[j]];  // non-marker pixels: saved == 0 so far
      } else {
        for (int j = 0; j < nlog && same; ++j) same = (o[lg[j]] == -sv[lg[j]]);
      }
      if (same && var + 1 < static_cast<int>(variants))
        for (int j = 0; j < nlog; ++j) o[lg[j]] = 0;
    }
    if (!same) atomicExch(&c[2], 1);
  }
}

// tie seen: put the saved markers back so that the exact whole-image kernel starts from them
__global__ void k_wsg_restore(int* __restrict__ lab, const int* __restrict__ saved,
                              const int* __restrict__ ctl, int hw) {
  if (ctl[4 * blockIdx.y + 2] == 0) return;
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x)
    lab[base + p] = max(saved[base + p], 0);  // k_wsg_certify parks its labels there as negatives
}

__global__ void k_wsg_flag(const int* __restrict__ ctl, int* __restrict__ run_flag, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) run_flag[i] = ctl[4 * i + 2];
}

// ------------------------------------------------------------------ gland / lumen
// loader/postproc.py:277-286 / :319-327
// single != 0: PostProcInstErodedMap (loader/postproc.py:154-155 etc.), one channel > thr
__global__ void k_gl_threshold(const float* __restrict__ canvas, int C, int ch0, float thr,
                               uint8_t* __restrict__ fg, int hw, int single) {
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    const float inner = canvas[(base + p) * C + ch0];
    if (single) {
      fg[base + p] = inner > thr;
      continue;
    }
    const float cnt = canvas[(base + p) * C + ch0 + 1];
    const float c01 = cnt > 0.5f ? 1.0f : 0.0f;
    fg[base + p] = __fsub_rn(inner, c01) > thr;
  }
}

// bb: [n][max_inst][4] = ymin, ymax, xmin, xmax (inclusive)
__global__ void k_bbox_init(int* __restrict__ bb, int total) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    bb[i] = (i & 1) ? -1 : INT_MAX;
}

__global__ void k_bbox(const int* __restrict__ lab, int* __restrict__ bb, int H, int W, int max_inst,
                       int* __restrict__ err) {
  const int hw = H * W;
  const size_t base = static_cast<size_t>(blockIdx.y) * hw;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    const int id = lab[base + p];
    if (id == 0) continue;
    if (id > max_inst) { atomicExch(err, 10); continue; }
    int* b = bb + (static_cast<size_t>(blockIdx.y) * max_inst + (id - 1)) * 4;
    const int x = p % W, y = p / W;
    atomicMin(&b[0], y);
    atomicMax(&b[1], y);
    atomicMin(&b[2], x);
    atomicMax(&b[3], x);
  }
}

struct EllipseRows {
  int k;
  int j1[32];
  int j2[32];
};

// 32 bits of a bit-row starting at bit position `bit` (may be negative / past the end).
__device__ __forceinline__ uint32_t bits_at(const uint32_t* row, int nwords, int bit) {
  const int w = bit >> 5;  // floor
  const int s = bit & 31;
  const uint32_t lo = (w >= 0 && w < nwords) ? row[w] : 0u;
  const uint32_t hi = (w + 1 >= 0 && w + 1 < nwords) ? row[w + 1] : 0u;
  return s == 0 ? lo : ((lo >> s) | (hi << (32 - s)));
}

// One block per (instance, image): crop (bbox +- 2k unless that would cross the border),
// dilate inside the crop, fill holes inside the crop, paint max id.
__global__ void __launch_bounds__(kThreads)
k_gl_instance(const int* __restrict__ lab, const int* __restrict__ count,
              const int* __restrict__ has_bg, const int* __restrict__ bb,
              int* __restrict__ out, int H, int W, int max_inst, EllipseRows ell, int smem_words,
              uint32_t* __restrict__ gpool, unsigned int gpool_words, unsigned int* __restrict__ gpool_top,
              int* __restrict__ err) {
  extern __shared__ uint32_t bitmem[];
  __shared__ unsigned int s_goff;
  const int img = blockIdx.y;
  const int id = blockIdx.x + 1;
  if (id > count[img]) return;
  // Reference quirk (loader/postproc.py:291,332): np.unique(inst_lab)[1:] drops the smallest
  // label assuming it is the background; an image without any background pixel loses id 1.
  if (id == 1 && !has_bg[img]) return;
  const int hw = H * W;
  const int* L = lab + static_cast<size_t>(img) * hw;
  int* O = out + static_cast<size_t>(img) * hw;
  const int* b = bb + (static_cast<size_t>(img) * max_inst + (id - 1)) * 4;
  // get_bounding_box (misc/utils.py:82-91): max is exclusive
  int y1 = b[0], y2 = b[1] + 1, x1 = b[2], x2 = b[3] + 1;
  const int k = ell.k;
  const int pad = k * 2;
  if (y1 - pad >= 0) y1 -= pad;
  if (x1 - pad >= 0) x1 -= pad;
  if (x2 + pad <= W - 1) x2 += pad;
  if (y2 + pad <= H - 1) y2 += pad;
  const int ch = y2 - y1, cw = x2 - x1;
  const int nw = (cw + 31) >> 5;
  // Crops whose two bit planes do not fit in shared memory (merged glands of ~900 px and more)
  // take a slice of a global-memory pool instead: same code, the planes just live in L2 / HBM.
  uint32_t* planes = bitmem;
  if (2 * ch * nw > smem_words) {
    const unsigned int need = 2u * static_cast<unsigned int>(ch) * static_cast<unsigned int>(nw);
    if (threadIdx.x == 0) s_goff = atomicAdd(gpool_top, need);
    __syncthreads();
    if (gpool == nullptr || s_goff + need > gpool_words) {
      if (threadIdx.x == 0) atomicExch(err, 11);
      return;
    }
    planes = gpool + s_goff;
  }
  uint32_t* A = planes;            // instance bits, later the "outside" flood
  uint32_t* B = planes + ch * nw;  // dilated bits
  const int total = ch * nw;
  const uint32_t tail = (cw & 31) ? ((1u << (cw & 31)) - 1u) : 0xffffffffu;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int y = i / nw, wi = i % nw;
    uint32_t bits = 0;
    const int xb = wi * 32;
    const int lim = min(32, cw - xb);
    const int* row = L + static_cast<size_t>(y1 + y) * W + x1 + xb;
    for (int j = 0; j < lim; ++j) bits |= (row[j] == id ? 1u : 0u) << j;
    A[i] = bits;
  }
  __syncthreads();
  // cv2.dilate, anchor = (k/2, k/2): dst(y,x) = OR src(y + ky - a, x + kx - a) over the element
  const int a = k / 2;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int y = i / nw, wi = i % nw;
    uint32_t acc = 0;
    for (int ky = 0; ky < k; ++ky) {
      const int ry = y + ky - a;
      if (ry < 0 || ry >= ch) continue;
      const uint32_t* row = A + ry * nw;
      for (int kx = ell.j1[ky]; kx < ell.j2[ky]; ++kx) acc |= bits_at(row, nw, wi * 32 + kx - a);
    }
    if (wi == nw - 1) acc &= tail;
    B[i] = acc;
  }
  __syncthreads();
  // flood the background from the crop border (4-connectivity); A becomes "outside"
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int y = i / nw, wi = i % nw;
    uint32_t bg = ~B[i];
    if (wi == nw - 1) bg &= tail;
    uint32_t border = 0;
    if (y == 0 || y == ch - 1) border = 0xffffffffu;
    if (wi == 0) border |= 1u;
    if (wi == nw - 1) border |= 1u << ((cw - 1) & 31);
    A[i] = bg & border;
  }
  __syncthreads();
  for (;;) {
    int changed = 0;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int y = i / nw, wi = i % nw;
      uint32_t bg = ~B[i];
      if (wi == nw - 1) bg &= tail;
      const uint32_t cur = A[i];
      uint32_t nb = cur;
      if (y > 0) nb |= A[i - nw];
      if (y < ch - 1) nb |= A[i + nw];
      if (wi > 0) nb |= A[i - 1] >> 31;
      if (wi < nw - 1) nb |= A[i + 1] << 31;
      nb &= bg;
      // run the horizontal propagation inside the word to convergence
      for (int it = 0; it < 32; ++it) {
        const uint32_t g = (nb | (nb << 1) | (nb >> 1)) & bg;
        if (g == nb) break;
        nb = g;
      }
      if (nb != cur) {
        A[i] = nb;  // monotone growth: a racing neighbour read sees the old or the new value
        changed = 1;
      }
    }
    if (!__syncthreads_or(changed)) break;
  }
  // result = everything that is not outside; ascending-id painting == max id
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int y = i / nw, wi = i % nw;
    uint32_t in = ~A[i];
    if (wi == nw - 1) in &= tail;
    int* row = O + static_cast<size_t>(y1 + y) * W + x1 + wi * 32;
    while (in) {
      const int j = __ffs(in) - 1;
      in &= in - 1;
      atomicMax(&row[j], id);
    }
  }
}

__global__ void k_mask_by(int* __restrict__ lumen, const int* __restrict__ gland, size_t total) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    if (gland[i] <= 0) lumen[i] = 0;
}

// ------------------------------------------------------------------ host side
struct Workspace {
  size_t pixels = 0;  // n * hw capacity
  uint8_t *m0 = nullptr, *m1 = nullptr, *m2 = nullptr;
  int *L = nullptr, *size = nullptr, *rank = nullptr, *lab = nullptr, *aux = nullptr;
  float *val = nullptr, *heap_v = nullptr;
  int *heap_a = nullptr, *heap_i = nullptr;
  unsigned long long* heap_k = nullptr;
  int *count = nullptr, *any_fg = nullptr;
  int* ctl = nullptr;  // [n][4] large-image watershed: pool top, component count, redo flag, tied components
  int *tie_log = nullptr, *tie_seeds = nullptr;  // large images only: k_wsg_certify scratch planes
  size_t tie_cap = 0;
  float* canvas = nullptr;
  size_t canvas_elems = 0;
  int* bb = nullptr;
  size_t bb_elems = 0;
  size_t count_cap = 0;
};

Workspace* g_ws_for(cerb_ctx* ctx);

template <typename T>
cudaError_t grow(cerb_ctx* ctx, T*& p, size_t& cap_elems, size_t need) {
  if (need <= cap_elems && p != nullptr) return cudaSuccess;
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, need * sizeof(T));
  if (e != cudaSuccess) return e;
  if (p != nullptr) {
    // the outgrown buffer may still be read by queued kernels: cudaFree synchronises the device
    for (size_t i = 0; i < ctx->scratch.size(); ++i)
      if (ctx->scratch[i] == p) {
        ctx->scratch.erase(ctx->scratch.begin() + i);
        break;
      }
    cudaFree(p);
  }
  ctx->scratch.push_back(q);
  p = static_cast<T*>(q);
  cap_elems = need;
  return cudaSuccess;
}

// The workspace lives and dies with its ctx (a table keyed by the ctx address would hand a stale
// workspace, full of freed pointers, to a later ctx allocated at the same address).
Workspace* g_ws_for(cerb_ctx* ctx) {
  if (ctx->postproc_ws == nullptr) {
    ctx->postproc_ws = new Workspace();
    ctx->postproc_ws_free = [](void* w) { delete static_cast<Workspace*>(w); };
  }
  return static_cast<Workspace*>(ctx->postproc_ws);
}

int ensure_ws(cerb_ctx* ctx, Workspace*& ws, int n, int hw) {
  ws = g_ws_for(ctx);
  if (!ws) return fail(CERB_ERR_ARG, "postproc: too many contexts");
  const size_t need = static_cast<size_t>(n) * hw;
  if (need > ws->pixels) {
    size_t c;
#define GROW(field, type)                                                              \
  c = 0;                                                                               \
  CERB_CUDA(grow<type>(ctx, ws->field, c, need));
    GROW(m0, uint8_t) GROW(m1, uint8_t) GROW(m2, uint8_t)
    GROW(L, int) GROW(size, int) GROW(rank, int) GROW(lab, int) GROW(aux, int)
    GROW(val, float) GROW(heap_v, float) GROW(heap_a, int) GROW(heap_i, int)
    GROW(heap_k, unsigned long long)
#undef GROW
    ws->pixels = need;
  }
  if (static_cast<size_t>(n) > ws->count_cap) {
    size_t c = 0;
    CERB_CUDA(grow<int>(ctx, ws->count, c, static_cast<size_t>(n)));
    c = 0;
    CERB_CUDA(grow<int>(ctx, ws->any_fg, c, static_cast<size_t>(n)));
    c = 0;
    CERB_CUDA(grow<int>(ctx, ws->ctl, c, static_cast<size_t>(n) * 4));
    ws->count_cap = n;
  }
  return CERB_OK;
}

dim3 grid2(int hw, int n) {
  int gx = (hw + kThreads - 1) / kThreads;
  if (gx > 148 * 4) gx = 148 * 4;
  return dim3(gx, n);
}

// CC of `fg` (in place size filter when min_size > 0). Leaves roots in ws->L, sizes in ws->size.
void cc_label(cerb_ctx* ctx, Workspace* ws, uint8_t* fg, int n, int H, int W, int min_size) {
  const int hw = H * W;
  cudaStream_t s = ctx->stream;
  k_cc_init<<<grid2(hw, n), kThreads, 0, s>>>(fg, ws->L, ws->size, hw, W);
  k_cc_merge<<<grid2(hw, n), kThreads, 0, s>>>(fg, ws->L, H, W);
  k_cc_flatten_count<<<grid2(hw, n), kThreads, 0, s>>>(fg, ws->L, ws->size, hw);
  ctx->launches += 3;
  if (min_size > 0) {
    k_filter_small<<<grid2(hw, n), kThreads, 0, s>>>(fg, ws->L, ws->size, hw, min_size);
    ctx->launches += 1;
  }
}

int stage_canvas(cerb_ctx* ctx, Workspace* ws, const float* canvas, size_t elems, int on_device,
                 const float** dev) {
  if (on_device) {
    *dev = canvas;
    return CERB_OK;
  }
  CERB_CUDA(grow<float>(ctx, ws->canvas, ws->canvas_elems, elems));
  CERB_CUDA(cudaMemcpyAsync(ws->canvas, canvas, elems * sizeof(float), cudaMemcpyHostToDevice,
                            ctx->stream));
  *dev = ws->canvas;
  return CERB_OK;
}

int finish(cerb_ctx* ctx, const int* dev_labels, int32_t* labels_out, size_t elems, int out_on_device) {
  CERB_CUDA(cudaMemcpyAsync(labels_out, dev_labels, elems * sizeof(int),
                            out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                            ctx->stream));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CERB_ERR_CUDA, "postproc launch: %s", cudaGetErrorString(e));
  if (!out_on_device) {
    int rc = cerb_ctx_sync(ctx);
    if (rc) return rc;
    const int code = ctx->err_flag_host[1];
    if (code != 0) {
      ctx->err_flag_host[1] = 0;
      return fail(CERB_ERR_KERNEL,
                  code == 10   ? "postproc: more instances than the size filter allows"
                  : code == 11 ? "postproc: an instance crop does not fit in shared memory "
                                 "(crop bits > on-chip capacity)"
                               : "postproc: kernel error %d",
                  code);
    }
  }
  return CERB_OK;
}

// cv2.getStructuringElement(MORPH_ELLIPSE, (k,k)): row i covers [c-dx, c+dx] with
// dx = round(c * sqrt((r^2 - dy^2) / r^2)), r = c = k/2, dy = i - r.
void ellipse_rows(int k, EllipseRows& e) {
  e.k = k;
  const int r = k / 2, c = k / 2;
  const double inv_r2 = r ? 1.0 / (static_cast<double>(r) * r) : 0.0;
  for (int i = 0; i < k; ++i) {
    int j1 = 0, j2 = 0;
    const int dy = i - r;
    if (std::abs(dy) <= r) {
      const int dx = static_cast<int>(std::lrint(c * std::sqrt((r * r - dy * dy) * inv_r2)));
      j1 = std::max(c - dx, 0);
      j2 = std::min(c + dx + 1, k);
    }
    e.j1[i] = j1;
    e.j2[i] = j2;
  }
}

}  // namespace

extern "C" int cerb_ellipse_rows(int k, int32_t* j1, int32_t* j2) {
  if (k < 1 || k > 32 || !j1 || !j2) return fail(CERB_ERR_ARG, "cerb_ellipse_rows: k must be in 1..32");
  EllipseRows e;
  ellipse_rows(k, e);
  for (int i = 0; i < k; ++i) {
    j1[i] = e.j1[i];
    j2[i] = e.j2[i];
  }
  return CERB_OK;
}

extern "C" int cerb_postproc_nuclei(cerb_ctx* ctx, const float* canvas, int n, int H, int W, int C,
                                    int ch0, int32_t* labels_out, int32_t* any_fg_out, int flags) {
  if (!ctx || !canvas || !labels_out || n <= 0 || H <= 0 || W <= 0 || C < 2 || ch0 < 0 ||
      ch0 + 2 > C)
    return fail(CERB_ERR_ARG, "cerb_postproc_nuclei: bad arguments");
  if (static_cast<size_t>(H) * W > static_cast<size_t>(INT_MAX) / 4)
    return fail(CERB_ERR_ARG, "cerb_postproc_nuclei: image too large");
  CERB_CUDA(cudaSetDevice(ctx->device));
  const int hw = H * W;
  Workspace* ws = nullptr;
  int rc = ensure_ws(ctx, ws, n, hw);
  if (rc) return rc;
  const float* dcanvas = nullptr;
  rc = stage_canvas(ctx, ws, canvas, static_cast<size_t>(n) * hw * C, flags & 1, &dcanvas);
  if (rc) return rc;
  cudaStream_t s = ctx->stream;
  const dim3 g = grid2(hw, n);
  CERB_CUDA(cudaMemsetAsync(ws->any_fg, 0, sizeof(int) * n, s));
  uint8_t* msk0 = ws->m0;
  uint8_t* mrk = ws->m1;
  uint8_t* msk = ws->m2;
  k_nuc_threshold<<<g, kThreads, 0, s>>>(dcanvas, C, ch0, msk0, mrk, ws->val, ws->any_fg, hw);
  k_erode_cross<<<g, kThreads, 0, s>>>(msk0, msk, H, W);
  ctx->launches += 2;
  cc_label(ctx, ws, mrk, n, H, W, 4);  // :370-373
  // :375-376 binary_fill_holes(marker)
  uint8_t* bg = msk0;  // msk0 is dead after the erosion
  k_invert<<<g, kThreads, 0, s>>>(mrk, bg, hw);
  ctx->launches += 1;
  cc_label(ctx, ws, bg, n, H, W, 0);
  k_mark_border<<<dim3((2 * (H + W) + kThreads - 1) / kThreads, n), kThreads, 0, s>>>(bg, ws->L, ws->size, H, W);
  k_fill_holes<<<g, kThreads, 0, s>>>(mrk, bg, ws->L, ws->size, hw);
  ctx->launches += 2;
  // :377 label -> raster-order ids
  cc_label(ctx, ws, mrk, n, H, W, 0);
  k_rank_roots<<<n, 1024, 0, s>>>(mrk, ws->L, ws->rank, ws->count, nullptr, hw);
  k_apply_rank<<<g, kThreads, 0, s>>>(mrk, ws->L, ws->rank, ws->lab, hw);
  // :366-368 (done last so that ws->L / ws->size describe the mask components for the watershed)
  cc_label(ctx, ws, msk, n, H, W, 8);
  k_mask_markers<<<g, kThreads, 0, s>>>(ws->lab, msk, hw);
  // :378 watershed(-inner, marker, mask)
  if (hw <= 65536) {
    const size_t smem = sizeof(uint64_t) * kWsHeapSmem + 2u * ((hw + 7) & ~7) + kWsHeapSmem + 16;
    static bool attr_set = false;
    if (!attr_set) {
      CERB_CUDA(cudaFuncSetAttribute(k_watershed_smem, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     226 * 1024));
      attr_set = true;
    }
    const size_t smem_c = sizeof(uint64_t) * kWcPool + 2u * ((hw + 7) & ~7) + 12u * kWcMaxComp + 16;
    static bool attr_c = false;
    if (!attr_c) {
      CERB_CUDA(cudaFuncSetAttribute(k_watershed_comp, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     226 * 1024));
      attr_c = true;
    }
    const bool fast = (ctx->ws_mode != 1);
    if (fast) {
      k_watershed_comp<<<n, kWcThreads, smem_c, s>>>(ws->val, msk, ws->L, ws->size, ws->lab,
                                                     ws->heap_a, ws->count, H, W, ctx->stat_dev,
                                                     ctx->ws_mode == 2 ? 1 : 0);
      ctx->launches += 1;
    }
    k_watershed_smem<<<n, kWsThreads, smem, s>>>(ws->val, msk, ws->lab,
                                                 reinterpret_cast<uint64_t*>(ws->heap_k), ws->m0,
                                                 ws->rank, ws->heap_v, fast ? ws->count : nullptr, H,
                                                 W);
  } else if (ctx->ws_mode == 1) {
    k_watershed<<<n, 32, 0, s>>>(ws->val, msk, ws->lab, ws->heap_v, ws->heap_a, ws->heap_i, H, W,
                                 nullptr);
  } else {
    // large images: one thread per mask component, heaps in a global pool (see k_wsg_flood).
    // ws->rank = slice offsets, ws->heap_a = per-root counters / root list is ws->heap_i's upper
    // neighbour ws->m1-free int array: roots go to ws->rank2 (= heap_v reinterpreted as int).
    if (static_cast<size_t>(n) * hw > ws->tie_cap) {
      size_t c = 0;
      CERB_CUDA(grow<int>(ctx, ws->tie_log, c, static_cast<size_t>(n) * hw));
      c = 0;
      CERB_CUDA(grow<int>(ctx, ws->tie_seeds, c, static_cast<size_t>(n) * hw));
      ws->tie_cap = static_cast<size_t>(n) * hw;
    }
    int* off = ws->rank;
    int* cnt = ws->heap_a;
    int* roots = reinterpret_cast<int*>(ws->heap_v);
    CERB_CUDA(cudaMemsetAsync(ws->ctl, 0, sizeof(int) * 4 * n, s));
    k_wsg_alloc<<<g, kThreads, 0, s>>>(msk, ws->L, ws->size, off, cnt, ws->aux, roots, ws->ctl, hw);
    // ws->size is dead now: it keeps the markers for the tie fallback
    CERB_CUDA(cudaMemcpyAsync(ws->size, ws->lab, sizeof(int) * static_cast<size_t>(n) * hw,
                              cudaMemcpyDeviceToDevice, s));
    k_wsg_seed<<<g, kThreads, 0, s>>>(ws->val, msk, ws->L, ws->lab, off, cnt, ws->aux, ws->heap_k,
                                      ws->heap_i, H, W, ws->ctl, 0);
    k_wsg_flood<<<dim3(148 * 8, n), 64, 0, s>>>(ws->val, msk, ws->lab, off, cnt, ws->aux, roots,
                                                ws->heap_k, ws->heap_i, ws->ctl, H, W);
    // components that saw a marker tie: all orders of the tied entries (each kernel returns at
    // once when the image has none)
    k_wsg_tied_reset<<<g, kThreads, 0, s>>>(msk, ws->L, ws->aux, ws->size, ws->lab, cnt, ws->ctl, hw);
    k_wsg_seed<<<g, kThreads, 0, s>>>(ws->val, msk, ws->L, ws->lab, off, cnt, ws->aux, ws->heap_k,
                                      ws->heap_i, H, W, ws->ctl, 1);
    k_wsg_certify<<<dim3(148 * 8, n), 64, 0, s>>>(ws->val, msk, ws->lab, off, cnt, ws->aux, roots,
                                                  ws->heap_k, ws->heap_i, ws->ctl, ws->size,
                                                  ws->tie_log, ws->tie_seeds, H, W);
    ctx->launches += 3;
    k_wsg_restore<<<g, kThreads, 0, s>>>(ws->lab, ws->size, ws->ctl, hw);
    k_wsg_flag<<<(n + 255) / 256, 256, 0, s>>>(ws->ctl, ws->count, n);
    k_watershed<<<n, 32, 0, s>>>(ws->val, msk, ws->lab, ws->heap_v, ws->heap_a, ws->heap_i, H, W,
                                 ws->count);
    ctx->launches += 6;
  }
  ctx->launches += 4;
  if (any_fg_out) {
    CERB_CUDA(cudaMemcpyAsync(any_fg_out, ws->any_fg, sizeof(int) * n,
                              (flags & 2) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  }
  const bool large = hw > 65536 && ctx->ws_mode != 1;
  rc = finish(ctx, ws->lab, labels_out, static_cast<size_t>(n) * hw, flags & 2);
  if (rc == CERB_OK && large) {
    // account for tied components / tiles that needed the exact whole-image fallback (the copy
    // synchronises; the large-image path is the WSI mode's, which waits for the labels anyway)
    std::vector<int> ctl(static_cast<size_t>(4) * n);
    CERB_CUDA(cudaMemcpy(ctl.data(), ws->ctl, sizeof(int) * 4 * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
      ctx->stat_ws_large += 1;
      ctx->stat_ws_fallback += ctl[4 * i + 2] != 0;
      ctx->stat_ws_tied += ctl[4 * i + 3];
    }
  }
  return rc;
}

extern "C" int64_t cerb_ctx_stat(cerb_ctx* ctx, const char* name) {
  if (!ctx || !name) return -1;
  if (strcmp(name, "ws_images") == 0 || strcmp(name, "ws_tie_fallbacks") == 0 ||
      strcmp(name, "ws_capacity_fallbacks") == 0) {
    // tiles <= 65536 px: images seen by the component-parallel watershed / redone by the exact
    // whole-tile emulation because of a marker tie / because the tile exceeded the heap pool
    if (ctx->stat_dev == nullptr) return 0;
    unsigned long long h[4] = {0, 0, 0, 0};
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (cudaMemcpy(h, ctx->stat_dev, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return static_cast<int64_t>(name[3] == 'i' ? h[0] : name[3] == 't' ? h[2] : h[3]);
  }
  if (strcmp(name, "ws_large_images") == 0) return ctx->stat_ws_large;
  if (strcmp(name, "ws_large_fallbacks") == 0) return ctx->stat_ws_fallback;
  if (strcmp(name, "ws_large_tied_components") == 0) return ctx->stat_ws_tied;
  return -1;
}

// Shared tail of the per-instance "dilate inside a padded crop, fill holes, paint" post-processing
// (loader/postproc.py:147-265 and :270-350): threshold -> remove_small_objects -> label -> one
// block per instance. `single`: the foreground is one channel > thr, else inner - (contour > 0.5).
static int gl_pipeline(cerb_ctx* ctx, const char* what, const float* canvas, int n, int H, int W,
                       int C, int ch0, int single, float thr, int k, int min_size,
                       int32_t* labels_out, int flags) {
  if (static_cast<size_t>(H) * W > static_cast<size_t>(INT_MAX) / 4)
    return fail(CERB_ERR_ARG, "%s: image too large", what);
  if (k < 1 || k > 32)
    return fail(CERB_ERR_ARG, "%s: structuring element size %d unsupported", what, k);
  CERB_CUDA(cudaSetDevice(ctx->device));
  const int hw = H * W;
  Workspace* ws = nullptr;
  int rc = ensure_ws(ctx, ws, n, hw);
  if (rc) return rc;
  const float* dcanvas = nullptr;
  rc = stage_canvas(ctx, ws, canvas, static_cast<size_t>(n) * hw * C, flags & 1, &dcanvas);
  if (rc) return rc;
  cudaStream_t s = ctx->stream;
  const dim3 g = grid2(hw, n);
  uint8_t* fg = ws->m0;
  k_gl_threshold<<<g, kThreads, 0, s>>>(dcanvas, C, ch0, thr, fg, hw, single);
  ctx->launches += 1;
  cc_label(ctx, ws, fg, n, H, W, min_size > 0 ? min_size : 0);
  // ws->any_fg doubles as the per-image "has a background pixel" flag here
  k_rank_roots<<<n, 1024, 0, s>>>(fg, ws->L, ws->rank, ws->count, ws->any_fg, hw);
  k_apply_rank<<<g, kThreads, 0, s>>>(fg, ws->L, ws->rank, ws->lab, hw);
  ctx->launches += 2;
  int max_inst = min_size > 1 ? hw / min_size + 1 : hw;
  if (max_inst > 65535) max_inst = 65535;
  const size_t bb_need = static_cast<size_t>(n) * max_inst * 4;
  CERB_CUDA(grow<int>(ctx, ws->bb, ws->bb_elems, bb_need));
  k_bbox_init<<<148, kThreads, 0, s>>>(ws->bb, static_cast<int>(bb_need));
  k_bbox<<<g, kThreads, 0, s>>>(ws->lab, ws->bb, H, W, max_inst, ctx->err_flag_dev + 1);
  // output map (ws->size is free now)
  int* out = ws->size;
  CERB_CUDA(cudaMemsetAsync(out, 0, sizeof(int) * static_cast<size_t>(n) * hw, s));
  EllipseRows ell;
  ellipse_rows(k, ell);
  const int full_words = 2 * H * ((W + 31) / 32);
  int smem_words = full_words < 50 * 1024 ? full_words : 50 * 1024;  // <= 200 KB
  static bool attr_set = false;
  if (!attr_set) {
    CERB_CUDA(cudaFuncSetAttribute(k_gl_instance, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   200 * 1024));
    attr_set = true;
  }
  // overflow pool for oversized crops: the connected-component parent plane (n * hw words, free by
  // now) holds 16 n whole-image crops; its fill pointer is one of the watershed control words
  unsigned int* pool_top = reinterpret_cast<unsigned int*>(ws->ctl);
  CERB_CUDA(cudaMemsetAsync(pool_top, 0, sizeof(unsigned int), s));
  const size_t pool_words = static_cast<size_t>(n) * hw;
  k_gl_instance<<<dim3(max_inst, n), kThreads, static_cast<size_t>(smem_words) * 4, s>>>(
      ws->lab, ws->count, ws->any_fg, ws->bb, out, H, W, max_inst, ell, smem_words,
      reinterpret_cast<uint32_t*>(ws->L),
      static_cast<unsigned int>(pool_words > 0xffffffffull ? 0xffffffffull : pool_words), pool_top,
      ctx->err_flag_dev + 1);
  ctx->launches += 3;
  return finish(ctx, out, labels_out, static_cast<size_t>(n) * hw, flags & 2);
}

extern "C" int cerb_postproc_gland_lumen(cerb_ctx* ctx, const float* canvas, int n, int H, int W,
                                         int C, int ch0, int tissue, double ds_factor,
                                         int32_t* labels_out, int flags) {
  if (!ctx || !canvas || !labels_out || n <= 0 || H <= 0 || W <= 0 || C < 2 || ch0 < 0 ||
      ch0 + 2 > C || (tissue != 0 && tissue != 1) || !(ds_factor > 0.0))
    return fail(CERB_ERR_ARG, "cerb_postproc_gland_lumen: bad arguments");
  // loader/postproc.py:272-275,287 (gland) / :314-317,328 (lumen)
  const int ksize_ = tissue == 0 ? 11 : 3;
  const int k = static_cast<int>((ksize_ - 1) * ds_factor);
  const int min_size = static_cast<int>((tissue == 0 ? 1000 : 150) * (ds_factor * ds_factor));
  const float thr = tissue == 0 ? 0.55f : 0.5f;
  return gl_pipeline(ctx, "cerb_postproc_gland_lumen", canvas, n, H, W, C, ch0, 0, thr, k, min_size,
                     labels_out, flags);
}

// PostProcInstErodedMap (loader/postproc.py:147-265): tissue 0 gland (ellipse 11,
// min size 1500), 1 lumen (3, 150), 2 nuclei (3, 8); foreground = channel ch0 > 0.5.
extern "C" int cerb_postproc_eroded_map(cerb_ctx* ctx, const float* canvas, int n, int H, int W,
                                        int C, int ch0, int tissue, int32_t* labels_out,
                                        int flags) {
  if (!ctx || !canvas || !labels_out || n <= 0 || H <= 0 || W <= 0 || C < 1 || ch0 < 0 ||
      ch0 + 1 > C || tissue < 0 || tissue > 2)
    return fail(CERB_ERR_ARG, "cerb_postproc_eroded_map: bad arguments");
  const int k = tissue == 0 ? 11 : 3;
  const int min_size = tissue == 0 ? 1500 : tissue == 1 ? 150 : 8;
  return gl_pipeline(ctx, "cerb_postproc_eroded_map", canvas, n, H, W, C, ch0, 1, 0.5f, k, min_size,
                     labels_out, flags);
}

// infer/tile.py:187-191: lumen *= (gland > 0)
extern "C" int cerb_mask_lumen(cerb_ctx* ctx, int32_t* lumen_dev, const int32_t* gland_dev,
                               size_t elems) {
  if (!ctx || !lumen_dev || !gland_dev) return fail(CERB_ERR_ARG, "cerb_mask_lumen: bad arguments");
  CERB_CUDA(cudaSetDevice(ctx->device));
  k_mask_by<<<148 * 4, kThreads, 0, ctx->stream>>>(lumen_dev, gland_dev, elems);
  ctx->launches += 1;
  CERB_CUDA(cudaGetLastError());
  return CERB_OK;
}
