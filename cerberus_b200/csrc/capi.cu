// C ABI of the forward pass: context, plan (tensors + ops), execution.
// Declared in include/cerberus_b200.h; bound from Python with ctypes (cerberus_b200/_lib.py).
//
// Replaces the reference's run_step closure (infer/base.py:51-53 -> models/run_desc.py:439-502
// -> models/net_desc.py:144-200). The plan is a flat list of ops over NHWC tensors; Python
// builds it once per batch shape from the model directory (cerberus_b200/plan.py).
#include "../../include/cerberus_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "capi_internal.cuh"
#include "conv3x3.cuh"
#include "conv3x3c2.cuh"
#include "conv64.cuh"
#include "conv64s.cuh"
#include "conv64x.cuh"
#include "conv_tc.cuh"
#include "ops.cuh"

namespace cerb {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

}  // namespace cerb

using namespace cerb;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Tensor {
  cerb_tensor_desc d;
  void* plane[2] = {nullptr, nullptr};
  size_t bytes = 0;
};

struct Step {
  int kind = 0;
  ConvKParams conv;
  Conv64Params c64;
  Conv3Params c3;
  Conv3c2Params c3p;
  Conv64sParams c64s;
  bool use64s = false;  // split-precision 64->64 kernel (csrc/conv64s.cu)
  bool use3p = false;  // CTA-pair kernel (csrc/conv3x3c2.cu)
  int chain_len = 0;     // > 1: this launch also runs the next chain_len - 1 steps (Conv3c2Params::layers)
  bool chained = false;  // executed by the launch of an earlier step of its chain
  Conv64xParams c64x;
  bool use64x = false;
  bool use64 = false;
  bool use3 = false;
  bool side = false;  // may overlap the ops that follow (nothing later in the plan reads its output)
  bool split = false;
  // misc ops
  ActRef a, b, c;
  // grouped UPADD (op.cout = G > 1): prev / out of every group; a = the shared skip
  int up_groups = 0;
  ActRef up_prev[8], up_out[8];
  const uint8_t* u8_in = nullptr;
  HeadParams head;
  PClassParams pclass;
};

size_t dtype_size(int dt) {
  switch (dt) {
    case CERB_U8: return 1;
    case CERB_F16: return 2;
    case CERB_F32: return 4;
    case CERB_I32: return 4;
  }
  return 0;
}

}  // namespace

struct cerb_plan {
  cerb_ctx* ctx = nullptr;
  std::vector<Tensor> tensors;
  std::vector<Step> steps;
  uint8_t* blob = nullptr;
  size_t blob_bytes = 0;
  bool owns_blob = true;  // false: the blob belongs to a cerb_model shared by several plans
  int prep_in_tensor = -1;
  // the op list is launch-bound on the host (~100 small launches): it is captured into a CUDA
  // graph on the second run (the first run doubles as warm-up / attribute setup) and replayed
  cudaGraphExec_t graph_exec = nullptr;
  int runs = 0;
  int* cur_counter = nullptr;    // counter of the op being built
  int* tile_counters = nullptr;  // one per op (dynamic tile scheduling), zeroed at the start of a run
  size_t n_counters = 0;         // ... followed by the per-image completion counters of layer chains
  std::vector<void*> chain_tables;  // device layer tables of the chains
  int n_launches = 0;            // kernel launches of one run (chained steps do not launch)
};

namespace {

ActRef act_ref(const Tensor& t) {
  ActRef a;
  a.hi = static_cast<__half*>(t.plane[0]);
  a.lo = static_cast<__half*>(t.plane[1]);
  a.n = t.d.n;
  a.h = t.d.h;
  a.w = t.d.w;
  a.c = t.d.c;
  return a;
}

int encode_map(cerb_ctx* ctx, CUtensorMap* m, void* base, int rank, const cuuint64_t* dims,
               const cuuint64_t* strides_bytes, const cuuint32_t* box, int x_stride = 1) {
  cuuint32_t estr[5] = {1, static_cast<cuuint32_t>(x_stride), 1, 1, 1};
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), base, dims,
                  strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(CERB_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] strides "
                "[%llu,%llu,%llu] box [%u,%u,%u,%u]",
                static_cast<int>(r), rank, (unsigned long long)dims[0],
                (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                (unsigned long long)(rank > 3 ? dims[3] : 0), (unsigned long long)strides_bytes[0],
                (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
                (unsigned long long)(rank > 3 ? strides_bytes[2] : 0), box[0], box[1],
                rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  }
  return CERB_OK;
}

int floor_log2(int v) {
  int l = 0;
  while ((1 << (l + 1)) <= v) ++l;
  return l;
}

// Picks the 128-pixel box (BW x BH, both powers of two) that wastes the fewest MMA rows.
int pick_box_w(int H, int W) {
  long best_area = -1;
  int best = 128;
  for (int bw = 128; bw >= 1; bw >>= 1) {
    const int bh = 128 / bw;
    const long area = static_cast<long>((W + bw - 1) / bw) * bw * ((H + bh - 1) / bh) * bh;
    if (best_area < 0 || area < best_area) {
      best_area = area;
      best = bw;
    }
  }
  return best;
}

int pick_bn(int cout, long m_tiles, int num_sms) {
  if (cout <= 128) return cout;
  // Larger BN re-reads the activation slab fewer times; keep at least ~2 tiles per SM.
  const int cands[2] = {256, 128};
  for (int c : cands) {
    if (cout % c != 0) continue;
    if (m_tiles * (cout / c) >= 2L * num_sms) return c;
  }
  if (cout % 128 == 0) return 128;
  if (cout % 64 == 0) return 64;
  return 0;
}

int build_conv64(cerb_plan* pl, const cerb_op& op, Step& st) {
  cerb_ctx* ctx = pl->ctx;
  const Tensor& in = pl->tensors[op.in0];
  const Tensor& out = pl->tensors[op.out];
  Conv64Params& p = st.c64;
  memset(&p, 0, sizeof(p));
  st.use64 = true;
  const bool tail = op.aux_classes > 0;  // fused classification head: `out` is the fp32 canvas
  const int H = tail ? in.d.h : out.d.h, W = tail ? in.d.w : out.d.w, N = out.d.n;
  p.mode = ctx->conv64_mode == 3 ? 1 : ctx->conv64_mode;  // mode 3 routes plain convs to conv64x.cu
  p.n_taps = 9;
  p.halo_x = 1;
  p.halo_y = 1;
  p.debug = ctx->conv64_debug;
  p.n_img = N;
  p.H = H;
  p.W = W;
  p.tiles_x = (W + conv64_tile_w() - 1) / conv64_tile_w();
  p.tiles_y = (H + conv64_tile_h() - 1) / conv64_tile_h();
  p.n_tiles = N * p.tiles_x * p.tiles_y;
  const size_t es = 2;
  if (op.in_coff % 8 != 0 || op.in_coff + 64 > in.d.c || in.d.c % 8 != 0)
    return fail(CERB_ERR_ARG, "conv64: bad input channels");
  {
    const cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                                static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(in.d.c) * es,
                                   static_cast<cuuint64_t>(W) * in.d.c * es,
                                   static_cast<cuuint64_t>(H) * W * in.d.c * es};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(conv64_box_w(p.mode)),
                               static_cast<cuuint32_t>(conv64_tile_h() + 2), 1};
    int rc = encode_map(ctx, &p.in_map, static_cast<__half*>(in.plane[0]) + op.in_coff, 4, dims,
                        strides, box);
    if (rc) return rc;
  }
  if (op.w_off < 0 || op.w_off % 16 != 0 ||
      static_cast<size_t>(op.w_off) + 64u * 576u * es > pl->blob_bytes)
    return fail(CERB_ERR_ARG, "conv64: weight offset out of range");
  {
    const cuuint64_t dims[2] = {576, 64};
    const cuuint64_t strides[1] = {576 * es};
    const cuuint32_t box[2] = {64, 64};
    int rc = encode_map(ctx, &p.w_map, pl->blob + op.w_off, 2, dims, strides, box);
    if (rc) return rc;
  }
  if (op.b_off >= 0) {
    if (op.b_off % 16 != 0 || static_cast<size_t>(op.b_off) + 64 * 4u > pl->blob_bytes)
      return fail(CERB_ERR_ARG, "conv64: bias offset out of range");
    p.bias = reinterpret_cast<const float*>(pl->blob + op.b_off);
  }
  if (tail) {
    const int width = op.head_mode == CERB_HEAD_INST ? op.aux_classes - 1 : 1;
    const int nt = static_cast<int>(pl->tensors.size());
    if (ctx->conv64_mode != 1 || op.in1 >= 0 || op.up_prev1 > 0 || op.aux_classes < 2 ||
        op.aux_classes > 8 || out.d.dtype != CERB_F32 || out.d.h > H || out.d.w > W ||
        op.out_coff < 0 || op.out_coff + width > out.d.c || op.tail_w_off < 0 ||
        op.tail_w_off % 16 != 0 || op.tail_b_off < 0 || op.tail_b_off % 16 != 0 ||
        op.aux_w_off < 0 || op.aux_b_off < 0 ||
        static_cast<size_t>(op.tail_w_off) + 96u * 64u * es > pl->blob_bytes ||
        static_cast<size_t>(op.tail_b_off) + 96u * 4u > pl->blob_bytes ||
        static_cast<size_t>(op.aux_w_off) + op.aux_classes * 96 * 4u > pl->blob_bytes ||
        static_cast<size_t>(op.aux_b_off) + op.aux_classes * 4u > pl->blob_bytes ||
        op.tail_w_shift < -60 || op.tail_w_shift > 60)
      return fail(CERB_ERR_ARG, "conv64: bad fused-head description");
    const cuuint64_t dims[2] = {64, 96};
    const cuuint64_t strides[1] = {64 * es};
    const cuuint32_t box[2] = {64, 96};
    int rc = encode_map(ctx, &p.w1_map, pl->blob + op.tail_w_off, 2, dims, strides, box);
    if (rc) return rc;
    p.has_tail = 1;
    p.tail_b1 = reinterpret_cast<const float*>(pl->blob + op.tail_b_off);
    p.tail_scale = ldexpf(1.0f, -op.tail_w_shift);
    p.tail_w2 = reinterpret_cast<const float*>(pl->blob + op.aux_w_off);
    p.tail_b2 = reinterpret_cast<const float*>(pl->blob + op.aux_b_off);
    p.tail_classes = op.aux_classes;
    p.tail_mode = op.head_mode;
    p.canvas = static_cast<float*>(out.plane[0]);
    p.oh = out.d.h;
    p.ow = out.d.w;
    p.canvas_c = out.d.c;
    p.canvas_coff = op.out_coff;
    if (op.logits_out >= 0) {
      if (op.logits_out >= nt) return fail(CERB_ERR_ARG, "conv64: logits id out of range");
      const Tensor& lg = pl->tensors[op.logits_out];
      if (lg.d.dtype != CERB_F32 || lg.d.n != N || lg.d.h != H || lg.d.w != W ||
          lg.d.c != op.aux_classes)
        return fail(CERB_ERR_ARG, "conv64: fused-head logits tensor mismatch");
      p.logits = static_cast<float*>(lg.plane[0]);
    }
  } else {
    if (op.out_coff % 8 != 0 || op.out_coff + 64 > out.d.c || out.d.c % 8 != 0)
      return fail(CERB_ERR_ARG, "conv64: bad output channels");
    const cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                                static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(out.d.c) * es,
                                   static_cast<cuuint64_t>(W) * out.d.c * es,
                                   static_cast<cuuint64_t>(H) * W * out.d.c * es};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(conv64_tile_w()),
                               static_cast<cuuint32_t>(conv64_tile_h()), 1};
    int rc = encode_map(ctx, &p.out_map, static_cast<__half*>(out.plane[0]) + op.out_coff, 4, dims,
                        strides, box);
    if (rc) return rc;
  }
  if (op.in1 >= 0) {
    if (op.in1 >= static_cast<int>(pl->tensors.size()))
      return fail(CERB_ERR_ARG, "conv64: residual id out of range");
    const Tensor& res = pl->tensors[op.in1];
    if (res.d.n != N || res.d.h != H || res.d.w != W || res.d.c < 64 || res.d.c % 8 != 0 ||
        res.d.dtype != CERB_F16)
      return fail(CERB_ERR_ARG, "conv64: residual shape mismatch");
    const cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                                static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(res.d.c) * es,
                                   static_cast<cuuint64_t>(W) * res.d.c * es,
                                   static_cast<cuuint64_t>(H) * W * res.d.c * es};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(conv64_tile_w()),
                               static_cast<cuuint32_t>(conv64_tile_h()), 1};
    int rc = encode_map(ctx, &p.res_map, res.plane[0], 4, dims, strides, box);
    if (rc) return rc;
    p.has_res = 1;
  }
  if (op.up_prev1 > 0) {
    const int pid = op.up_prev1 - 1;
    if (pid >= static_cast<int>(pl->tensors.size()))
      return fail(CERB_ERR_ARG, "conv64: up_prev id out of range");
    const Tensor& pv = pl->tensors[pid];
    if (pv.d.dtype != CERB_F16 || pv.d.n != N || pv.d.h * 2 != H || pv.d.w * 2 != W || pv.d.c < 64 ||
        pv.d.c % 8 != 0 || op.in_coff != 0)
      return fail(CERB_ERR_ARG, "conv64: fused upsample+add shape mismatch");
    p.up_skip = static_cast<const __half*>(in.plane[0]);
    p.up_prev = static_cast<const __half*>(pv.plane[0]);
    p.up_skip_cs = in.d.c;
    p.up_prev_cs = pv.d.c;
  }
  conv64_plan(p);  // after up_prev is known: the fused producer needs staging shared memory
  p.relu = op.relu;
  if (op.w_shift < -60 || op.w_shift > 60) return fail(CERB_ERR_ARG, "conv64: w_shift out of range");
  p.acc_scale = ldexpf(1.0f, -op.w_shift);
  p.err_flag = ctx->err_flag_dev;
  p.prof = ctx->prof_dev;
  return CERB_OK;
}

// 7x7 stem on the resident-weight halo kernel (conv64.cu mode 4)
int build_stem64(cerb_plan* pl, const cerb_op& op, Step& st) {
  cerb_ctx* ctx = pl->ctx;
  const Tensor& in = pl->tensors[op.in0];
  const Tensor& out = pl->tensors[op.out];
  Conv64Params& p = st.c64;
  memset(&p, 0, sizeof(p));
  st.use64 = true;
  const int H = out.d.h, W = out.d.w, N = out.d.n;
  const size_t es = 2;
  if (in.d.c != 8 || in.d.h != H || in.d.w != W + 8 || op.kh != 7 || op.kw != 7 || op.stride != 1 ||
      op.pad != 3 || op.cout != 64 || op.in1 >= 0)
    return fail(CERB_ERR_ARG, "conv: stem expects a 7x7 s1 p3 3->64 conv over a PREP tensor");
  p.mode = 4;
  p.n_taps = 7;
  p.halo_x = 0;
  p.halo_y = 3;
  p.n_img = N;
  p.H = H;
  p.W = W;
  p.tiles_x = (W + conv64_tile_w() - 1) / conv64_tile_w();
  p.tiles_y = (H + conv64_tile_h() - 1) / conv64_tile_h();
  p.n_tiles = N * p.tiles_x * p.tiles_y;
  {
    // overlapping windows: "pixel" x of the view = padded columns x .. x+7 (8 px x 8 ch = 64 K)
    const cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                                static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {8 * es, static_cast<cuuint64_t>(W + 8) * 8 * es,
                                   static_cast<cuuint64_t>(H) * (W + 8) * 8 * es};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(conv64_tile_w()),
                               static_cast<cuuint32_t>(conv64_tile_h() + 6), 1};
    int rc = encode_map(ctx, &p.in_map, in.plane[0], 4, dims, strides, box);
    if (rc) return rc;
  }
  if (op.w_off < 0 || op.w_off % 16 != 0 ||
      static_cast<size_t>(op.w_off) + 64u * 448u * es > pl->blob_bytes)
    return fail(CERB_ERR_ARG, "conv: stem weight offset out of range");
  {
    const cuuint64_t dims[2] = {448, 64};
    const cuuint64_t strides[1] = {448 * es};
    const cuuint32_t box[2] = {64, 64};
    int rc = encode_map(ctx, &p.w_map, pl->blob + op.w_off, 2, dims, strides, box);
    if (rc) return rc;
  }
  if (op.b_off >= 0) {
    if (op.b_off % 16 != 0 || static_cast<size_t>(op.b_off) + 64 * 4u > pl->blob_bytes)
      return fail(CERB_ERR_ARG, "conv: stem bias offset out of range");
    p.bias = reinterpret_cast<const float*>(pl->blob + op.b_off);
  }
  if (op.out_coff % 8 != 0 || op.out_coff + 64 > out.d.c || out.d.c % 8 != 0)
    return fail(CERB_ERR_ARG, "conv: stem output channels");
  {
    const cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                                static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(out.d.c) * es,
                                   static_cast<cuuint64_t>(W) * out.d.c * es,
                                   static_cast<cuuint64_t>(H) * W * out.d.c * es};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(conv64_tile_w()),
                               static_cast<cuuint32_t>(conv64_tile_h()), 1};
    int rc = encode_map(ctx, &p.out_map, static_cast<__half*>(out.plane[0]) + op.out_coff, 4, dims,
                        strides, box);
    if (rc) return rc;
  }
  conv64_plan(p);
  p.relu = op.relu;
  if (op.w_shift < -60 || op.w_shift > 60) return fail(CERB_ERR_ARG, "conv: w_shift out of range");
  p.acc_scale = ldexpf(1.0f, -op.w_shift);
  p.err_flag = ctx->err_flag_dev;
  p.prof = ctx->prof_dev;
  return CERB_OK;
}

int build_conv64x(cerb_plan* pl, const cerb_op& op, Step& st) {
  cerb_ctx* ctx = pl->ctx;
  const Tensor& in = pl->tensors[op.in0];
  const Tensor& out = pl->tensors[op.out];
  Conv64xParams& p = st.c64x;
  memset(&p, 0, sizeof(p));
  st.use64x = true;
  const int H = out.d.h, W = out.d.w, N = out.d.n;
  const size_t es = 2;
  p.n_img = N;
  p.H = H;
  p.W = W;
  if (op.in_coff % 8 != 0 || op.in_coff + 64 > in.d.c || in.d.c % 8 != 0)
    return fail(CERB_ERR_ARG, "conv64x: bad input channels");
  if (op.out_coff % 8 != 0 || op.out_coff + 64 > out.d.c || out.d.c % 8 != 0)
    return fail(CERB_ERR_ARG, "conv64x: bad output channels");
  const cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                              static_cast<cuuint64_t>(N)};
  {
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(in.d.c) * es,
                                   static_cast<cuuint64_t>(W) * in.d.c * es,
                                   static_cast<cuuint64_t>(H) * W * in.d.c * es};
    const cuuint32_t box[4] = {64, 18, 18, 1};  // element stride 2 in x: 9 columns of one parity
    int rc = encode_map(ctx, &p.in_map, static_cast<__half*>(in.plane[0]) + op.in_coff, 4, dims,
                        strides, box, 2);
    if (rc) return rc;
  }
  if (op.w_off < 0 || op.w_off % 16 != 0 ||
      static_cast<size_t>(op.w_off) + 64u * 576u * es > pl->blob_bytes)
    return fail(CERB_ERR_ARG, "conv64x: weight offset out of range");
  {
    const cuuint64_t wd[2] = {576, 64};
    const cuuint64_t ws[1] = {576 * es};
    const cuuint32_t wb[2] = {64, 64};
    int rc = encode_map(ctx, &p.w_map, pl->blob + op.w_off, 2, wd, ws, wb);
    if (rc) return rc;
  }
  if (op.b_off >= 0) {
    if (op.b_off % 16 != 0 || static_cast<size_t>(op.b_off) + 64 * 4u > pl->blob_bytes)
      return fail(CERB_ERR_ARG, "conv64x: bias offset out of range");
    p.bias = reinterpret_cast<const float*>(pl->blob + op.b_off);
  }
  {
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(out.d.c) * es,
                                   static_cast<cuuint64_t>(W) * out.d.c * es,
                                   static_cast<cuuint64_t>(H) * W * out.d.c * es};
    const cuuint32_t box[4] = {64, 16, 16, 1};  // element stride 2 in x: the 8 pixels of one parity
    int rc = encode_map(ctx, &p.out_map, static_cast<__half*>(out.plane[0]) + op.out_coff, 4, dims,
                        strides, box, 2);
    if (rc) return rc;
  }
  if (op.in1 >= 0) {
    if (op.in1 >= static_cast<int>(pl->tensors.size()))
      return fail(CERB_ERR_ARG, "conv64x: residual id out of range");
    const Tensor& res = pl->tensors[op.in1];
    if (res.d.n != N || res.d.h != H || res.d.w != W || res.d.c < 64 || res.d.c % 8 != 0 ||
        res.d.dtype != CERB_F16)
      return fail(CERB_ERR_ARG, "conv64x: residual shape mismatch");
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(res.d.c) * es,
                                   static_cast<cuuint64_t>(W) * res.d.c * es,
                                   static_cast<cuuint64_t>(H) * W * res.d.c * es};
    const cuuint32_t box[4] = {64, 16, 16, 1};
    int rc = encode_map(ctx, &p.res_map, res.plane[0], 4, dims, strides, box, 2);
    if (rc) return rc;
    p.has_res = 1;
  }
  if (op.up_prev1 > 0) {
    // input = skip (in0) + bilinear_x2(low): the halo fix-up of conv64x.cu
    const int pid = op.up_prev1 - 1;
    if (pid >= static_cast<int>(pl->tensors.size()))
      return fail(CERB_ERR_ARG, "conv64x: up_prev id out of range");
    const Tensor& pv = pl->tensors[pid];
    if (pv.d.dtype != CERB_F16 || pv.d.n != N || pv.d.h * 2 != H || pv.d.w * 2 != W || pv.d.c < 64 ||
        pv.d.c % 8 != 0)
      return fail(CERB_ERR_ARG, "conv64x: fused upsample+add shape mismatch");
    const cuuint64_t ldims[4] = {64, static_cast<cuuint64_t>(pv.d.w), static_cast<cuuint64_t>(pv.d.h),
                                 static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(pv.d.c) * es,
                                   static_cast<cuuint64_t>(pv.d.w) * pv.d.c * es,
                                   static_cast<cuuint64_t>(pv.d.h) * pv.d.w * pv.d.c * es};
    const cuuint32_t box[4] = {64, 10, 10, 1};
    int rc = encode_map(ctx, &p.low_map, pv.plane[0], 4, ldims, strides, box);
    if (rc) return rc;
    p.fuse_up = 1;
  }
  p.relu = op.relu;
  if (op.w_shift < -60 || op.w_shift > 60) return fail(CERB_ERR_ARG, "conv64x: w_shift out of range");
  p.acc_scale = ldexpf(1.0f, -op.w_shift);
  p.err_flag = ctx->err_flag_dev;
  p.prof = ctx->prof_dev;
  p.tile_counter = ctx->dyn_sched ? pl->cur_counter : nullptr;
  conv64x_plan(p);
  return CERB_OK;
}

int build_conv3(cerb_plan* pl, const cerb_op& op, Step& st) {
  cerb_ctx* ctx = pl->ctx;
  const Tensor& in = pl->tensors[op.in0];
  const Tensor& out = pl->tensors[op.out];
  Conv3Params& p = st.c3;
  memset(&p, 0, sizeof(p));
  st.use3 = true;
  const int H = out.d.h, W = out.d.w, N = out.d.n;
  const size_t es = 2;
  p.n_img = N;
  p.H = H;
  p.W = W;
  p.n_chunks = op.in_c / 64;
  p.BN = op.cout % 128 == 0 ? 128 : 64;
  p.n_ntiles = op.cout / p.BN;
  if (op.in_coff % 8 != 0 || op.in_coff + op.in_c > in.d.c || in.d.c % 8 != 0)
    return fail(CERB_ERR_ARG, "conv3x3: bad input channels");
  if (op.out_coff % 8 != 0 || op.out_coff + op.cout > out.d.c || out.d.c % 8 != 0)
    return fail(CERB_ERR_ARG, "conv3x3: bad output channels");
  {
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(op.in_c), static_cast<cuuint64_t>(W),
                                static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(in.d.c) * es,
                                   static_cast<cuuint64_t>(W) * in.d.c * es,
                                   static_cast<cuuint64_t>(H) * W * in.d.c * es};
    const cuuint32_t box[4] = {64, 18, 18, 1};
    int rc = encode_map(ctx, &p.l0.in_map, static_cast<__half*>(in.plane[0]) + op.in_coff, 4, dims,
                        strides, box);
    if (rc) return rc;
  }
  const size_t k_total = static_cast<size_t>(9) * op.in_c;
  if (op.w_off < 0 || op.w_off % 16 != 0 ||
      static_cast<size_t>(op.w_off) + static_cast<size_t>(op.cout) * k_total * es > pl->blob_bytes)
    return fail(CERB_ERR_ARG, "conv3x3: weight offset out of range");
  {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(k_total), static_cast<cuuint64_t>(op.cout)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(k_total) * es};
    const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(p.BN)};
    int rc = encode_map(ctx, &p.l0.w_map, pl->blob + op.w_off, 2, dims, strides, box);
    if (rc) return rc;
  }
  if (op.b_off >= 0) {
    if (op.b_off % 16 != 0 || static_cast<size_t>(op.b_off) + op.cout * 4u > pl->blob_bytes)
      return fail(CERB_ERR_ARG, "conv3x3: bias offset out of range");
    p.l0.bias = reinterpret_cast<const float*>(pl->blob + op.b_off);
  }
  {
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(op.cout), static_cast<cuuint64_t>(W),
                                static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(out.d.c) * es,
                                   static_cast<cuuint64_t>(W) * out.d.c * es,
                                   static_cast<cuuint64_t>(H) * W * out.d.c * es};
    const cuuint32_t box[4] = {64, 8, 16, 1};
    int rc = encode_map(ctx, &p.l0.out_map, static_cast<__half*>(out.plane[0]) + op.out_coff, 4, dims,
                        strides, box);
    if (rc) return rc;
  }
  if (op.in1 >= 0) {
    if (op.in1 >= static_cast<int>(pl->tensors.size()))
      return fail(CERB_ERR_ARG, "conv3x3: residual id out of range");
    const Tensor& res = pl->tensors[op.in1];
    if (res.d.n != N || res.d.h != H || res.d.w != W || res.d.c < op.cout || res.d.c % 8 != 0 ||
        res.d.dtype != CERB_F16)
      return fail(CERB_ERR_ARG, "conv3x3: residual shape mismatch");
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(op.cout), static_cast<cuuint64_t>(W),
                                static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(res.d.c) * es,
                                   static_cast<cuuint64_t>(W) * res.d.c * es,
                                   static_cast<cuuint64_t>(H) * W * res.d.c * es};
    const cuuint32_t box[4] = {64, 8, 16, 1};
    int rc = encode_map(ctx, &p.l0.res_map, res.plane[0], 4, dims, strides, box);
    if (rc) return rc;
    p.l0.has_res = 1;
  }
  p.l0.relu = op.relu;
  if (op.w_shift < -60 || op.w_shift > 60) return fail(CERB_ERR_ARG, "conv3x3: w_shift out of range");
  p.l0.acc_scale = ldexpf(1.0f, -op.w_shift);
  p.err_flag = ctx->err_flag_dev;
  p.prof = ctx->prof_dev;
  p.tile_counter = ctx->dyn_sched ? pl->cur_counter : nullptr;
  // a K walk rotated per CTA would make the fp32 accumulation order of a work item depend on
  // which CTA drew it: only with the static split (it made no measurable difference anyway)
  p.rotate = ctx->k_rotate && p.tile_counter == nullptr;
  conv3x3_plan(p);
  return CERB_OK;
}

// 64->64 3x3 stride-1 in the split-precision mode: halo reuse + resident hi weights (csrc/conv64s.cu).
int build_conv64s(cerb_plan* pl, const cerb_op& op, Step& st) {
  cerb_ctx* ctx = pl->ctx;
  const Tensor& in = pl->tensors[op.in0];
  const Tensor& out = pl->tensors[op.out];
  Conv64sParams& p = st.c64s;
  memset(&p, 0, sizeof(p));
  st.use64s = true;
  const int H = out.d.h, W = out.d.w, N = out.d.n;
  const size_t es = 2;
  p.n_img = N;
  p.H = H;
  p.W = W;
  if (!in.plane[1] || !out.plane[1]) return fail(CERB_ERR_ARG, "conv64s: tensors lack a lo plane");
  if (op.in_coff % 8 != 0 || op.in_coff + 64 > in.d.c || in.d.c % 8 != 0)
    return fail(CERB_ERR_ARG, "conv64s: bad input channels");
  if (op.out_coff % 8 != 0 || op.out_coff + 64 > out.d.c || out.d.c % 8 != 0)
    return fail(CERB_ERR_ARG, "conv64s: bad output channels");
  const cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                              static_cast<cuuint64_t>(N)};
  auto act_map = [&](CUtensorMap* m, const Tensor& t, int plane, int coff, const cuuint32_t* box) {
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(t.d.c) * es,
                                   static_cast<cuuint64_t>(W) * t.d.c * es,
                                   static_cast<cuuint64_t>(H) * W * t.d.c * es};
    return encode_map(ctx, m, static_cast<__half*>(t.plane[plane]) + coff, 4, dims, strides, box);
  };
  const cuuint32_t halo_box[4] = {64, 10, 18, 1};
  const cuuint32_t tile_box[4] = {64, 8, 16, 1};
  int rc;
  if ((rc = act_map(&p.in_hi, in, 0, op.in_coff, halo_box))) return rc;
  if ((rc = act_map(&p.in_lo, in, 1, op.in_coff, halo_box))) return rc;
  if ((rc = act_map(&p.out_hi, out, 0, op.out_coff, tile_box))) return rc;
  if ((rc = act_map(&p.out_lo, out, 1, op.out_coff, tile_box))) return rc;
  const size_t wbytes = 64u * 576u * es;
  if (op.w_off < 0 || op.w_off % 16 != 0 || static_cast<size_t>(op.w_off) + wbytes > pl->blob_bytes ||
      op.w_lo_off < 0 || op.w_lo_off % 16 != 0 || static_cast<size_t>(op.w_lo_off) + wbytes > pl->blob_bytes)
    return fail(CERB_ERR_ARG, "conv64s: weight offsets out of range");
  {
    const cuuint64_t wd[2] = {576, 64};
    const cuuint64_t ws[1] = {576 * es};
    const cuuint32_t wb[2] = {64, 64};
    if ((rc = encode_map(ctx, &p.w_hi, pl->blob + op.w_off, 2, wd, ws, wb))) return rc;
    if ((rc = encode_map(ctx, &p.w_lo, pl->blob + op.w_lo_off, 2, wd, ws, wb))) return rc;
  }
  if (op.b_off >= 0) {
    if (op.b_off % 16 != 0 || static_cast<size_t>(op.b_off) + 64 * 4u > pl->blob_bytes)
      return fail(CERB_ERR_ARG, "conv64s: bias offset out of range");
    p.bias = reinterpret_cast<const float*>(pl->blob + op.b_off);
  }
  if (op.in1 >= 0) {
    if (op.in1 >= static_cast<int>(pl->tensors.size()))
      return fail(CERB_ERR_ARG, "conv64s: residual id out of range");
    const Tensor& res = pl->tensors[op.in1];
    if (res.d.n != N || res.d.h != H || res.d.w != W || res.d.c < 64 || res.d.c % 8 != 0 ||
        res.d.dtype != CERB_F16 || !res.plane[1])
      return fail(CERB_ERR_ARG, "conv64s: residual shape mismatch");
    if ((rc = act_map(&p.res_hi, res, 0, 0, tile_box))) return rc;
    if ((rc = act_map(&p.res_lo, res, 1, 0, tile_box))) return rc;
    p.has_res = 1;
  }
  p.relu = op.relu;
  if (op.w_shift < -60 || op.w_shift > 60) return fail(CERB_ERR_ARG, "conv64s: w_shift out of range");
  p.acc_scale = ldexpf(1.0f, -op.w_shift);
  p.err_flag = ctx->err_flag_dev;
  p.prof = ctx->prof_dev;
  p.tile_counter = ctx->dyn_sched ? pl->cur_counter : nullptr;
  conv64s_plan(p);
  return CERB_OK;
}

// Wide 3x3 stride-1 layers with Cout % 256 == 0 on CTA pairs (csrc/conv3x3c2.cu).
int build_conv3_pair(cerb_plan* pl, const cerb_op& op, Step& st) {
  cerb_ctx* ctx = pl->ctx;
  const Tensor& in = pl->tensors[op.in0];
  const Tensor& out = pl->tensors[op.out];
  Conv3c2Params& p = st.c3p;
  memset(&p, 0, sizeof(p));
  st.use3p = true;
  const int H = out.d.h, W = out.d.w, N = out.d.n;
  const size_t es = 2;
  p.n_img = N;
  p.H = H;
  p.W = W;
  p.n_chunks = op.in_c / 64;
  p.BN = op.cout % 256 == 0 ? 256 : op.cout;  // 256, or the layer's 128 / 64
  p.n_ntiles = op.cout / p.BN;
  if (op.in_coff % 8 != 0 || op.in_coff + op.in_c > in.d.c || in.d.c % 8 != 0)
    return fail(CERB_ERR_ARG, "conv3x3c2: bad input channels");
  if (op.out_coff % 8 != 0 || op.out_coff + op.cout > out.d.c || out.d.c % 8 != 0)
    return fail(CERB_ERR_ARG, "conv3x3c2: bad output channels");
  {
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(op.in_c), static_cast<cuuint64_t>(W),
                                static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(in.d.c) * es,
                                   static_cast<cuuint64_t>(W) * in.d.c * es,
                                   static_cast<cuuint64_t>(H) * W * in.d.c * es};
    const cuuint32_t box[4] = {64, 10, 18, 1};
    int rc = encode_map(ctx, &p.l0.in_map, static_cast<__half*>(in.plane[0]) + op.in_coff, 4, dims,
                        strides, box);
    if (rc) return rc;
  }
  const size_t k_total = static_cast<size_t>(9) * op.in_c;
  if (op.w_off < 0 || op.w_off % 16 != 0 ||
      static_cast<size_t>(op.w_off) + static_cast<size_t>(op.cout) * k_total * es > pl->blob_bytes)
    return fail(CERB_ERR_ARG, "conv3x3c2: weight offset out of range");
  {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(k_total), static_cast<cuuint64_t>(op.cout)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(k_total) * es};
    const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(p.BN / 2)};
    int rc = encode_map(ctx, &p.l0.w_map, pl->blob + op.w_off, 2, dims, strides, box);
    if (rc) return rc;
  }
  if (op.b_off >= 0) {
    if (op.b_off % 16 != 0 || static_cast<size_t>(op.b_off) + op.cout * 4u > pl->blob_bytes)
      return fail(CERB_ERR_ARG, "conv3x3c2: bias offset out of range");
    p.l0.bias = reinterpret_cast<const float*>(pl->blob + op.b_off);
  }
  const cuuint64_t odims[4] = {static_cast<cuuint64_t>(op.cout), static_cast<cuuint64_t>(W),
                               static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
  const cuuint32_t obox[4] = {64, 8, 16, 1};
  {
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(out.d.c) * es,
                                   static_cast<cuuint64_t>(W) * out.d.c * es,
                                   static_cast<cuuint64_t>(H) * W * out.d.c * es};
    int rc = encode_map(ctx, &p.l0.out_map, static_cast<__half*>(out.plane[0]) + op.out_coff, 4, odims,
                        strides, obox);
    if (rc) return rc;
  }
  if (op.in1 >= 0) {
    if (op.in1 >= static_cast<int>(pl->tensors.size()))
      return fail(CERB_ERR_ARG, "conv3x3c2: residual id out of range");
    const Tensor& res = pl->tensors[op.in1];
    if (res.d.n != N || res.d.h != H || res.d.w != W || res.d.c < op.cout || res.d.c % 8 != 0 ||
        res.d.dtype != CERB_F16)
      return fail(CERB_ERR_ARG, "conv3x3c2: residual shape mismatch");
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(res.d.c) * es,
                                   static_cast<cuuint64_t>(W) * res.d.c * es,
                                   static_cast<cuuint64_t>(H) * W * res.d.c * es};
    int rc = encode_map(ctx, &p.l0.res_map, res.plane[0], 4, odims, strides, obox);
    if (rc) return rc;
    p.l0.has_res = 1;
  }
  p.l0.relu = op.relu;
  if (op.w_shift < -60 || op.w_shift > 60) return fail(CERB_ERR_ARG, "conv3x3c2: w_shift out of range");
  p.l0.acc_scale = ldexpf(1.0f, -op.w_shift);
  p.err_flag = ctx->err_flag_dev;
  p.prof = ctx->prof_dev;
  p.tile_counter = ctx->dyn_sched ? pl->cur_counter : nullptr;
  conv3x3c2_plan(p);
  return CERB_OK;
}

int build_conv(cerb_plan* pl, const cerb_op& op, Step& st) {
  cerb_ctx* ctx = pl->ctx;
  const int nt = static_cast<int>(pl->tensors.size());
  if (op.in0 < 0 || op.in0 >= nt || op.out < 0 || op.out >= nt)
    return fail(CERB_ERR_ARG, "conv: tensor id out of range");
  const Tensor& in = pl->tensors[op.in0];
  const Tensor& out = pl->tensors[op.out];
  if (op.aux_classes > 0 && op.kh == 3) {
    // 64->64 3x3 conv with the whole classification head fused behind it (csrc/conv64.cu)
    if (ctx->precision != CERB_PREC_F16 || op.stem || op.kw != 3 || op.stride != 1 || op.pad != 1 ||
        op.in_c != 64 || op.cout != 64 || in.d.dtype != CERB_F16 || in.d.n != out.d.n)
      return fail(CERB_ERR_ARG, "conv: a fused head tail needs the 64->64 3x3 kernel in CERB_PREC_F16");
    return build_conv64(pl, op, st);
  }
  const bool fused_head = op.aux_classes > 0;
  if (in.d.dtype != CERB_F16 || out.d.dtype != (fused_head ? CERB_F32 : CERB_F16))
    return fail(CERB_ERR_ARG, "conv: tensors must be fp16 (fp32 canvas for a fused head)");
  const bool split = ctx->precision == CERB_PREC_F16X2;
  ConvKParams& p = st.conv;
  memset(&p, 0, sizeof(p));
  st.split = split;

  // a fused head writes the (centre-cropped) canvas; its geometry is the input's
  const int H = fused_head ? in.d.h : out.d.h, W = fused_head ? in.d.w : out.d.w, N = out.d.n;
  if (in.d.n != N) return fail(CERB_ERR_ARG, "conv: batch mismatch");
  if (fused_head) {
    const int width = op.head_mode == CERB_HEAD_INST ? op.aux_classes - 1 : 1;
    if (op.cout != 96 || op.kh != 1 || op.kw != 1 || op.stride != 1 || op.stem || op.in1 >= 0 ||
        op.aux_classes < 2 || op.aux_classes > 8 || out.d.h > H || out.d.w > W || op.out_coff < 0 ||
        op.out_coff + width > out.d.c || op.aux_w_off < 0 || op.aux_b_off < 0 ||
        op.aux_w_off % 16 != 0 ||
        static_cast<size_t>(op.aux_w_off) + op.aux_classes * 96 * 4u > pl->blob_bytes ||
        static_cast<size_t>(op.aux_b_off) + op.aux_classes * 4u > pl->blob_bytes)
      return fail(CERB_ERR_ARG, "conv: bad fused-head description");
  } else if (op.cout <= 0 || op.cout % 16 != 0 || op.out_coff % 8 != 0 ||
             op.out_coff + op.cout > out.d.c || out.d.c % 8 != 0) {
    return fail(CERB_ERR_ARG, "conv: bad output channels (cout %d coff %d c %d)", op.cout,
                op.out_coff, out.d.c);
  }

  if (split && ctx->conv64s && !op.stem && !fused_head && op.kh == 3 && op.kw == 3 && op.stride == 1 &&
      op.pad == 1 && op.in_c == 64 && op.cout == 64 && op.up_prev1 <= 0 && in.d.dtype == CERB_F16 &&
      out.d.dtype == CERB_F16 && in.d.h == out.d.h && in.d.w == out.d.w && in.d.n == out.d.n)
    return build_conv64s(pl, op, st);
  const bool conv64_ok = !split && !op.stem && !fused_head && ctx->conv64_mode >= 0 && op.kh == 3 &&
                         op.kw == 3 && op.stride == 1 && op.pad == 1 && op.in_c == 64 &&
                         op.cout == 64 && in.d.h == H && in.d.w == W;
  if (op.up_prev1 > 0 && !(conv64_ok && (ctx->conv64_mode == 1 || ctx->conv64_mode == 3)))
    return fail(CERB_ERR_ARG, "conv: fused upsample+add needs the 64->64 3x3 kernel (conv64_mode 1 or 3, "
                "CERB_PREC_F16)");
  if (op.stem && !split && !fused_head && ctx->conv64_mode >= 1 && ctx->stem_mode == 1 &&
      in.d.dtype == CERB_F16 && out.d.dtype == CERB_F16)
    return build_stem64(pl, op, st);
  if (conv64_ok && ctx->conv64_mode == 3) return build_conv64x(pl, op, st);
  if (conv64_ok) return build_conv64(pl, op, st);
  // wide 3x3 stride-1 layers: halo reuse + two M tiles per weight slab (csrc/conv3x3.cu)
  const bool conv3_ok = !split && !op.stem && !fused_head && ctx->conv3_mode > 0 && op.kh == 3 &&
                        op.kw == 3 && op.stride == 1 && op.pad == 1 && op.in_c >= 128 &&
                        op.in_c % 64 == 0 && op.cout % 128 == 0 &&
                        (ctx->conv3_mode > 1 || op.cout <= 512) && in.d.h == H && in.d.w == W &&
                        op.in0 < nt && op.out < nt;
  const bool conv3_shape = !split && !op.stem && !fused_head && ctx->conv3_mode > 0 && op.kh == 3 &&
                           op.kw == 3 && op.stride == 1 && op.pad == 1 && op.in_c >= 128 &&
                           op.in_c % 64 == 0 && in.d.h == H && in.d.w == W;
  // conv3_pair 1 (default): the layers with Cout % 256 == 0 (layer3 / layer4), where the pair kernel
  // measured faster (0.044 vs 0.048 ms at batch 32); 2: also Cout = 128 / 64, where it measured
  // SLOWER than conv3x3.cu (0.063 vs 0.048 ms on 128->128 at 64x64: both kernels sit at the L2 -> SM
  // delivery limit there and the pair kernel's items are half as long)
  if (conv3_shape && ctx->conv3_pair > 0 &&
      ((op.cout % 256 == 0 && (ctx->conv3_mode > 1 || op.cout <= 512)) ||
       (ctx->conv3_pair > 1 && (op.cout == 64 || op.cout == 128))))
    return build_conv3_pair(pl, op, st);
  if (conv3_ok) return build_conv3(pl, op, st);
  int bw = op.box_w > 0 ? op.box_w : pick_box_w(H, W);
  if (bw > 128 || (bw & (bw - 1)) != 0) return fail(CERB_ERR_ARG, "conv: bad box_w %d", bw);
  const int bh = 128 / bw;
  p.bw_log2 = floor_log2(bw);
  p.n_img = N;
  p.H = H;
  p.W = W;
  p.tiles_x = (W + bw - 1) / bw;
  p.tiles_y = (H + bh - 1) / bh;
  const long m_tiles = static_cast<long>(N) * p.tiles_x * p.tiles_y;
  p.BN = pick_bn(op.cout, m_tiles, ctx->num_sms);
  if (split && p.BN > 128) p.BN = 128;  // three TMEM accumulators per tile in split mode
  if (p.BN <= 0 || p.BN > 256 || p.BN % 16 != 0 || op.cout % p.BN != 0)
    return fail(CERB_ERR_ARG, "conv: unsupported cout %d", op.cout);
  p.n_ntiles = op.cout / p.BN;
  p.n_tiles = static_cast<int>(m_tiles * p.n_ntiles);

  const size_t es = 2;
  int k_total = 0;
  if (op.stem) {
    // in = PREP tensor [N, H, W+8, 8]; one K chunk per filter row = 8 px x 8 ch window.
    if (in.d.c != 8 || in.d.h != H || in.d.w != W + 8 || op.kh != 7 || op.kw != 7 ||
        op.stride != 1 || op.pad != 3)
      return fail(CERB_ERR_ARG, "conv: stem expects a 7x7 s1 p3 conv over a PREP tensor");
    p.n_taps = 7;
    p.n_chunks = 1;
    for (int r = 0; r < 7; ++r) {
      p.taps[r].map = 0;
      p.taps[r].dx = 0;
      p.taps[r].dy = static_cast<int8_t>(r - 3);
    }
    const cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                                static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {8 * es, static_cast<cuuint64_t>(W + 8) * 8 * es,
                                   static_cast<cuuint64_t>(H) * (W + 8) * 8 * es};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), 1};
    for (int pln = 0; pln < (split ? 2 : 1); ++pln) {
      CUtensorMap* m = pln == 0 ? &p.in_hi[0] : &p.in_lo[0];
      int rc = encode_map(ctx, m, in.plane[pln], 4, dims, strides, box);
      if (rc) return rc;
    }
    k_total = 7 * 64;
    p.a_lo_zero = split ? 1 : 0;  // PREP writes integers 0..255 (exact in fp16) and zeroes the lo plane
  } else {
    const int s = op.stride;
    if (s != 1 && s != 2) return fail(CERB_ERR_ARG, "conv: stride must be 1 or 2");
    if (op.in_c <= 0 || op.in_c % 64 != 0 || op.in_coff % 8 != 0 ||
        op.in_coff + op.in_c > in.d.c || in.d.c % 8 != 0)
      return fail(CERB_ERR_ARG, "conv: bad input channels (in_c %d coff %d c %d)", op.in_c,
                  op.in_coff, in.d.c);
    const int eh = (in.d.h + 2 * op.pad - op.kh) / s + 1;
    const int ew = (in.d.w + 2 * op.pad - op.kw) / s + 1;
    if (eh != H || ew != W)
      return fail(CERB_ERR_ARG, "conv: output is %dx%d but geometry gives %dx%d", H, W, eh, ew);
    if (op.kh * op.kw > kConvMaxTaps) return fail(CERB_ERR_ARG, "conv: too many taps");
    p.n_chunks = op.in_c / 64;
    p.n_taps = op.kh * op.kw;
    bool used[4] = {false, false, false, false};
    for (int r = 0; r < op.kh; ++r) {
      for (int q = 0; q < op.kw; ++q) {
        ConvTap& t = p.taps[r * op.kw + q];
        const int ty = r - op.pad, tx = q - op.pad;
        if (s == 1) {
          t.map = 0;
          t.dy = static_cast<int8_t>(ty);
          t.dx = static_cast<int8_t>(tx);
        } else {
          const int py = ((ty % 2) + 2) % 2, px = ((tx % 2) + 2) % 2;
          t.map = static_cast<int8_t>(py * 2 + px);
          t.dy = static_cast<int8_t>((ty - py) / 2);
          t.dx = static_cast<int8_t>((tx - px) / 2);
        }
        used[t.map] = true;
      }
    }
    for (int mi = 0; mi < 4; ++mi) {
      if (!used[mi]) continue;
      const int py = mi >> 1, px = mi & 1;
      const int vh = (in.d.h - py + s - 1) / s, vw = (in.d.w - px + s - 1) / s;
      if (vh <= 0 || vw <= 0) return fail(CERB_ERR_ARG, "conv: empty parity view");
      const cuuint64_t dims[4] = {static_cast<cuuint64_t>(op.in_c), static_cast<cuuint64_t>(vw),
                                  static_cast<cuuint64_t>(vh), static_cast<cuuint64_t>(N)};
      const cuuint64_t strides[3] = {static_cast<cuuint64_t>(s) * in.d.c * es,
                                     static_cast<cuuint64_t>(s) * in.d.w * in.d.c * es,
                                     static_cast<cuuint64_t>(in.d.h) * in.d.w * in.d.c * es};
      const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), 1};
      const size_t off = (static_cast<size_t>(py) * in.d.w + px) * in.d.c + op.in_coff;
      for (int pln = 0; pln < (split ? 2 : 1); ++pln) {
        CUtensorMap* m = pln == 0 ? &p.in_hi[mi] : &p.in_lo[mi];
        int rc = encode_map(ctx, m, static_cast<__half*>(in.plane[pln]) + off, 4, dims, strides,
                            box);
        if (rc) return rc;
      }
    }
    k_total = p.n_taps * op.in_c;
  }

  // weights [cout][k_total] fp16, K-major
  if (op.w_off < 0 || static_cast<size_t>(op.w_off) + static_cast<size_t>(op.cout) * k_total * es >
                          pl->blob_bytes || op.w_off % 16 != 0)
    return fail(CERB_ERR_ARG, "conv: weight offset out of range");
  if (split && (op.w_lo_off < 0 || static_cast<size_t>(op.w_lo_off) +
                                           static_cast<size_t>(op.cout) * k_total * es >
                                       pl->blob_bytes || op.w_lo_off % 16 != 0))
    return fail(CERB_ERR_ARG, "conv: lo weight offset out of range");
  {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(k_total), static_cast<cuuint64_t>(op.cout)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(k_total) * es};
    const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(p.BN)};
    int rc = encode_map(ctx, &p.w_hi, pl->blob + op.w_off, 2, dims, strides, box);
    if (rc) return rc;
    if (split) {
      rc = encode_map(ctx, &p.w_lo, pl->blob + op.w_lo_off, 2, dims, strides, box);
      if (rc) return rc;
    }
  }
  if (op.b_off >= 0) {
    if (op.b_off % 16 != 0 || static_cast<size_t>(op.b_off) + op.cout * 4u > pl->blob_bytes)
      return fail(CERB_ERR_ARG, "conv: bias offset out of range");
    p.bias = reinterpret_cast<const float*>(pl->blob + op.b_off);
  }
  if (fused_head) {
    p.head_w = reinterpret_cast<const float*>(pl->blob + op.aux_w_off);
    p.head_b = reinterpret_cast<const float*>(pl->blob + op.aux_b_off);
    p.head_classes = op.aux_classes;
    p.head_mode = op.head_mode;
    if (op.cout == 96 && op.b_off >= 0) {  // biases as kernel parameters (see conv_tc.cuh)
      CERB_CUDA(cudaMemcpy(p.head_hbias, pl->blob + op.b_off, 96 * sizeof(float), cudaMemcpyDeviceToHost));
      CERB_CUDA(cudaMemcpy(p.head_obias, pl->blob + op.aux_b_off, op.aux_classes * sizeof(float),
                           cudaMemcpyDeviceToHost));
    }
    p.canvas = static_cast<float*>(out.plane[0]);
    p.oh = out.d.h;
    p.ow = out.d.w;
    p.canvas_c = out.d.c;
    p.canvas_coff = op.out_coff;
    if (op.logits_out >= 0) {
      if (op.logits_out >= nt) return fail(CERB_ERR_ARG, "conv: logits id out of range");
      const Tensor& lg = pl->tensors[op.logits_out];
      if (lg.d.dtype != CERB_F32 || lg.d.n != N || lg.d.h != H || lg.d.w != W ||
          lg.d.c != op.aux_classes)
        return fail(CERB_ERR_ARG, "conv: fused-head logits tensor mismatch");
      p.logits = static_cast<float*>(lg.plane[0]);
    }
  } else {
    p.out_hi = static_cast<__half*>(out.plane[0]);
    p.out_lo = static_cast<__half*>(out.plane[1]);
    p.out_cs = out.d.c;
    p.out_coff = op.out_coff;
  }
  if (op.in1 >= 0) {
    if (op.in1 >= nt) return fail(CERB_ERR_ARG, "conv: residual id out of range");
    const Tensor& res = pl->tensors[op.in1];
    if (res.d.n != N || res.d.h != H || res.d.w != W || res.d.c < op.cout || res.d.c % 8 != 0 ||
        res.d.dtype != CERB_F16)
      return fail(CERB_ERR_ARG, "conv: residual shape mismatch");
    p.res_hi = static_cast<const __half*>(res.plane[0]);
    p.res_lo = static_cast<const __half*>(res.plane[1]);
    p.res_cs = res.d.c;
    p.res_coff = 0;
  }
  p.relu = op.relu;
  if (op.w_shift < -60 || op.w_shift > 60) return fail(CERB_ERR_ARG, "conv: w_shift out of range");
  p.acc_scale = ldexpf(1.0f, -op.w_shift);
  p.err_flag = ctx->err_flag_dev;
  p.prof = ctx->prof_dev;
  p.tile_counter = ctx->dyn_sched ? pl->cur_counter : nullptr;
  conv_tc_plan_pipeline(p, split);
  return CERB_OK;
}

int check_id(const cerb_plan* pl, int id, const char* what) {
  if (id < 0 || id >= static_cast<int>(pl->tensors.size()))
    return fail(CERB_ERR_ARG, "%s: tensor id %d out of range", what, id);
  return CERB_OK;
}

}  // namespace

// ------------------------------------------------------------------------ context
extern "C" int cerb_ctx_create(int device, int precision, cerb_ctx** out) {
  if (!out) return fail(CERB_ERR_ARG, "cerb_ctx_create: out is null");
  *out = nullptr;
  if (precision != CERB_PREC_F16 && precision != CERB_PREC_F16X2)
    return fail(CERB_ERR_ARG, "cerb_ctx_create: unknown precision %d", precision);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(CERB_ERR_NO_DEVICE, "no CUDA device (%s); cerberus_b200 has no CPU path",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= count)
    return fail(CERB_ERR_ARG, "cerb_ctx_create: device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  CERB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(CERB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
  CERB_CUDA(cudaSetDevice(device));
  cerb_ctx* ctx = new cerb_ctx();
  ctx->device = device;
  ctx->precision = precision;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->conv_sms = ctx->num_sms;
  CERB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CERB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  CERB_CUDA(cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
  CERB_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
  CERB_CUDA(cudaEventCreateWithFlags(&ctx->side_event, cudaEventDisableTiming));
  for (int i = 0; i < 8; ++i) {
    CERB_CUDA(cudaEventCreateWithFlags(&ctx->slot_event[i], cudaEventDisableTiming));
    CERB_CUDA(cudaEventCreateWithFlags(&ctx->mark_event[i], cudaEventDisableTiming));
  }
  CERB_CUDA(cudaEventCreateWithFlags(&ctx->order_event, cudaEventDisableTiming));
  CERB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ctx->err_flag_host), sizeof(int) * 4,
                          cudaHostAllocMapped));
  memset(ctx->err_flag_host, 0, sizeof(int) * 4);
  CERB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->err_flag_dev),
                                     ctx->err_flag_host, 0));
  CERB_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->stat_dev), 4 * sizeof(unsigned long long)));
  CERB_CUDA(cudaMemset(ctx->stat_dev, 0, 4 * sizeof(unsigned long long)));
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  CERB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess)
    return fail(CERB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  ctx->encode_tiled = fn;
  *out = ctx;
  return CERB_OK;
}

extern "C" void cerb_ctx_destroy(cerb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->up_stream) cudaStreamDestroy(ctx->up_stream);
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  if (ctx->side_event) cudaEventDestroy(ctx->side_event);
  for (int i = 0; i < 8; ++i) {
    if (ctx->slot_event[i]) cudaEventDestroy(ctx->slot_event[i]);
    if (ctx->mark_event[i]) cudaEventDestroy(ctx->mark_event[i]);
  }
  if (ctx->order_event) cudaEventDestroy(ctx->order_event);
  if (ctx->err_flag_host) cudaFreeHost(ctx->err_flag_host);
  if (ctx->prof_dev) cudaFree(ctx->prof_dev);
  if (ctx->stat_dev) cudaFree(ctx->stat_dev);
  for (void* p : ctx->scratch) cudaFree(p);
  if (ctx->postproc_ws && ctx->postproc_ws_free) ctx->postproc_ws_free(ctx->postproc_ws);
  if (ctx->instinfo_ws && ctx->instinfo_ws_free) ctx->instinfo_ws_free(ctx->instinfo_ws);
  for (int i = 0; i < cerb_ctx::kParamSlots; ++i)
    if (ctx->param_event[i]) cudaEventDestroy(ctx->param_event[i]);
  if (ctx->param_host) cudaFreeHost(ctx->param_host);
  if (ctx->param_dev) cudaFree(ctx->param_dev);
  delete ctx;
}

extern "C" const char* cerb_last_error(void) { return g_last_error.c_str(); }

extern "C" int cerb_ctx_sync(cerb_ctx* ctx) {
  if (!ctx) return fail(CERB_ERR_ARG, "cerb_ctx_sync: null ctx");
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  const int flag = ctx->err_flag_host ? ctx->err_flag_host[0] : 0;
  if (flag != 0) {
    return fail(CERB_ERR_KERNEL, "kernel pipeline watchdog fired (code %d: %s)", flag,
                flag == 1 ? "TMA producer waiting for a free smem stage"
                : flag == 2 ? "MMA issuer waiting for a drained accumulator"
                : flag == 3 ? "MMA issuer waiting for TMA data"
                : flag == 4 ? "epilogue waiting for the accumulator"
                            : "unknown");
  }
  if (e != cudaSuccess) return fail(CERB_ERR_CUDA, "stream sync: %s", cudaGetErrorString(e));
  const int pp = ctx->err_flag_host ? ctx->err_flag_host[1] : 0;
  if (pp != 0) {
    ctx->err_flag_host[1] = 0;
    return fail(CERB_ERR_KERNEL,
                pp == 10   ? "postproc: more instances than the size filter allows"
                : pp == 11 ? "postproc: an instance crop does not fit in shared memory"
                           : "postproc: kernel error %d",
                pp);
  }
  return CERB_OK;
}

extern "C" int cerb_ctx_set_option(cerb_ctx* ctx, const char* name, int value) {
  if (!ctx || !name) return fail(CERB_ERR_ARG, "cerb_ctx_set_option: bad arguments");
  if (strcmp(name, "conv64_mode") == 0) {
    if (value < -1 || value > 3) return fail(CERB_ERR_ARG, "conv64_mode must be -1, 0, 1, 2 or 3");
    ctx->conv64_mode = value;
    return CERB_OK;
  }
  if (strcmp(name, "conv64s") == 0) {
    ctx->conv64s = value != 0;
    return CERB_OK;
  }
  if (strcmp(name, "conv3_pair") == 0) {
    ctx->conv3_pair = value;
    return CERB_OK;
  }
  if (strcmp(name, "fuse_upadd") == 0) {
    ctx->fuse_upadd = value != 0;  // fold skip + bilinear_x2(low) into the 64->64 convolution that reads it
    return CERB_OK;
  }
  if (strcmp(name, "conv3_chain") == 0) {
    ctx->conv3_chain = value != 0;  // consecutive pair-kernel layers of one geometry in one launch
    return CERB_OK;
  }
  if (strcmp(name, "conv3_mode") == 0) {
    if (value < 0 || value > 2) return fail(CERB_ERR_ARG, "conv3_mode must be 0, 1 or 2");
    ctx->conv3_mode = value;
    return CERB_OK;
  }
  if (strcmp(name, "conv_sms") == 0) {
    // persistent convolution kernels launch at most this many CTAs (one per SM); the rest of the
    // SMs stay free for kernels of other contexts (post-processing blocks that need a whole SM).
    // Applies to launches / graph captures made afterwards.
    if (value < 1) return fail(CERB_ERR_ARG, "conv_sms must be >= 1");
    ctx->conv_sms = value < ctx->num_sms ? value : ctx->num_sms;
    return CERB_OK;
  }
  if (strcmp(name, "dyn_sched") == 0) {
    ctx->dyn_sched = value != 0;  // persistent conv kernels take tiles from a global counter
    return CERB_OK;
  }
  if (strcmp(name, "stem_mode") == 0) {
    ctx->stem_mode = value;  // 1: stem on the resident-weight halo kernel, 0: generic kernel
    return CERB_OK;
  }
  if (strcmp(name, "use_pdl") == 0) {
    ctx->use_pdl = value != 0;  // programmatic dependent launch of the convolution kernels
    return CERB_OK;
  }
  if (strcmp(name, "k_rotate") == 0) {
    ctx->k_rotate = value != 0;
    return CERB_OK;
  }
  if (strcmp(name, "use_graphs") == 0) {
    ctx->use_graphs = value != 0;
    return CERB_OK;
  }
  if (strcmp(name, "ws_mode") == 0) {
    ctx->ws_mode = value;
    return CERB_OK;
  }
  if (strcmp(name, "conv64_debug") == 0) {
    ctx->conv64_debug = value;
    return CERB_OK;
  }
  if (strcmp(name, "kernel_prof") == 0) {
    // plans created afterwards hand the 64->64 kernel a counter buffer (cerb_ctx_read_prof)
    CERB_CUDA(cudaSetDevice(ctx->device));
    if (value != 0 && ctx->prof_dev == nullptr) {
      CERB_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->prof_dev), kProfSlots * sizeof(long long)));
      CERB_CUDA(cudaMemset(ctx->prof_dev, 0, kProfSlots * sizeof(long long)));
    } else if (value == 0 && ctx->prof_dev != nullptr) {
      cudaFree(ctx->prof_dev);
      ctx->prof_dev = nullptr;
    }
    return CERB_OK;
  }
  return fail(CERB_ERR_ARG, "cerb_ctx_set_option: unknown option %s", name);
}

extern "C" int cerb_copy_async(cerb_ctx* ctx, void* dst, const void* src, size_t bytes, int kind) {
  if (!ctx || !dst || !src || (kind != 1 && kind != 2))
    return fail(CERB_ERR_ARG, "cerb_copy_async: bad arguments");
  CERB_CUDA(cudaSetDevice(ctx->device));
  CERB_CUDA(cudaMemcpyAsync(dst, src, bytes,
                            kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                            kind == 1 ? ctx->up_stream : ctx->copy_stream));
  return CERB_OK;
}

extern "C" int cerb_copy_mark(cerb_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot > 7) return fail(CERB_ERR_ARG, "cerb_copy_mark: bad arguments");
  CERB_CUDA(cudaSetDevice(ctx->device));
  CERB_CUDA(cudaEventRecord(ctx->slot_event[slot], ctx->copy_stream));
  return CERB_OK;
}

extern "C" int cerb_copy_wait(cerb_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot > 7) return fail(CERB_ERR_ARG, "cerb_copy_wait: bad arguments");
  CERB_CUDA(cudaSetDevice(ctx->device));
  CERB_CUDA(cudaEventSynchronize(ctx->slot_event[slot]));
  const int flag = ctx->err_flag_host ? ctx->err_flag_host[0] : 0;
  if (flag != 0) return fail(CERB_ERR_KERNEL, "kernel pipeline watchdog fired (code %d)", flag);
  return CERB_OK;
}

extern "C" int cerb_stream_order(cerb_ctx* ctx, int copy_waits_for_compute) {
  if (!ctx) return fail(CERB_ERR_ARG, "cerb_stream_order: null ctx");
  CERB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t from = copy_waits_for_compute ? ctx->stream : ctx->up_stream;
  cudaStream_t to = copy_waits_for_compute ? ctx->copy_stream : ctx->stream;
  CERB_CUDA(cudaEventRecord(ctx->order_event, from));
  CERB_CUDA(cudaStreamWaitEvent(to, ctx->order_event, 0));
  return CERB_OK;
}

extern "C" int cerb_ctx_wait(cerb_ctx* waiter, cerb_ctx* signal) {
  if (!waiter || !signal || waiter->device != signal->device)
    return fail(CERB_ERR_ARG, "cerb_ctx_wait: both contexts must live on the same device");
  CERB_CUDA(cudaSetDevice(waiter->device));
  CERB_CUDA(cudaEventRecord(signal->order_event, signal->stream));
  CERB_CUDA(cudaStreamWaitEvent(waiter->stream, signal->order_event, 0));
  return CERB_OK;
}

extern "C" int cerb_ctx_mark(cerb_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot > 7) return fail(CERB_ERR_ARG, "cerb_ctx_mark: bad arguments");
  CERB_CUDA(cudaSetDevice(ctx->device));
  CERB_CUDA(cudaEventRecord(ctx->mark_event[slot], ctx->stream));
  return CERB_OK;
}

extern "C" int cerb_ctx_wait_mark(cerb_ctx* waiter, cerb_ctx* signal, int slot) {
  if (!waiter || !signal || waiter->device != signal->device || slot < 0 || slot > 7)
    return fail(CERB_ERR_ARG, "cerb_ctx_wait_mark: bad arguments");
  CERB_CUDA(cudaSetDevice(waiter->device));
  CERB_CUDA(cudaStreamWaitEvent(waiter->stream, signal->mark_event[slot], 0));
  return CERB_OK;
}

extern "C" int cerb_copy_sync(cerb_ctx* ctx) {
  if (!ctx) return fail(CERB_ERR_ARG, "cerb_copy_sync: null ctx");
  CERB_CUDA(cudaSetDevice(ctx->device));
  CERB_CUDA(cudaStreamSynchronize(ctx->up_stream));
  CERB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  return CERB_OK;
}

extern "C" int cerb_ctx_read_prof(cerb_ctx* ctx, int64_t* out, int n, int reset) {
  if (!ctx || !out || n <= 0 || n > kProfSlots) return fail(CERB_ERR_ARG, "cerb_ctx_read_prof: bad arguments");
  if (ctx->prof_dev == nullptr) return fail(CERB_ERR_ARG, "cerb_ctx_read_prof: option kernel_prof is off");
  CERB_CUDA(cudaSetDevice(ctx->device));
  CERB_CUDA(cudaStreamSynchronize(ctx->stream));
  CERB_CUDA(cudaMemcpy(out, ctx->prof_dev, n * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (reset) CERB_CUDA(cudaMemset(ctx->prof_dev, 0, kProfSlots * sizeof(long long)));
  return CERB_OK;
}

extern "C" int64_t cerb_ctx_launch_count(cerb_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void* cerb_ctx_stream(cerb_ctx* ctx) { return ctx ? ctx->stream : nullptr; }

// ------------------------------------------------------------------------ plan
namespace {
// Runs of consecutive wide 3x3 convolutions of identical geometry on one kernel (CTA-pair kernel or
// conv3x3.cu), each reading the output of the one before (the bodies of encoder layer2 / layer3 /
// layer4), become ONE launch: see conv_chain.cuh.
// A residual may be any tensor written before the step that consumes it (the chain's dependency
// rule - layer l of an image starts when layer l - 1 of that image is stored - orders it too).
// out = skip + bilinear_x2(low) followed by the 64->64 3x3 convolution that is its ONLY reader
// (decoder levels u2 / u1, models/net_desc.py:183-189): the convolution reads the skip tensor
// itself and its fix-up warps add the upsampled tensor inside the halo (csrc/conv64x.cu), so the
// sum never travels through HBM. fp16 mode, default 64->64 kernel (conv64_mode 3) only.
void fuse_upadd_ops(const cerb_ctx* ctx, const cerb_tensor_desc* tensors, int n_tensors,
                    std::vector<cerb_op>& ops, std::vector<char>& folded) {
  if (!ctx->fuse_upadd || ctx->precision != CERB_PREC_F16 || ctx->conv64_mode != 3) return;
  const int n = static_cast<int>(ops.size());
  auto valid = [&](int t) { return t >= 0 && t < n_tensors; };
  for (int i = 0; i + 1 < n; ++i) {
    const cerb_op& u = ops[i];
    cerb_op& c = ops[i + 1];
    if (u.kind != CERB_OP_UPADD || u.cout > 1 || c.kind != CERB_OP_CONV || u.side || c.side) continue;
    if (!valid(u.in0) || !valid(u.in1) || !valid(u.out) || !valid(c.out)) continue;
    if (c.in0 != u.out || c.in_coff != 0 || c.in_c != 64 || c.cout != 64 || c.kh != 3 || c.kw != 3 ||
        c.stride != 1 || c.pad != 1 || c.stem || c.aux_classes > 0 || c.up_prev1 > 0 || c.in1 == u.out ||
        c.out == u.in0 || c.out == u.in1)
      continue;
    const cerb_tensor_desc& sk = tensors[u.in0];
    const cerb_tensor_desc& lo = tensors[u.in1];
    const cerb_tensor_desc& su = tensors[u.out];
    if (sk.dtype != CERB_F16 || lo.dtype != CERB_F16 || su.dtype != CERB_F16 || sk.c != 64 || lo.c != 64 ||
        su.c != 64 || sk.h != su.h || sk.w != su.w || sk.n != su.n || lo.n != su.n || lo.h * 2 != su.h ||
        lo.w * 2 != su.w)
      continue;
    // does anything else read THIS sum? (the tensor may be reused: scan up to its next writer)
    bool other_reader = false;
    for (int k = i + 2; k < n; ++k) {
      const cerb_op& o = ops[k];
      if (o.in0 == u.out || o.in1 == u.out || o.up_prev1 - 1 == u.out) {
        other_reader = true;
        break;
      }
      if (o.out == u.out) break;
    }
    if (other_reader) continue;
    c.in0 = u.in0;
    c.up_prev1 = u.in1 + 1;
    folded[i] = 1;
  }
}

// Geometry of a pair-kernel / conv3x3.cu step as far as chaining is concerned.
struct ChainGeom {
  int kernel, n_img, H, W, n_chunks, BN, n_ntiles, n_bstages;
  bool operator==(const ChainGeom& o) const {
    return kernel == o.kernel && n_img == o.n_img && H == o.H && W == o.W && n_chunks == o.n_chunks &&
           BN == o.BN && n_ntiles == o.n_ntiles && n_bstages == o.n_bstages;
  }
};
bool chain_geom(const Step& st, ChainGeom& g) {
  if (st.kind != CERB_OP_CONV || st.side) return false;
  if (st.use3p) {
    const Conv3c2Params& p = st.c3p;
    g = {1, p.n_img, p.H, p.W, p.n_chunks, p.BN, p.n_ntiles, p.n_bstages};
    return true;
  }
  if (st.use3 && !st.c3.rotate) {
    const Conv3Params& p = st.c3;
    g = {2, p.n_img, p.H, p.W, p.n_chunks, p.BN, p.n_ntiles, p.n_bstages};
    return true;
  }
  return false;
}

int link_chains(cerb_plan* pl, const cerb_op* ops, int n_ops) {
  cerb_ctx* ctx = pl->ctx;
  pl->n_launches = 0;
  for (const Step& st : pl->steps) pl->n_launches += st.chained ? 0 : 1;
  if (!ctx->conv3_chain || !ctx->dyn_sched) return CERB_OK;
  int* done_next = pl->tile_counters + n_ops;
  for (int i = 0; i < n_ops;) {
    Step& first = pl->steps[i];
    ChainGeom a;
    if (!chain_geom(first, a)) { ++i; continue; }
    int j = i + 1;
    while (j < n_ops) {
      ChainGeom b;
      if (!chain_geom(pl->steps[j], b) || !(b == a) || ops[j].in0 != ops[j - 1].out ||
          ops[j].in_coff != ops[j - 1].out_coff || ops[j].out == ops[j].in0)
        break;
      ++j;
    }
    const int len = j - i;
    if (len > 1) {
      std::vector<ConvChainLayer> table(static_cast<size_t>(len));
      for (int k = 0; k < len; ++k)
        table[k] = a.kernel == 1 ? pl->steps[i + k].c3p.l0 : pl->steps[i + k].c3.l0;
      void* dev = nullptr;
      const size_t bytes = sizeof(ConvChainLayer) * table.size();
      if (cudaMalloc(&dev, bytes) != cudaSuccess ||
          cudaMemcpy(dev, table.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(CERB_ERR_CUDA, "layer table of a convolution chain: %s",
                    cudaGetErrorString(cudaGetLastError()));
      pl->chain_tables.push_back(dev);
      if (a.kernel == 1) {
        Conv3c2Params& p = first.c3p;
        p.layers = static_cast<const ConvChainLayer*>(dev);
        p.n_layers = len;
        p.done = done_next;
        conv3x3c2_plan(p);
      } else {
        Conv3Params& p = first.c3;
        p.layers = static_cast<const ConvChainLayer*>(dev);
        p.n_layers = len;
        p.done = done_next;
        conv3x3_plan(p);
      }
      done_next += static_cast<size_t>(len) * a.n_img;
      first.chain_len = len;
      for (int k = 1; k < len; ++k) pl->steps[i + k].chained = true;
      pl->n_launches -= len - 1;
    }
    i = j;
  }
  return CERB_OK;
}
}  // namespace

extern "C" int cerb_plan_preview_folding(int precision, const cerb_tensor_desc* tensors, int n_tensors,
                                         const cerb_op* ops, int n_ops, int32_t* folded) {
  if (!tensors || !ops || !folded || n_tensors <= 0 || n_ops <= 0)
    return fail(CERB_ERR_ARG, "cerb_plan_preview_folding: bad arguments");
  cerb_ctx defaults;  // option defaults of a fresh context; never touches a device
  defaults.precision = precision;
  defaults.conv64_mode = 3;
  std::vector<cerb_op> opv(ops, ops + n_ops);
  std::vector<char> f(static_cast<size_t>(n_ops), 0);
  fuse_upadd_ops(&defaults, tensors, n_tensors, opv, f);
  for (int i = 0; i < n_ops; ++i) folded[i] = f[i];
  return CERB_OK;
}

extern "C" void cerb_plan_destroy(cerb_plan* pl) {
  if (!pl) return;
  cudaSetDevice(pl->ctx->device);
  for (Tensor& t : pl->tensors) {
    if (t.plane[0]) cudaFree(t.plane[0]);
    if (t.plane[1]) cudaFree(t.plane[1]);
  }
  if (pl->blob && pl->owns_blob) cudaFree(pl->blob);
  if (pl->tile_counters) cudaFree(pl->tile_counters);
  for (void* t : pl->chain_tables) cudaFree(t);
  if (pl->graph_exec) cudaGraphExecDestroy(pl->graph_exec);
  delete pl;
}

extern "C" int cerb_plan_create(cerb_ctx* ctx, const cerb_tensor_desc* tensors, int n_tensors,
                                const cerb_op* ops, int n_ops, const void* weight_blob,
                                size_t blob_bytes, cerb_plan** out) {
  return cerb::plan_create_impl(ctx, tensors, n_tensors, ops, n_ops, weight_blob, blob_bytes, nullptr, out);
}

int cerb::plan_create_impl(cerb_ctx* ctx, const cerb_tensor_desc* tensors, int n_tensors,
                           const cerb_op* ops_in, int n_ops, const void* weight_blob, size_t blob_bytes,
                           uint8_t* shared_dev_blob, cerb_plan** out) {
  if (!ctx || !tensors || !ops_in || !out || n_tensors <= 0 || n_ops <= 0)
    return fail(CERB_ERR_ARG, "cerb_plan_create: bad arguments");
  *out = nullptr;
  // the op list as executed: UPADD ops whose only reader is an eligible 64->64 convolution are
  // folded into that convolution (fuse_upadd_ops)
  std::vector<cerb_op> opv(ops_in, ops_in + n_ops);
  std::vector<char> folded(static_cast<size_t>(n_ops), 0);
  fuse_upadd_ops(ctx, tensors, n_tensors, opv, folded);
  const cerb_op* ops = opv.data();
  CERB_CUDA(cudaSetDevice(ctx->device));
  cerb_plan* pl = new cerb_plan();
  pl->ctx = ctx;
  int rc = CERB_OK;
  auto bail = [&](int code) {
    cerb_plan_destroy(pl);
    return code;
  };
  const bool split = ctx->precision == CERB_PREC_F16X2;
  pl->blob_bytes = blob_bytes;
  if (shared_dev_blob != nullptr) {
    pl->blob = shared_dev_blob;
    pl->owns_blob = false;
  } else if (blob_bytes > 0) {
    if (cudaMalloc(reinterpret_cast<void**>(&pl->blob), blob_bytes) != cudaSuccess)
      return bail(fail(CERB_ERR_CUDA, "cudaMalloc(%zu) for the weight blob failed", blob_bytes));
    if (cudaMemcpy(pl->blob, weight_blob, blob_bytes, cudaMemcpyHostToDevice) != cudaSuccess)
      return bail(fail(CERB_ERR_CUDA, "weight blob H2D copy failed"));
  }
  pl->tensors.resize(n_tensors);
  for (int i = 0; i < n_tensors; ++i) {
    Tensor& t = pl->tensors[i];
    t.d = tensors[i];
    const size_t es = dtype_size(t.d.dtype);
    if (es == 0 || t.d.n <= 0 || t.d.h <= 0 || t.d.w <= 0 || t.d.c <= 0)
      return bail(fail(CERB_ERR_ARG, "tensor %d: bad descriptor", i));
    t.bytes = static_cast<size_t>(t.d.n) * t.d.h * t.d.w * t.d.c * es;
    const int planes = (t.d.dtype == CERB_F16 && split) ? 2 : 1;
    for (int pnum = 0; pnum < planes; ++pnum) {
      cudaError_t e = cudaMalloc(&t.plane[pnum], t.bytes);
      if (e != cudaSuccess)
        return bail(fail(CERB_ERR_CUDA, "cudaMalloc(%zu) for tensor %d failed: %s", t.bytes, i,
                         cudaGetErrorString(e)));
      cudaMemsetAsync(t.plane[pnum], 0, t.bytes, ctx->stream);
    }
  }
  int max_n = 1;
  for (const Tensor& t : pl->tensors) max_n = t.d.n > max_n ? t.d.n : max_n;
  pl->n_counters = static_cast<size_t>(n_ops) * (1 + static_cast<size_t>(max_n));
  if (cudaMalloc(reinterpret_cast<void**>(&pl->tile_counters), sizeof(int) * pl->n_counters) != cudaSuccess)
    return bail(fail(CERB_ERR_CUDA, "cudaMalloc for the tile counters failed"));
  pl->steps.resize(n_ops);
  for (int i = 0; i < n_ops; ++i) {
    const cerb_op& op = ops[i];
    Step& st = pl->steps[i];
    pl->cur_counter = pl->tile_counters + i;
    st.kind = op.kind;
    st.side = op.side != 0;
    if (folded[i]) {  // an UPADD executed by the producer of the convolution that follows
      st.chained = true;
      continue;
    }
    switch (op.kind) {
      case CERB_OP_PREP: {
        if ((rc = check_id(pl, op.in0, "prep")) || (rc = check_id(pl, op.out, "prep")))
          return bail(rc);
        const Tensor& in = pl->tensors[op.in0];
        const Tensor& o = pl->tensors[op.out];
        if (in.d.dtype != CERB_U8 || in.d.c != 3 || o.d.dtype != CERB_F16 || o.d.c != 8 ||
            o.d.w != in.d.w + 8 || o.d.h != in.d.h || o.d.n != in.d.n)
          return bail(fail(CERB_ERR_ARG, "op %d: PREP expects u8 [N,H,W,3] -> f16 [N,H,W+8,8]", i));
        st.u8_in = static_cast<const uint8_t*>(in.plane[0]);
        st.a = act_ref(o);
        pl->prep_in_tensor = op.in0;
        break;
      }
      case CERB_OP_CONV:
        if ((rc = build_conv(pl, op, st))) {
          g_last_error = "op " + std::to_string(i) + ": " + g_last_error;
          return bail(rc);
        }
        break;
      case CERB_OP_MAXPOOL: {
        if ((rc = check_id(pl, op.in0, "maxpool")) || (rc = check_id(pl, op.out, "maxpool")))
          return bail(rc);
        st.a = act_ref(pl->tensors[op.in0]);
        st.b = act_ref(pl->tensors[op.out]);
        if (st.b.h != (st.a.h + 1) / 2 || st.b.w != (st.a.w + 1) / 2 || st.a.c != st.b.c ||
            st.a.c % 8 != 0 || st.a.n != st.b.n)
          return bail(fail(CERB_ERR_ARG, "op %d: MAXPOOL shape mismatch", i));
        break;
      }
      case CERB_OP_UPADD: {
        if ((rc = check_id(pl, op.in0, "upadd")) || (rc = check_id(pl, op.in1, "upadd")) ||
            (rc = check_id(pl, op.out, "upadd")))
          return bail(rc);
        if (op.cout > 1) {
          // grouped form: G = op.cout decoders share the skip tensor in0; their low-resolution
          // inputs are the tensors in1, in1+1, .., in1+G-1 and their outputs out, .., out+G-1
          if (op.cout > 8) return bail(fail(CERB_ERR_ARG, "op %d: UPADD groups > 8", i));
          st.a = act_ref(pl->tensors[op.in0]);
          st.up_groups = op.cout;
          for (int d = 0; d < op.cout; ++d) {
            if ((rc = check_id(pl, op.in1 + d, "upadd")) || (rc = check_id(pl, op.out + d, "upadd")))
              return bail(rc);
            ActRef pv = act_ref(pl->tensors[op.in1 + d]);
            const ActRef ov = act_ref(pl->tensors[op.out + d]);
            if (st.a.h != 2 * pv.h || st.a.w != 2 * pv.w || ov.h != st.a.h || ov.w != st.a.w ||
                st.a.c != ov.c || st.a.c % 8 != 0 || op.in_coff % 8 != 0 || op.in_coff + st.a.c > pv.c ||
                st.a.n != pv.n || st.a.n != ov.n)
              return bail(fail(CERB_ERR_ARG, "op %d: UPADD (group %d) shape mismatch", i, d));
            pv.hi += op.in_coff;
            if (pv.lo) pv.lo += op.in_coff;
            st.up_prev[d] = pv;
            st.up_out[d] = ov;
          }
          break;
        }
        st.a = act_ref(pl->tensors[op.in0]);  // skip
        st.b = act_ref(pl->tensors[op.in1]);  // prev (low res)
        st.c = act_ref(pl->tensors[op.out]);
        // `prev` may be a channel slice of a wider tensor (fused first-stage convs).
        if (st.a.h != 2 * st.b.h || st.a.w != 2 * st.b.w || st.c.h != st.a.h || st.c.w != st.a.w ||
            st.a.c != st.c.c || st.a.c % 8 != 0 || op.in_coff % 8 != 0 ||
            op.in_coff + st.a.c > st.b.c || st.a.n != st.b.n || st.a.n != st.c.n)
          return bail(fail(CERB_ERR_ARG, "op %d: UPADD shape mismatch", i));
        st.b.hi += op.in_coff;
        if (st.b.lo) st.b.lo += op.in_coff;
        break;
      }
      case CERB_OP_HEAD: {
        if ((rc = check_id(pl, op.in0, "head")) || (rc = check_id(pl, op.out, "head")))
          return bail(rc);
        const Tensor& in = pl->tensors[op.in0];
        const Tensor& cv = pl->tensors[op.out];
        HeadParams& h = st.head;
        memset(&h, 0, sizeof(h));
        h.in = act_ref(in);
        if (in.d.dtype != CERB_F16 || in.d.c != 96 || cv.d.dtype != CERB_F32 || cv.d.n != in.d.n ||
            cv.d.h > in.d.h || cv.d.w > in.d.w || op.cout < 2 || op.cout > 8)
          return bail(fail(CERB_ERR_ARG, "op %d: HEAD shape mismatch", i));
        const int width = op.head_mode == CERB_HEAD_INST ? op.cout - 1 : 1;
        if (op.out_coff < 0 || op.out_coff + width > cv.d.c)
          return bail(fail(CERB_ERR_ARG, "op %d: HEAD canvas channels out of range", i));
        if (op.w_off < 0 || op.b_off < 0 ||
            static_cast<size_t>(op.w_off) + op.cout * 96 * 4u > blob_bytes ||
            static_cast<size_t>(op.b_off) + op.cout * 4u > blob_bytes)
          return bail(fail(CERB_ERR_ARG, "op %d: HEAD weight offsets out of range", i));
        h.w = reinterpret_cast<const float*>(pl->blob + op.w_off);
        h.b = reinterpret_cast<const float*>(pl->blob + op.b_off);
        h.classes = op.cout;
        h.mode = op.head_mode;
        h.canvas = static_cast<float*>(cv.plane[0]);
        h.oh = cv.d.h;
        h.ow = cv.d.w;
        h.canvas_c = cv.d.c;
        h.canvas_coff = op.out_coff;
        if (op.logits_out >= 0) {
          if ((rc = check_id(pl, op.logits_out, "head logits"))) return bail(rc);
          const Tensor& lg = pl->tensors[op.logits_out];
          if (lg.d.dtype != CERB_F32 || lg.d.n != in.d.n || lg.d.h != in.d.h ||
              lg.d.w != in.d.w || lg.d.c != op.cout)
            return bail(fail(CERB_ERR_ARG, "op %d: HEAD logits tensor mismatch", i));
          h.logits = static_cast<float*>(lg.plane[0]);
        }
        break;
      }
      case CERB_OP_PCLASS: {
        if ((rc = check_id(pl, op.in0, "pclass")) || (rc = check_id(pl, op.out, "pclass")))
          return bail(rc);
        const Tensor& in = pl->tensors[op.in0];
        const Tensor& cv = pl->tensors[op.out];
        PClassParams& q = st.pclass;
        memset(&q, 0, sizeof(q));
        q.x4 = act_ref(in);
        const size_t need = (512u + 512u + 256u * 512u + 256u + op.cout * 256u + op.cout) * 4u;
        if (in.d.dtype != CERB_F16 || in.d.c != 512 || cv.d.dtype != CERB_F32 ||
            cv.d.n != in.d.n || op.cout < 1 || op.cout > 16 || op.out_coff < 0 ||
            op.out_coff >= cv.d.c || op.w_off < 0 ||
            static_cast<size_t>(op.w_off) + need > blob_bytes)
          return bail(fail(CERB_ERR_ARG, "op %d: PCLASS shape mismatch", i));
        q.params = reinterpret_cast<const float*>(pl->blob + op.w_off);
        q.classes = op.cout;
        q.canvas = static_cast<float*>(cv.plane[0]);
        q.oh = cv.d.h;
        q.ow = cv.d.w;
        q.canvas_c = cv.d.c;
        q.canvas_coff = op.out_coff;
        if (op.logits_out >= 0) {
          if ((rc = check_id(pl, op.logits_out, "pclass logits"))) return bail(rc);
          const Tensor& lg = pl->tensors[op.logits_out];
          if (lg.d.dtype != CERB_F32 ||
              static_cast<size_t>(lg.d.n) * lg.d.h * lg.d.w * lg.d.c !=
                  static_cast<size_t>(in.d.n) * op.cout)
            return bail(fail(CERB_ERR_ARG, "op %d: PCLASS logits tensor mismatch", i));
          q.logits = static_cast<float*>(lg.plane[0]);
        }
        break;
      }
      default:
        return bail(fail(CERB_ERR_ARG, "op %d: unknown kind %d", i, op.kind));
    }
  }
  if ((rc = link_chains(pl, ops, n_ops))) return bail(rc);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return bail(fail(CERB_ERR_CUDA, "plan init: %s", cudaGetErrorString(e)));
  *out = pl;
  return CERB_OK;
}

namespace {
cudaError_t launch_step(cerb_ctx* ctx, Step& st, cudaStream_t s) {
  cudaError_t e = cudaSuccess;
  switch (st.kind) {
    case CERB_OP_PREP:
      e = launch_prep(st.u8_in, st.a.hi, st.a.n, st.a.h, st.a.w - 8, s);
      if (e == cudaSuccess && st.a.lo != nullptr) {
        // integers 0..255 are exact in fp16: the lo plane of the stem input is zero.
        e = cudaMemsetAsync(st.a.lo, 0, static_cast<size_t>(st.a.n) * st.a.h * st.a.w * st.a.c * 2, s);
      }
      break;
    case CERB_OP_CONV:
      if (st.chained) break;  // part of the launch of an earlier step
      e = st.use64x ? conv64x_launch(st.c64x, ctx->conv_sms, s, ctx->use_pdl)
          : st.use64  ? conv64_launch(st.c64, ctx->conv_sms, s, ctx->use_pdl)
          : st.use64s ? conv64s_launch(st.c64s, ctx->conv_sms, s, ctx->use_pdl)
          : st.use3p ? conv3x3c2_launch(st.c3p, ctx->conv_sms, s, ctx->use_pdl)
          : st.use3 ? conv3x3_launch(st.c3, ctx->conv_sms, s, ctx->use_pdl)
                    : conv_tc_launch(st.conv, st.split, ctx->conv_sms, s, ctx->use_pdl);
      break;
    case CERB_OP_MAXPOOL:
      e = launch_maxpool(st.a, st.b, s);
      break;
    case CERB_OP_UPADD:
      if (st.chained) break;  // folded into the next convolution
      e = st.up_groups > 1 ? launch_upadd_multi(st.a, st.up_prev, st.up_out, st.up_groups, s)
                           : launch_upadd(st.a, st.b, st.c, s);
      break;
    case CERB_OP_HEAD:
      e = launch_head(st.head, s);
      break;
    case CERB_OP_PCLASS:
      e = launch_pclass(st.pclass, s);
      break;
  }
  return e;
}
}  // namespace

extern "C" int cerb_plan_run(cerb_plan* pl, const uint8_t* input_u8, int input_on_device) {
  if (!pl) return fail(CERB_ERR_ARG, "cerb_plan_run: null plan");
  cerb_ctx* ctx = pl->ctx;
  CERB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (input_u8 != nullptr) {
    if (pl->prep_in_tensor < 0) return fail(CERB_ERR_ARG, "cerb_plan_run: plan has no PREP op");
    Tensor& t = pl->tensors[pl->prep_in_tensor];
    CERB_CUDA(cudaMemcpyAsync(t.plane[0], input_u8, t.bytes,
                              input_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                              s));
  }
  if (pl->graph_exec != nullptr) {
    CERB_CUDA(cudaGraphLaunch(pl->graph_exec, s));
    ctx->launches += static_cast<int64_t>(pl->n_launches);
    return CERB_OK;
  }
  const bool capture = ctx->use_graphs && pl->runs >= 1;
  if (capture) CERB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  CERB_CUDA(cudaMemsetAsync(pl->tile_counters, 0, sizeof(int) * pl->n_counters, s));
  bool forked = false;
  for (Step& st : pl->steps) {
    cudaStream_t ls = s;
    if (capture && st.side) {
      // graph branch: the op depends on everything queued so far and joins at the end of the plan
      CERB_CUDA(cudaEventRecord(ctx->order_event, s));
      CERB_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->order_event, 0));
      ls = ctx->side_stream;
      forked = true;
    }
    cudaError_t e = launch_step(ctx, st, ls);
    if (e != cudaSuccess) {
      if (capture) {
        cudaGraph_t g = nullptr;
        cudaStreamEndCapture(s, &g);
        if (g) cudaGraphDestroy(g);
      }
      return fail(CERB_ERR_CUDA, "launch of op %d (kind %d%s) failed: %s",
                  static_cast<int>(&st - pl->steps.data()), st.kind,
                  st.use3p ? ", pair kernel" : st.use3 ? ", conv3x3" : st.use64x ? ", conv64x" : "",
                  cudaGetErrorString(e));
    }
    if (!capture && !st.chained) ctx->launches += 1;
  }
  pl->runs += 1;
  if (capture) {
    if (forked) {
      CERB_CUDA(cudaEventRecord(ctx->side_event, ctx->side_stream));
      CERB_CUDA(cudaStreamWaitEvent(s, ctx->side_event, 0));
    }
    cudaGraph_t g = nullptr;
    CERB_CUDA(cudaStreamEndCapture(s, &g));
    cudaError_t e = cudaGraphInstantiate(&pl->graph_exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
      pl->graph_exec = nullptr;
      return fail(CERB_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    CERB_CUDA(cudaGraphLaunch(pl->graph_exec, s));
    ctx->launches += static_cast<int64_t>(pl->n_launches);
  }
  return CERB_OK;
}


extern "C" int cerb_plan_num_ops(cerb_plan* pl) { return pl ? static_cast<int>(pl->steps.size()) : 0; }

extern "C" int cerb_plan_profile(cerb_plan* pl, int reps, float* ms_per_op, int32_t* kinds) {
  if (!pl || reps <= 0 || !ms_per_op) return fail(CERB_ERR_ARG, "cerb_plan_profile: bad arguments");
  cerb_ctx* ctx = pl->ctx;
  CERB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t n = pl->steps.size();
  std::vector<cudaEvent_t> ev((n + 1) * static_cast<size_t>(reps));
  for (cudaEvent_t& e : ev) CERB_CUDA(cudaEventCreate(&e));
  for (int r = 0; r < reps; ++r) {
    cudaEvent_t* e = ev.data() + static_cast<size_t>(r) * (n + 1);
    CERB_CUDA(cudaMemsetAsync(pl->tile_counters, 0, sizeof(int) * pl->n_counters, s));
    CERB_CUDA(cudaEventRecord(e[0], s));
    for (size_t i = 0; i < n; ++i) {
      cudaError_t le = launch_step(ctx, pl->steps[i], s);
      if (le != cudaSuccess)
        return fail(CERB_ERR_CUDA, "profile: launch of op %zu failed: %s", i, cudaGetErrorString(le));
      if (!pl->steps[i].chained) ctx->launches += 1;
      CERB_CUDA(cudaEventRecord(e[i + 1], s));
    }
  }
  int rc = cerb_ctx_sync(ctx);
  if (rc) return rc;
  for (size_t i = 0; i < n; ++i) {
    double acc = 0.0;
    for (int r = 0; r < reps; ++r) {
      cudaEvent_t* e = ev.data() + static_cast<size_t>(r) * (n + 1);
      float ms = 0.f;
      CERB_CUDA(cudaEventElapsedTime(&ms, e[i], e[i + 1]));
      acc += ms;
    }
    ms_per_op[i] = static_cast<float>(acc / reps);
    if (kinds) kinds[i] = pl->steps[i].kind;
  }
  // a layer chain is one launch: its time is spread evenly over its (identically shaped) layers
  for (size_t i = 0; i < n; ++i) {
    const int len = pl->steps[i].chain_len;
    if (len <= 1) continue;
    double tot = 0.0;
    for (int k = 0; k < len; ++k) tot += ms_per_op[i + k];
    for (int k = 0; k < len; ++k) ms_per_op[i + k] = static_cast<float>(tot / len);
  }
  for (cudaEvent_t& e : ev) cudaEventDestroy(e);
  return CERB_OK;
}

extern "C" void* cerb_plan_tensor_ptr(cerb_plan* pl, int tensor_id, int plane) {
  if (!pl || tensor_id < 0 || tensor_id >= static_cast<int>(pl->tensors.size()) || plane < 0 ||
      plane > 1)
    return nullptr;
  return pl->tensors[tensor_id].plane[plane];
}

extern "C" int cerb_plan_read_tensor(cerb_plan* pl, int tensor_id, int plane, void* host_dst,
                                     size_t bytes) {
  if (!pl || !host_dst) return fail(CERB_ERR_ARG, "cerb_plan_read_tensor: bad arguments");
  void* p = cerb_plan_tensor_ptr(pl, tensor_id, plane);
  if (!p) return fail(CERB_ERR_ARG, "cerb_plan_read_tensor: tensor %d plane %d absent", tensor_id, plane);
  if (bytes != pl->tensors[tensor_id].bytes)
    return fail(CERB_ERR_ARG, "cerb_plan_read_tensor: %zu bytes given, tensor holds %zu", bytes,
                pl->tensors[tensor_id].bytes);
  cudaSetDevice(pl->ctx->device);
  cudaError_t e = cudaMemcpyAsync(host_dst, p, bytes, cudaMemcpyDeviceToHost, pl->ctx->stream);
  if (e != cudaSuccess) return fail(CERB_ERR_CUDA, "D2H copy: %s", cudaGetErrorString(e));
  return cerb_ctx_sync(pl->ctx);
}

extern "C" int cerb_plan_write_tensor(cerb_plan* pl, int tensor_id, int plane, const void* host_src,
                                      size_t bytes) {
  if (!pl || !host_src) return fail(CERB_ERR_ARG, "cerb_plan_write_tensor: bad arguments");
  void* p = cerb_plan_tensor_ptr(pl, tensor_id, plane);
  if (!p) return fail(CERB_ERR_ARG, "cerb_plan_write_tensor: tensor %d plane %d absent", tensor_id, plane);
  if (bytes != pl->tensors[tensor_id].bytes)
    return fail(CERB_ERR_ARG, "cerb_plan_write_tensor: %zu bytes given, tensor holds %zu", bytes,
                pl->tensors[tensor_id].bytes);
  cudaSetDevice(pl->ctx->device);
  cudaError_t e = cudaMemcpyAsync(p, host_src, bytes, cudaMemcpyHostToDevice, pl->ctx->stream);
  if (e != cudaSuccess) return fail(CERB_ERR_CUDA, "H2D copy: %s", cudaGetErrorString(e));
  return cerb_ctx_sync(pl->ctx);
}
