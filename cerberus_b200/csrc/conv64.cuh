// 64->64 3x3 stride-1 convolution with resident weights and halo reuse (see conv64.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace cerb {

// warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5 (+ 6-9): epilogue groups; with the fused
// upsample+add input, warps 6-13 are its producers and only warps 2-5 run the epilogue
constexpr int kConv64Threads = 320;
constexpr int kConv64ThreadsUp = 448;

struct Conv64Params {
  CUtensorMap in_map;  // [64 ch, W, H, N], box {64, conv64_box_w(mode), 18, 1}
  CUtensorMap w_map;   // [576, 64], box {64, 64}
  CUtensorMap out_map; // output [64 ch, W, H, N], box {64, 8, 16, 1} (TMA store)
  CUtensorMap res_map; // residual, same geometry (TMA load into the staging tile)
  int has_res;
  // fused classification head (has_tail): hidden 1x1 64 -> 96 + BN + ReLU, 1x1 96 -> C, softmax /
  // argmax / centre crop into the canvas; the conv's own 64-channel output is not stored
  CUtensorMap w1_map;  // [64, 96] fp16 hidden weights, box {64, 96}
  int has_tail;
  const float* tail_b1;  // [96]
  float tail_scale;      // 2^-w_shift of the hidden layer
  const float* tail_w2;  // [C][96] fp32
  const float* tail_b2;  // [C]
  int tail_classes, tail_mode;
  float* canvas;         // [N, oh, ow, canvas_c]
  float* logits;         // optional [N, H, W, C]
  int oh, ow, canvas_c, canvas_coff;
  int mode;            // halo layout, see conv64.cu (4 = 7x7 stem over the PREP tensor)
  int n_taps;          // 9, or 7 for the stem
  int halo_x, halo_y;  // halo origin = tile origin - (halo_x, halo_y): (1, 1), stem (0, 3)
  int n_img, H, W;
  int tiles_x, tiles_y, n_tiles;
  const float* bias;
  float acc_scale;
  int relu;
  // fused input: A = up_skip + bilinear_x2(up_prev) instead of a TMA load of in_map (mode 1 only)
  const __half* up_skip;
  const __half* up_prev;
  int up_skip_cs, up_prev_cs;
  int debug;  // attribution experiments: 1 = no global stores, 4 = no epilogue math
  // filled by conv64_plan
  int pitch_px, copy_bytes, stage_bytes, tx_bytes, sbo_bytes, n_stages;
  int* err_flag;
  long long* prof;  // optional [grid][16] per-role wait cycles (cerb_ctx_set_option "kernel_prof")
};

void conv64_plan(Conv64Params& p);
int conv64_box_w(int mode);
int conv64_tile_w();
int conv64_tile_h();
size_t conv64_smem_bytes(const Conv64Params& p);
// pdl: launch with programmatic stream serialisation (the kernel's prologue overlaps the tail of
// the previous kernel in the stream; see ptx::grid_dep_wait)
cudaError_t conv64_launch(const Conv64Params& p, int num_sms, cudaStream_t stream, bool pdl = false);

}  // namespace cerb
