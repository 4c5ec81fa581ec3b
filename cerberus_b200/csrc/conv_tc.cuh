// Implicit-GEMM NHWC convolution on tcgen05 tensor cores (sm_100a).
//
// GEMM view: M = output pixels, N = Cout, K = taps * Cin.
// One CTA tile = 128 output pixels (a BH x BW spatial box of one image) x BN output
// channels. For every filter tap the 128 x 64-channel activation slab is fetched by ONE
// 4-D TMA box load whose (x, y) coordinates are the tile origin shifted by the tap
// offset; TMA zero-fills out-of-bounds pixels, which is exactly the convolution's zero
// padding, so no im2col buffer ever exists. Stride-2 layers read four parity views of
// the input (one tensor map per (y&1, x&1)), the 7x7 stem reads a pre-padded 8-channel
// image through an overlapping-window tensor map (7 row taps x 64 = 8 px * 8 ch).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace cerb {

constexpr int kConvMaxTaps = 16;
// warp 0: TMA, warp 1: MMA issue + TMEM alloc, warps 2-5 / 6-9: epilogue groups 0 / 1
constexpr int kConvThreads = 576;  // launch bound; the launch uses 64 + 128 * n_acc threads

struct ConvTap {
  int8_t map;  // which input tensor map (parity view) this tap reads
  int8_t dx;   // x offset added to the tile origin, in that view's pixel grid
  int8_t dy;
  int8_t pad;
};

struct ConvKParams {
  CUtensorMap in_hi[4];
  CUtensorMap in_lo[4];  // only read in split-precision mode
  CUtensorMap w_hi;
  CUtensorMap w_lo;
  ConvTap taps[kConvMaxTaps];
  int n_taps;
  int n_chunks;  // Cin / 64 per tap
  // output geometry
  int n_img, H, W;
  int bw_log2;  // tile box width = 1 << bw_log2, height = 128 >> bw_log2
  int tiles_x, tiles_y, n_ntiles, n_tiles;
  int BN;  // output channels per tile (multiple of 32, <= 256)
  // epilogue
  const float* bias;  // [Cout] fp32 (BN folded), may be null
  float acc_scale;    // 2^-w_shift: weights are stored pre-scaled by a power of two
  __half* out_hi;
  __half* out_lo;  // split-precision mode only
  const __half* res_hi;
  const __half* res_lo;
  int out_cs;    // channel stride (elements per pixel) of the output tensor
  int out_coff;  // channel offset inside the output tensor
  int res_cs;
  int res_coff;
  int relu;
  // fused classification-head tail (head_classes > 0): see CERB_OP_CONV.aux_classes
  const float* head_w;   // [C][96] fp32
  const float* head_b;   // [C]
  int head_classes, head_mode;
  // MMA-tail head: the hidden layer's 96 biases and the C head biases travel in the kernel
  // parameters (constant bank): the epilogue's FMAs read them as constant operands instead of
  // 24 broadcast LDS.128 per pixel, which cost 38 % of the kernel's shared-memory wavefronts
  float head_hbias[96];
  float head_obias[8];
  float* canvas;         // [N, oh, ow, canvas_c]
  float* logits;         // optional [N, H, W, C]
  int oh, ow, canvas_c, canvas_coff;
  // pipeline
  int n_acc;  // accumulator stages = epilogue groups (2 or 4; split mode: 1, or 2 when BN <= 64)
  int n_stages;
  int stage_bytes;
  int a_lo_zero;  // split mode: the lo plane of the input is all zeros (stem: uint8 pixels are exact
                  // in fp16) - its loads and the lo x hi products are skipped
  int mma_tail;  // fused head: 1x1 96 -> C on the tensor core (fp16 mode), else fp32 FMAs
  int* tile_counter;  // zeroed before the launch: dynamic tile scheduling; null = static split
  int* err_flag;
  long long* prof;  // optional [grid][16] per-role cycle counters (option "kernel_prof")
};

// Host side: encodes nothing, just launches. `split` selects the 3-MMA hi/lo mode.
cudaError_t conv_tc_launch(const ConvKParams& p, bool split, int num_sms, cudaStream_t stream,
                           bool pdl = false);
size_t conv_tc_smem_bytes(const ConvKParams& p);
// Fills n_stages / stage_bytes for a given BN and mode.
void conv_tc_plan_pipeline(ConvKParams& p, bool split);

}  // namespace cerb
