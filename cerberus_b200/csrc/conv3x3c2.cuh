// 3x3 stride-1 convolution for the wide layers on CTA PAIRS (tcgen05.mma.cta_group::2), Cin =
// 64*c, Cout = 256*n: see conv3x3c2.cu. fp16 operands, fp32 accumulation (CERB_PREC_F16 only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace cerb {

// per CTA - warp 0: TMA producer, warp 1: MMA issuer (leader CTA) + TMEM allocator, warps 2-5: epilogue
constexpr int kConv3c2Threads = 192;

struct Conv3c2Params {
  CUtensorMap in_map;   // [Cin, W, H, N], box {64, 10, 18, 1}: halo of one 16x8 half region, one 64-channel chunk
  CUtensorMap w_map;    // [9*Cin, Cout], box {64, BN / 2}: the half of a weight slab one CTA holds
  CUtensorMap out_map;  // [Cout, W, H, N], box {64, 8, 16, 1}
  CUtensorMap res_map;  // residual, same geometry as out_map
  int has_res;
  int n_img, H, W;
  int n_chunks;   // Cin / 64
  int BN;         // output channels per work item: 256, 128 or 64 (each CTA holds BN / 2 weight rows)
  int n_ntiles;   // Cout / BN
  int regions_x, regions_y, n_items;
  int n_bstages;  // weight-slab pipeline depth
  const float* bias;  // [Cout] fp32 (BN folded), may be null
  float acc_scale;    // 2^-w_shift
  int relu;
  int* tile_counter;  // zeroed before the launch: dynamic work-item scheduling; null = static split
  int* err_flag;
  long long* prof;
};

void conv3x3c2_plan(Conv3c2Params& p);
size_t conv3x3c2_smem_bytes(const Conv3c2Params& p);
cudaError_t conv3x3c2_launch(const Conv3c2Params& p, int num_sms, cudaStream_t stream, bool pdl = false);

}  // namespace cerb
