// 3x3 stride-1 convolution, Cin = 64*c, Cout = 64 or 128*n: halo reuse + two M tiles per CTA
// (see conv3x3.cu). fp16 operands, fp32 accumulation (CERB_PREC_F16 only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "conv_chain.cuh"

namespace cerb {

// warp 0: TMA producer, warp 1: MMA issuer + TMEM allocator, warps 2-5 / 6-9: epilogue of the
// left / right 16x8-pixel half of the 16x16 region
constexpr int kConv3Threads = 320;

struct Conv3Params {
  // l0 maps: in [Cin, W, H, N], box {64, 18, 18, 1} (halo of a 16x16 region, one 64-channel chunk);
  // w [9*Cin, Cout], box {64, BN}; out / res [Cout, W, H, N], box {64, 8, 16, 1}
  ConvChainLayer l0;             // the layer of a single-layer launch
  const ConvChainLayer* layers;  // chain of n_layers > 1 layers (conv_chain.cuh), table in global memory
  int n_layers;        // 1 = single layer (l0)
  int n_items_layer;   // work items per layer; n_items = n_layers * n_items_layer
  int items_per_img;   // regions_x * regions_y * n_ntiles
  int* done;           // [n_layers * n_img], zeroed before the launch; += 1 per epilogue group and item
  int n_img, H, W;
  int n_chunks;  // Cin / 64
  int BN;        // output channels per work item: 64 or 128
  int n_ntiles;  // Cout / BN
  int regions_x, regions_y, n_items;
  int n_bstages;  // weight-tile pipeline depth
  int rotate;     // 1: each CTA starts its K walk at a different (tap, chunk)
  int* tile_counter;  // zeroed before the launch: dynamic work-item scheduling; null = static split
  int* err_flag;
  long long* prof;
};

void conv3x3_plan(Conv3Params& p);
size_t conv3x3_smem_bytes(const Conv3Params& p);
cudaError_t conv3x3_launch(const Conv3Params& p, int num_sms, cudaStream_t stream, bool pdl = false);

}  // namespace cerb
