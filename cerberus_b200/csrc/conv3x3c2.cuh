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

// What differs between the layers of a chain (see Conv3c2Params::layers).
struct Conv3c2Layer {
  CUtensorMap in_map;   // [Cin, W, H, N], box {64, 10, 18, 1}: halo of one 16x8 half region, one 64-channel chunk
  CUtensorMap w_map;    // [9*Cin, Cout], box {64, BN / 2}: the half of a weight slab one CTA holds
  CUtensorMap out_map;  // [Cout, W, H, N], box {64, 8, 16, 1}
  CUtensorMap res_map;  // residual, same geometry as out_map
  const float* bias;    // [Cout] fp32 (BN folded), may be null
  float acc_scale;      // 2^-w_shift
  int has_res;
  int relu;
  int pad_[11];
};
static_assert(sizeof(Conv3c2Layer) % 64 == 0, "tensor maps of a layer table must stay 64-byte aligned");

struct Conv3c2Params {
  Conv3c2Layer l0;  // the layer of a single-layer launch (kernel parameter space)
  // A CHAIN of n_layers > 1 layers of identical geometry, each reading the output of the one
  // before (encoder layer3 / layer4 bodies): ONE persistent launch whose work items are numbered
  // layer-major and drawn in that order from the global counter. An item of layer l on image i
  // starts when all items of layer l - 1 on image i have been stored (`done` counters, release /
  // acquire through global memory), so the launch, pipeline-fill, tail and wave-quantisation cost
  // of a launch is paid once per chain instead of once per layer. Table in global memory.
  const Conv3c2Layer* layers;
  int n_layers;        // 1 = single layer (l0)
  int n_items_layer;   // work items per layer; n_items = n_layers * n_items_layer
  int items_per_img;   // regions_x * regions_y * n_ntiles
  int* done;           // [n_layers * n_img], zeroed before the launch; += 1 per CTA and item
  int n_img, H, W;
  int n_chunks;   // Cin / 64
  int BN;         // output channels per work item: 256, 128 or 64 (each CTA holds BN / 2 weight rows)
  int n_ntiles;   // Cout / BN
  int regions_x, regions_y, n_items;
  int n_bstages;  // weight-slab pipeline depth
  int* tile_counter;  // zeroed before the launch: dynamic work-item scheduling; null = static split
  int* err_flag;
  long long* prof;
};

void conv3x3c2_plan(Conv3c2Params& p);
size_t conv3x3c2_smem_bytes(const Conv3c2Params& p);
cudaError_t conv3x3c2_launch(const Conv3c2Params& p, int num_sms, cudaStream_t stream, bool pdl = false);

}  // namespace cerb
