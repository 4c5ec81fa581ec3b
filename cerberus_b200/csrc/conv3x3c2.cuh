// 3x3 stride-1 convolution for the wide layers on CTA PAIRS (tcgen05.mma.cta_group::2), Cin =
// 64*c, Cout = 256*n: see conv3x3c2.cu. fp16 operands, fp32 accumulation (CERB_PREC_F16 only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "conv_chain.cuh"

namespace cerb {

// per CTA - warp 0: TMA producer, warp 1: MMA issuer (leader CTA) + TMEM allocator, warps 2-5: epilogue
constexpr int kConv3c2Threads = 192;

using Conv3c2Layer = ConvChainLayer;

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
