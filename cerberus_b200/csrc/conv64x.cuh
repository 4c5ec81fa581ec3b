// 64->64 3x3 stride-1 convolution, "even/odd" formulation (see conv64x.cu): N = 128 MMAs that
// compute two horizontally adjacent output pixels per TMEM lane, lifting the N = 64
// shared-memory-bandwidth ceiling of conv64.cu. fp16 operands, fp32 accumulation.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace cerb {

// warp 0: TMA producer, warp 1: MMA issuer + TMEM allocator, warps 2-5 / 6-9: epilogue of the
// even / odd pixel of every lane
constexpr int kConv64xThreads = 320;
// fused upsample+add (Conv64xParams::fuse_up): warps 10-15 add bilinear_x2(low) to the skip halo
// in shared memory before the MMA warp may read it
constexpr int kConv64xFuseThreads = 512;

struct Conv64xParams {
  CUtensorMap in_map;   // [64 ch, W, H, N], box {64, 18 (element stride 2 -> 9 columns), 18, 1}
  CUtensorMap w_map;    // [576, 64], box {64, 64}: one tap
  CUtensorMap out_map;  // [64 ch, W, H, N], box {64, 16 (element stride 2 -> 8 columns), 16, 1}: TMA store
                        // of the even-pixel / odd-pixel plane of a region
  CUtensorMap res_map;  // residual, same geometry
  int has_res;
  // fuse_up: the input of the convolution is skip + bilinear_x2(low), align_corners=False
  // (models/net_desc.py:185-188): in_map addresses the SKIP tensor, low_map the half-resolution
  // tensor [64 ch, W/2, H/2, N], box {64, 10, 10, 1} = the low pixels under an 18x18 halo
  CUtensorMap low_map;
  int fuse_up;
  int n_img, H, W;
  int regions_x, regions_y, n_regions;
  const float* bias;
  float acc_scale;
  int relu;
  int n_stages;
  int* tile_counter;  // zeroed before the launch: dynamic region scheduling; null = static split
  int* err_flag;
  long long* prof;
};

void conv64x_plan(Conv64xParams& p);
size_t conv64x_smem_bytes(const Conv64xParams& p);
cudaError_t conv64x_launch(const Conv64xParams& p, int num_sms, cudaStream_t stream, bool pdl = false);

}  // namespace cerb
