// 64->64 3x3 stride-1 convolution in the SPLIT-PRECISION mode (CERB_PREC_F16X2): hi+lo fp16
// operands, three MMAs per product, halo reuse (see conv64s.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace cerb {

// warp 0: halo producer, warp 1: MMA issuer + TMEM allocator, warps 2-5: epilogue, warp 6: lo-weight producer
constexpr int kConv64sThreads = 224;

struct Conv64sParams {
  CUtensorMap in_hi, in_lo;    // [64, W, H, N], box {64, 10, 18, 1}: halo of a 16 (rows) x 8 (columns) tile
  CUtensorMap w_hi, w_lo;      // [576, 64], box {64, 64}: one tap
  CUtensorMap out_hi, out_lo;  // [64, W, H, N], box {64, 8, 16, 1}
  CUtensorMap res_hi, res_lo;  // residual, same geometry
  int has_res;
  int n_img, H, W;
  int tiles_x, tiles_y, n_tiles;
  const float* bias;
  float acc_scale;  // 2^-w_shift
  int relu;
  int* tile_counter;
  int* err_flag;
  long long* prof;
};

void conv64s_plan(Conv64sParams& p);
size_t conv64s_smem_bytes(const Conv64sParams& p);
cudaError_t conv64s_launch(const Conv64sParams& p, int num_sms, cudaStream_t stream, bool pdl = false);

}  // namespace cerb
