// out = skip + bilinear_x2(prev), align_corners=False (models/utils/net_layers.py:45-46,
// models/net_desc.py:185-188): the arithmetic of ONE output pixel x 8 channels, shared by the
// stand-alone pass (ops_misc.cu) and the producer fix-up of the fused 64->64 convolution
// (conv64x.cu) so that both round identically. An output pixel of the 2x2 block below / right of
// low-resolution pixel (j, i) mixes the four neighbours p00 (j, i), p01 (j, i+1), p10 (j+1, i),
// p11 (j+1, i+1) - indices clamped by the caller - with weights ly / lx in {0, 0.25, 0.75}.
#pragma once
#include <cuda_fp16.h>
#include <cstdint>

namespace cerb {

__device__ __forceinline__ void upadd_cvt8(const uint4& u, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __half22float2(h[e]);
    v[2 * e] = f.x;
    v[2 * e + 1] = f.y;
  }
}

__device__ __forceinline__ uint4 upadd_pixel8(const float (&p00)[8], const float (&p01)[8],
                                              const float (&p10)[8], const float (&p11)[8],
                                              const float (&s)[8], float ly, float lx) {
  const float hy = 1.0f - ly, hx = 1.0f - lx;
  uint4 o;
  uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float u0 = hy * (hx * p00[2 * e] + lx * p01[2 * e]) + ly * (hx * p10[2 * e] + lx * p11[2 * e]);
    const float u1 = hy * (hx * p00[2 * e + 1] + lx * p01[2 * e + 1]) +
                     ly * (hx * p10[2 * e + 1] + lx * p11[2 * e + 1]);
    const __half2 h = __floats2half2_rn(s[2 * e] + u0, s[2 * e + 1] + u1);
    ow[e] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return o;
}

// fp16 throughput mode: the same expression in packed half arithmetic (inputs and result are fp16
// anyway; four extra roundings of 2^-11 on the interpolated term). A third of the instructions of
// the fp32 form - the fix-up warps of conv64x.cu are issue-bound - and no conversions.
__device__ __forceinline__ uint4 upadd_pixel8_h2(const uint4& p00, const uint4& p01, const uint4& p10,
                                                 const uint4& p11, const uint4& s, float ly, float lx) {
  const __half2 ly2 = __float2half2_rn(ly), hy2 = __float2half2_rn(1.0f - ly);
  const __half2 lx2 = __float2half2_rn(lx), hx2 = __float2half2_rn(1.0f - lx);
  const __half2* a00 = reinterpret_cast<const __half2*>(&p00);
  const __half2* a01 = reinterpret_cast<const __half2*>(&p01);
  const __half2* a10 = reinterpret_cast<const __half2*>(&p10);
  const __half2* a11 = reinterpret_cast<const __half2*>(&p11);
  const __half2* sv = reinterpret_cast<const __half2*>(&s);
  uint4 o;
  __half2* ov = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __half2 t0 = __hfma2(lx2, a01[e], __hmul2(hx2, a00[e]));
    const __half2 t1 = __hfma2(lx2, a11[e], __hmul2(hx2, a10[e]));
    ov[e] = __hadd2(sv[e], __hfma2(ly2, t1, __hmul2(hy2, t0)));
  }
  return o;
}

}  // namespace cerb
