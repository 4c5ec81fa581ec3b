// Per-pixel tail of a classification head (models/run_desc.py:451-491): softmax over C logits in
// fp32, INST heads keep channels 1..C-1, TYPE heads keep argmax (first maximum wins, as
// torch.argmax), centre crop (misc/utils.py:94-104) is applied by the caller through `dst`.
#pragma once
#include <cuda_runtime.h>

namespace cerb {

constexpr int kHeadMaxC = 8;

// acc: C logits (bias already added). dst: canvas pixel + channel offset, or nullptr.
__device__ __forceinline__ void head_tail(const float (&acc)[kHeadMaxC], int classes, int mode,
                                          float* __restrict__ logits, float* __restrict__ dst) {
  if (logits != nullptr) {
#pragma unroll
    for (int c = 0; c < kHeadMaxC; ++c)
      if (c < classes) logits[c] = acc[c];
  }
  if (dst == nullptr) return;
  float m = acc[0];
#pragma unroll
  for (int c = 1; c < kHeadMaxC; ++c)
    if (c < classes) m = fmaxf(m, acc[c]);
  float e[kHeadMaxC], sum = 0.0f;
#pragma unroll
  for (int c = 0; c < kHeadMaxC; ++c) {
    e[c] = (c < classes) ? __expf(acc[c] - m) : 0.0f;  // 2 ulp: far inside the 1e-3 gate
    sum += e[c];
  }
  const float inv = __frcp_rn(sum);
  if (mode == 0) {
#pragma unroll
    for (int c = 1; c < kHeadMaxC; ++c)
      if (c < classes) dst[c - 1] = e[c] * inv;
  } else {
    // argmax of the probabilities = argmax of e[] (first maximum wins, as torch.argmax)
    int best = 0;
    float bp = e[0];
#pragma unroll
    for (int c = 1; c < kHeadMaxC; ++c) {
      if (c < classes && e[c] > bp) { bp = e[c]; best = c; }
    }
    dst[0] = static_cast<float>(best);
  }
}

}  // namespace cerb
