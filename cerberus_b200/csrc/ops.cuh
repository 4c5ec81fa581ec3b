// Launchers for the bandwidth-bound operators around the convolution kernel.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace cerb {

// fp16 NHWC activation, optionally carried as a hi+lo pair (F16X2 mode; lo == nullptr otherwise).
struct ActRef {
  __half* hi;
  __half* lo;
  int n, h, w, c;
};

// u8 [N,H,W,3] -> fp16 [N,H,W+8,8]: pixel x lands at column x+3, channels 3..7 and the
// 3 (left) + 5 (right) border columns are zero. Values stay integers 0..255 (exact in fp16);
// the reference's /255 (models/net_desc.py:147) is folded into the stem weights.
cudaError_t launch_prep(const uint8_t* in_u8, __half* out, int n, int h, int w, cudaStream_t s);

// 3x3 stride-2 pad-1 max pool with -inf padding (models/backbone/resnet.py:201).
cudaError_t launch_maxpool(ActRef in, ActRef out, cudaStream_t s);

// out = skip + bilinear_x2(prev), align_corners=False (models/utils/net_layers.py:45-46,
// models/net_desc.py:185-188). skip/out: [N,2h,2w,C], prev: [N,h,w,C].
cudaError_t launch_upadd(ActRef skip, ActRef prev, ActRef out, cudaStream_t s);
// n_groups (prev[d], out[d]) pairs against ONE skip tensor, read once (fp16 mode; else one pass each)
cudaError_t launch_upadd_multi(ActRef skip, const ActRef* prev, const ActRef* out, int n_groups,
                               cudaStream_t s);

struct HeadParams {
  ActRef in;            // [N,H,W,96] post-ReLU hidden layer of the classification head
  const float* w;       // [C][96] fp32
  const float* b;       // [C]
  int classes;          // C <= 8
  int mode;             // cerb_head_mode
  float* logits;        // optional [N,H,W,C] fp32
  float* canvas;        // [N,oh,ow,canvas_c] fp32 patch canvas, or nullptr
  int oh, ow, canvas_c, canvas_coff;
};
// 1x1 conv 96->C + bias, then softmax / argmax / centre crop (models/utils/net_layers.py:36-38,
// models/run_desc.py:451-491, misc/utils.py:94-104).
cudaError_t launch_head(const HeadParams& p, cudaStream_t s);

struct PClassParams {
  ActRef x4;            // [N,h4,w4,512] encoder bottom features
  const float* params;  // bn1 scale[512], bn1 shift[512], W1^T[512][256], b1[256], W2[C][256], b2[C]
  int classes;
  float* logits;        // optional [N,C]
  float* canvas;        // [N,oh,ow,canvas_c]
  int oh, ow, canvas_c, canvas_coff;
};
// Patch-Class branch: centre-crop 9x9 (iff h4 != 9 and w4 != 9), global average pool,
// BN+ReLU, 1x1 512->256 (+folded BN) + ReLU, 1x1 256->C, argmax broadcast over the output
// window (models/net_desc.py:64-76,169-180; models/run_desc.py:459-461,479-486).
cudaError_t launch_pclass(const PClassParams& p, cudaStream_t s);

}  // namespace cerb
