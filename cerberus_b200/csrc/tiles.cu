// Tile plumbing on the device: reflect-padded patch extraction (infer/tile.py:64-69 +
// loader/infer_loader.py:57-69) and canvas stitching (infer/tile.py:136-163). Byte / fp32
// copy kernels, HBM-bound; vectorised where the layout allows.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "capi_internal.cuh"

using namespace cerb;

namespace {

// numpy.pad(mode="reflect") index map: period 2(n-1), no edge repetition; n == 1 -> 0.
__device__ __forceinline__ int reflect_index(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  int m = i % period;
  if (m < 0) m += period;
  return m < n ? m : period - m;
}

__global__ void k_extract(const uint8_t* __restrict__ img, int H, int W, int pad_t, int pad_l,
                          const int* __restrict__ tl, int ph, int pw, uint8_t* __restrict__ out,
                          int zero_pad) {
  const int patch = blockIdx.y;
  const int ty = tl[2 * patch], tx = tl[2 * patch + 1];
  const int total = ph * pw;
  uint8_t* o = out + static_cast<size_t>(patch) * total * 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / pw, x = i % pw;
    int sy = ty + y - pad_t, sx = tx + x - pad_l;
    if (zero_pad) {  // WSI mode: read_bounds(..., pad_constant_values=0) (infer/wsi.py:936-942)
      if (sy < 0 || sy >= H || sx < 0 || sx >= W) {
        o[3 * i + 0] = 0; o[3 * i + 1] = 0; o[3 * i + 2] = 0;
        continue;
      }
    } else {
      sy = reflect_index(sy, H);
      sx = reflect_index(sx, W);
    }
    const uint8_t* s = img + (static_cast<size_t>(sy) * W + sx) * 3;
    o[3 * i + 0] = s[0];
    o[3 * i + 1] = s[1];
    o[3 * i + 2] = s[2];
  }
}

// One thread per output (pixel, channel): gathers the covering patches in list order so the
// fp32 summation order equals the reference's sequential `+=`.
__global__ void k_stitch(const float* __restrict__ patches, int n, int oh, int ow, int C,
                         const int* __restrict__ tl, int src_y, int src_x, int out_h, int out_w,
                         float* __restrict__ out) {
  const size_t total = static_cast<size_t>(out_h) * out_w * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t pix = i / C;
    const int x = static_cast<int>(pix % out_w) + src_x;
    const int y = static_cast<int>(pix / out_w) + src_y;
    float sum = 0.0f, cnt = 0.0f;
    for (int p = 0; p < n; ++p) {
      const int py = y - tl[2 * p], px = x - tl[2 * p + 1];
      if (py < 0 || py >= oh || px < 0 || px >= ow) continue;
      sum += patches[((static_cast<size_t>(p) * oh + py) * ow + px) * C + c];
      cnt += 1.0f;
    }
    out[i] = sum / (cnt + 1.0e-8f);
  }
}


// WSI canvas assembly (infer/wsi.py:463,615 -> tiatoolbox merge_prediction): with the reference's
// stride == patch_output_shape every canvas pixel receives exactly one patch, so the running
// average (old*cnt + new)/(cnt+1) over a zero-initialised canvas is a plain clipped write.
__global__ void k_scatter(const float* __restrict__ patches, int oh, int ow, int C,
                          const int* __restrict__ tl, float* __restrict__ canvas, int H, int W) {
  const int patch = blockIdx.y;
  const int ty = tl[2 * patch], tx = tl[2 * patch + 1];
  const int row_elems = ow * C;
  const int total = oh * row_elems;
  const float* src = patches + static_cast<size_t>(patch) * total;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / row_elems, r = i - y * row_elems;
    const int x = r / C;
    const int Y = ty + y, X = tx + x;
    if (Y < 0 || Y >= H || X < 0 || X >= W) continue;
    canvas[(static_cast<size_t>(Y) * W + X) * C + (r - x * C)] = src[i];
  }
}

// infer/wsi.py:763-788: tile_pred_map *= mask; cv2.resize(tile_pred_map, (0,0), fx=0.5, fy=0.5)
// (default INTER_LINEAR, float32). What OpenCV computes for an exact factor of two depends on the
// channel count (established against cv2 4.13 + IPP 2022.2 of this image, tests/test_gpu_wsi.py;
// the reference's pinned opencv-python wheel ships IPP as well):
//   * 1, 3, 4 channels -> IPP bilinear in lerp form, horizontal first, no FMA:
//         t = a + (b - a) * wx ;  u = c + (d - c) * wx ;  out = t + (u - t) * wy
//     with taps (2d, 2d+1) and weight 0.5, or taps (n-2, n-1) and weight 1 where 2d+1 leaves the
//     image (odd sizes);
//   * 2 channels (no IPP variant) -> cv::resize turns LINEAR with scale 2 into its INTER_AREA
//     fast path: (((a + b) + c) + d) * 0.25 over the 2x2 block in raster order, and sum / count
//     over the pixels that exist for the last row / column of odd sizes.
// Every operation below is a single-rounding intrinsic so the compiler cannot contract them.
__global__ void k_region_half(const float* __restrict__ canvas, int W, int C, int y0, int x0, int h,
                              int w, const uint8_t* __restrict__ mask, const int* __restrict__ chans,
                              int k, float* __restrict__ out, int oh, int ow) {
  const size_t total = static_cast<size_t>(oh) * ow * k;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % k);
    const size_t pix = i / k;
    const int dx = static_cast<int>(pix % ow), dy = static_cast<int>(pix / ow);
    const int ch = chans[c];
    auto at = [&](int yy, int xx) -> float {
      const float v = canvas[(static_cast<size_t>(y0 + yy) * W + (x0 + xx)) * C + ch];
      return mask != nullptr ? __fmul_rn(v, static_cast<float>(mask[static_cast<size_t>(yy) * w + xx])) : v;
    };
    const int sx = 2 * dx, sy = 2 * dy;
    if (k == 2) {
      if (sx + 2 <= w && sy + 2 <= h) {
        const float s4 = __fadd_rn(__fadd_rn(__fadd_rn(at(sy, sx), at(sy, sx + 1)), at(sy + 1, sx)),
                                   at(sy + 1, sx + 1));
        out[i] = __fmul_rn(s4, 0.25f);
      } else {
        float sum = 0.0f;
        int cnt = 0;
        for (int yy = sy; yy < sy + 2 && yy < h; ++yy)
          for (int xx = sx; xx < sx + 2 && xx < w; ++xx) {
            sum = __fadd_rn(sum, at(yy, xx));
            ++cnt;
          }
        out[i] = cnt > 0 ? __fdiv_rn(sum, static_cast<float>(cnt)) : 0.0f;
      }
      continue;
    }
    int xa = sx, ya = sy;
    float wx = 0.5f, wy = 0.5f;
    if (sx + 1 > w - 1) { xa = w >= 2 ? w - 2 : 0; wx = 1.0f; }
    if (sy + 1 > h - 1) { ya = h >= 2 ? h - 2 : 0; wy = 1.0f; }
    const int xb = w >= 2 ? xa + 1 : 0, yb = h >= 2 ? ya + 1 : 0;
    const float a = at(ya, xa), b = at(ya, xb), cc = at(yb, xa), d = at(yb, xb);
    const float t = __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), wx));
    const float u = __fadd_rn(cc, __fmul_rn(__fsub_rn(d, cc), wx));
    out[i] = __fadd_rn(t, __fmul_rn(__fsub_rn(u, t), wy));
  }
}

// cv2.resize(..., fx=f, fy=f, interpolation=INTER_NEAREST) of one canvas channel
// (infer/wsi.py:694-702, Patch-Class map at 0.25): src index = min(floor(d / f), n - 1).
__global__ void k_nearest_channel(const float* __restrict__ canvas, int H, int W, int C, int ch,
                                  float inv_scale, float* __restrict__ out, int oh, int ow) {
  const size_t total = static_cast<size_t>(oh) * ow;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int dx = static_cast<int>(i % ow), dy = static_cast<int>(i / ow);
    int sx = static_cast<int>(floorf(dx * inv_scale)), sy = static_cast<int>(floorf(dy * inv_scale));
    if (sx > W - 1) sx = W - 1;
    if (sy > H - 1) sy = H - 1;
    out[i] = canvas[(static_cast<size_t>(sy) * W + sx) * C + ch];
  }
}

}  // namespace

extern "C" void* cerb_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    fail(CERB_ERR_CUDA, "cerb_host_alloc(%zu) failed", bytes);
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

extern "C" void cerb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

extern "C" void* cerb_dev_alloc(cerb_ctx* ctx, size_t bytes) {
  if (!ctx || bytes == 0) return nullptr;
  cudaSetDevice(ctx->device);
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    fail(CERB_ERR_CUDA, "cerb_dev_alloc(%zu) failed", bytes);
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

extern "C" int cerb_dev_free(cerb_ctx* ctx, void* p) {
  if (!ctx) return fail(CERB_ERR_ARG, "cerb_dev_free: null ctx");
  cudaSetDevice(ctx->device);
  CERB_CUDA(cudaStreamSynchronize(ctx->stream));
  CERB_CUDA(cudaFree(p));
  return CERB_OK;
}

extern "C" int cerb_memcpy(cerb_ctx* ctx, void* dst, const void* src, size_t bytes, int kind) {
  if (!ctx || !dst || !src) return fail(CERB_ERR_ARG, "cerb_memcpy: bad arguments");
  cudaSetDevice(ctx->device);
  const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice
                           : kind == 2 ? cudaMemcpyDeviceToHost
                                       : cudaMemcpyDeviceToDevice;
  CERB_CUDA(cudaMemcpyAsync(dst, src, bytes, k, ctx->stream));
  if (kind == 2) return cerb_ctx_sync(ctx);
  return CERB_OK;
}

namespace {
struct DevBuf {
  void* p = nullptr;
  cerb_ctx* ctx;
  explicit DevBuf(cerb_ctx* c) : ctx(c) {}
  ~DevBuf() {
    if (p) {
      cudaStreamSynchronize(ctx->stream);
      cudaFree(p);
    }
  }
};
}  // namespace

namespace {
// Copies a small host array into the next slot of the ctx's parameter ring and returns its device
// address; the copy is queued on the ctx stream and the slot is reused 64 calls later (its event
// is waited for then). *dev = nullptr when the array does not fit a slot.
int stage_params(cerb_ctx* ctx, const void* host, size_t bytes, const void** dev) {
  *dev = nullptr;
  if (bytes > cerb_ctx::kParamBytes) return CERB_OK;
  if (!ctx->param_host) {
    CERB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&ctx->param_host),
                             cerb_ctx::kParamSlots * cerb_ctx::kParamBytes));
    CERB_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->param_dev),
                         cerb_ctx::kParamSlots * cerb_ctx::kParamBytes));
    for (int i = 0; i < cerb_ctx::kParamSlots; ++i)
      CERB_CUDA(cudaEventCreateWithFlags(&ctx->param_event[i], cudaEventDisableTiming));
  }
  const int slot = ctx->param_next;
  ctx->param_next = (slot + 1) % cerb_ctx::kParamSlots;
  CERB_CUDA(cudaEventSynchronize(ctx->param_event[slot]));  // never recorded: returns at once
  char* h = ctx->param_host + slot * cerb_ctx::kParamBytes;
  char* d = ctx->param_dev + slot * cerb_ctx::kParamBytes;
  memcpy(h, host, bytes);
  CERB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CERB_CUDA(cudaEventRecord(ctx->param_event[slot], ctx->stream));
  *dev = d;
  return CERB_OK;
}
}  // namespace

extern "C" int cerb_extract_patches(cerb_ctx* ctx, const uint8_t* img, int H, int W, int pad_t,
                                    int pad_l, const int32_t* tl_yx, int n, int ph, int pw,
                                    uint8_t* out, int flags) {
  if (!ctx || !img || !tl_yx || !out || H <= 0 || W <= 0 || n <= 0 || ph <= 0 || pw <= 0)
    return fail(CERB_ERR_ARG, "cerb_extract_patches: bad arguments");
  CERB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevBuf dimg(ctx), dtl(ctx), dout(ctx);
  const size_t img_bytes = static_cast<size_t>(H) * W * 3;
  const size_t out_bytes = static_cast<size_t>(n) * ph * pw * 3;
  const uint8_t* src = img;
  if (!(flags & 1)) {
    CERB_CUDA(cudaMalloc(&dimg.p, img_bytes));
    CERB_CUDA(cudaMemcpyAsync(dimg.p, img, img_bytes, cudaMemcpyHostToDevice, s));
    src = static_cast<const uint8_t*>(dimg.p);
  }
  // image and patches both resident: nothing of the caller's is read after the call returns
  // (the top-left table goes through the parameter ring), so the call stays asynchronous
  const bool resident = (flags & 1) && (flags & 2);
  const void* tl_dev = nullptr;
  if (resident) {
    const int rc = stage_params(ctx, tl_yx, sizeof(int) * 2 * n, &tl_dev);
    if (rc) return rc;
  }
  if (!tl_dev) {
    CERB_CUDA(cudaMalloc(&dtl.p, sizeof(int) * 2 * n));
    CERB_CUDA(cudaMemcpyAsync(dtl.p, tl_yx, sizeof(int) * 2 * n, cudaMemcpyHostToDevice, s));
    tl_dev = dtl.p;
  }
  uint8_t* dst = out;
  if (!(flags & 2)) {
    CERB_CUDA(cudaMalloc(&dout.p, out_bytes));
    dst = static_cast<uint8_t*>(dout.p);
  }
  int gx = (ph * pw + 255) / 256;
  if (gx > 64) gx = 64;
  k_extract<<<dim3(gx, n), 256, 0, s>>>(src, H, W, pad_t, pad_l, static_cast<const int*>(tl_dev), ph,
                                        pw, dst, (flags & 4) ? 1 : 0);
  CERB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  if (!(flags & 2)) {
    CERB_CUDA(cudaMemcpyAsync(out, dst, out_bytes, cudaMemcpyDeviceToHost, s));
  }
  if (resident && tl_dev != dtl.p) return CERB_OK;
  return cerb_ctx_sync(ctx);  // tl_yx / temporaries are released on return
}

extern "C" int cerb_stitch(cerb_ctx* ctx, const float* patches, int n, int oh, int ow, int C,
                           const int32_t* tl_yx, int canvas_h, int canvas_w, int src_y, int src_x,
                           int out_h, int out_w, float* out, int flags) {
  if (!ctx || !patches || !tl_yx || !out || n <= 0 || oh <= 0 || ow <= 0 || C <= 0 || out_h <= 0 ||
      out_w <= 0 || src_y < 0 || src_x < 0 || src_y + out_h > canvas_h || src_x + out_w > canvas_w)
    return fail(CERB_ERR_ARG, "cerb_stitch: bad arguments");
  CERB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevBuf dp(ctx), dtl(ctx), dout(ctx);
  const size_t p_bytes = static_cast<size_t>(n) * oh * ow * C * sizeof(float);
  const size_t o_bytes = static_cast<size_t>(out_h) * out_w * C * sizeof(float);
  const float* src = patches;
  if (!(flags & 1)) {
    CERB_CUDA(cudaMalloc(&dp.p, p_bytes));
    CERB_CUDA(cudaMemcpyAsync(dp.p, patches, p_bytes, cudaMemcpyHostToDevice, s));
    src = static_cast<const float*>(dp.p);
  }
  CERB_CUDA(cudaMalloc(&dtl.p, sizeof(int) * 2 * n));
  CERB_CUDA(cudaMemcpyAsync(dtl.p, tl_yx, sizeof(int) * 2 * n, cudaMemcpyHostToDevice, s));
  float* dst = out;
  if (!(flags & 2)) {
    CERB_CUDA(cudaMalloc(&dout.p, o_bytes));
    dst = static_cast<float*>(dout.p);
  }
  k_stitch<<<148 * 8, 256, 0, s>>>(src, n, oh, ow, C, static_cast<const int*>(dtl.p), src_y, src_x,
                                   out_h, out_w, dst);
  CERB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  if (!(flags & 2)) CERB_CUDA(cudaMemcpyAsync(out, dst, o_bytes, cudaMemcpyDeviceToHost, s));
  return cerb_ctx_sync(ctx);
}

// ------------------------------------------------------------------ WSI plumbing
extern "C" int cerb_scatter_patches(cerb_ctx* ctx, const float* patches_dev, int n, int oh, int ow,
                                    int C, const int32_t* tl_yx, float* canvas_dev, int H, int W) {
  if (!ctx || !patches_dev || !tl_yx || !canvas_dev || n <= 0 || oh <= 0 || ow <= 0 || C <= 0 ||
      H <= 0 || W <= 0)
    return fail(CERB_ERR_ARG, "cerb_scatter_patches: bad arguments");
  CERB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevBuf dtl(ctx);
  const void* tl_dev = nullptr;
  const int rc = stage_params(ctx, tl_yx, sizeof(int) * 2 * n, &tl_dev);
  if (rc) return rc;
  if (!tl_dev) {
    CERB_CUDA(cudaMalloc(&dtl.p, sizeof(int) * 2 * n));
    CERB_CUDA(cudaMemcpyAsync(dtl.p, tl_yx, sizeof(int) * 2 * n, cudaMemcpyHostToDevice, s));
    tl_dev = dtl.p;
  }
  int gx = (oh * ow * C + 255) / 256;
  if (gx > 64) gx = 64;
  k_scatter<<<dim3(gx, n), 256, 0, s>>>(patches_dev, oh, ow, C, static_cast<const int*>(tl_dev),
                                        canvas_dev, H, W);
  CERB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  if (tl_dev != dtl.p) return CERB_OK;  // asynchronous: everything involved is device memory
  return cerb_ctx_sync(ctx);
}

extern "C" int cerb_crop2d(cerb_ctx* ctx, const void* src_dev, int H, int W, int px_bytes, int y0,
                           int x0, int h, int w, void* dst, int flags) {
  if (!ctx || !src_dev || !dst || px_bytes <= 0 || y0 < 0 || x0 < 0 || h <= 0 || w <= 0 ||
      y0 + h > H || x0 + w > W)
    return fail(CERB_ERR_ARG, "cerb_crop2d: bad arguments");
  CERB_CUDA(cudaSetDevice(ctx->device));
  const uint8_t* s0 = static_cast<const uint8_t*>(src_dev) +
                      (static_cast<size_t>(y0) * W + x0) * px_bytes;
  CERB_CUDA(cudaMemcpy2DAsync(dst, static_cast<size_t>(w) * px_bytes, s0,
                              static_cast<size_t>(W) * px_bytes, static_cast<size_t>(w) * px_bytes, h,
                              (flags & 2) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                              ctx->stream));
  return (flags & 2) ? CERB_OK : cerb_ctx_sync(ctx);
}

extern "C" int cerb_region_half(cerb_ctx* ctx, const float* canvas_dev, int H, int W, int C, int y0,
                                int x0, int h, int w, const uint8_t* mask_host, const int32_t* chans,
                                int k, float* out_dev, int oh, int ow) {
  if (!ctx || !canvas_dev || !chans || !out_dev || k <= 0 || k > 16 || y0 < 0 || x0 < 0 || h <= 0 ||
      w <= 0 || y0 + h > H || x0 + w > W || oh <= 0 || ow <= 0 || 2 * (oh - 1) > h - 1 + 1 ||
      2 * (ow - 1) > w - 1 + 1)
    return fail(CERB_ERR_ARG, "cerb_region_half: bad arguments");
  for (int i = 0; i < k; ++i)
    if (chans[i] < 0 || chans[i] >= C) return fail(CERB_ERR_ARG, "cerb_region_half: bad channel");
  CERB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevBuf dm(ctx), dc(ctx);
  if (mask_host != nullptr) {
    CERB_CUDA(cudaMalloc(&dm.p, static_cast<size_t>(h) * w));
    CERB_CUDA(cudaMemcpyAsync(dm.p, mask_host, static_cast<size_t>(h) * w, cudaMemcpyHostToDevice, s));
  }
  CERB_CUDA(cudaMalloc(&dc.p, sizeof(int) * k));
  CERB_CUDA(cudaMemcpyAsync(dc.p, chans, sizeof(int) * k, cudaMemcpyHostToDevice, s));
  k_region_half<<<148 * 8, 256, 0, s>>>(canvas_dev, W, C, y0, x0, h, w,
                                        static_cast<const uint8_t*>(dm.p),
                                        static_cast<const int*>(dc.p), k, out_dev, oh, ow);
  CERB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return cerb_ctx_sync(ctx);
}

extern "C" int cerb_nearest_channel(cerb_ctx* ctx, const float* canvas_dev, int H, int W, int C,
                                    int ch, double scale, float* out_host, int oh, int ow) {
  if (!ctx || !canvas_dev || !out_host || ch < 0 || ch >= C || !(scale > 0.0) || oh <= 0 || ow <= 0)
    return fail(CERB_ERR_ARG, "cerb_nearest_channel: bad arguments");
  CERB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevBuf d(ctx);
  const size_t bytes = static_cast<size_t>(oh) * ow * sizeof(float);
  CERB_CUDA(cudaMalloc(&d.p, bytes));
  k_nearest_channel<<<148 * 8, 256, 0, s>>>(canvas_dev, H, W, C, ch, static_cast<float>(1.0 / scale),
                                            static_cast<float*>(d.p), oh, ow);
  CERB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  CERB_CUDA(cudaMemcpyAsync(out_host, d.p, bytes, cudaMemcpyDeviceToHost, s));
  return cerb_ctx_sync(ctx);
}

// ------------------------------------------------------------------ channel plane of an HWC canvas
namespace {
__global__ void k_channel_plane(const float* __restrict__ canvas, size_t hw, int C, int ch,
                                float* __restrict__ out) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < hw;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = canvas[i * C + ch];
}
}  // namespace

extern "C" int cerb_channel_plane(cerb_ctx* ctx, const float* canvas_dev, int H, int W, int C, int ch,
                                  float* out, int flags) {
  if (!ctx || !canvas_dev || !out || H <= 0 || W <= 0 || ch < 0 || ch >= C)
    return fail(CERB_ERR_ARG, "cerb_channel_plane: bad arguments");
  CERB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t hw = static_cast<size_t>(H) * W;
  DevBuf d(ctx);
  float* dst = out;
  if (!(flags & 2)) {
    CERB_CUDA(cudaMalloc(&d.p, hw * sizeof(float)));
    dst = static_cast<float*>(d.p);
  }
  k_channel_plane<<<148 * 8, 256, 0, s>>>(canvas_dev, hw, C, ch, dst);
  CERB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  if (flags & 2) return CERB_OK;  // queued on the ctx stream
  CERB_CUDA(cudaMemcpyAsync(out, dst, hw * sizeof(float), cudaMemcpyDeviceToHost, s));
  return cerb_ctx_sync(ctx);
}
