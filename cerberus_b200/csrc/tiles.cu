// Tile plumbing on the device: reflect-padded patch extraction (infer/tile.py:64-69 +
// loader/infer_loader.py:57-69) and canvas stitching (infer/tile.py:136-163). Byte / fp32
// copy kernels, HBM-bound; vectorised where the layout allows.
#include <cuda_runtime.h>

#include <cstdint>
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
                          const int* __restrict__ tl, int ph, int pw, uint8_t* __restrict__ out) {
  const int patch = blockIdx.y;
  const int ty = tl[2 * patch], tx = tl[2 * patch + 1];
  const int total = ph * pw;
  uint8_t* o = out + static_cast<size_t>(patch) * total * 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / pw, x = i % pw;
    const int sy = reflect_index(ty + y - pad_t, H);
    const int sx = reflect_index(tx + x - pad_l, W);
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
  CERB_CUDA(cudaMalloc(&dtl.p, sizeof(int) * 2 * n));
  CERB_CUDA(cudaMemcpyAsync(dtl.p, tl_yx, sizeof(int) * 2 * n, cudaMemcpyHostToDevice, s));
  uint8_t* dst = out;
  if (!(flags & 2)) {
    CERB_CUDA(cudaMalloc(&dout.p, out_bytes));
    dst = static_cast<uint8_t*>(dout.p);
  }
  int gx = (ph * pw + 255) / 256;
  if (gx > 64) gx = 64;
  k_extract<<<dim3(gx, n), 256, 0, s>>>(src, H, W, pad_t, pad_l, static_cast<const int*>(dtl.p), ph,
                                        pw, dst);
  CERB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  if (!(flags & 2)) {
    CERB_CUDA(cudaMemcpyAsync(out, dst, out_bytes, cudaMemcpyDeviceToHost, s));
  }
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
