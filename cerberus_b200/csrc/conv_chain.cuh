// Layer chains of the wide 3x3 kernels (conv3x3.cu, conv3x3c2.cu): a run of consecutive layers of
// identical geometry, each reading the output of the one before (the bodies of encoder layer2-4,
// models/backbone/resnet.py:203-211), executed by ONE persistent launch. Work items are numbered
// layer-major and drawn in that order from the global counter; an item of layer l on image i may
// request its first halo once `done[(l - 1) * n_img + i]` shows that every item of layer l - 1 on
// that image has been stored. See DESIGN.md 3.2b.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "ptx.cuh"

namespace cerb {

// What differs between the layers of a chain. Tables live in global memory (tensor maps are
// passed to TMA by generic address); a single-layer launch carries one in its kernel parameters.
struct ConvChainLayer {
  CUtensorMap in_map;   // input halo boxes
  CUtensorMap w_map;    // weight slabs
  CUtensorMap out_map;  // output slabs (TMA stores)
  CUtensorMap res_map;  // residual, same geometry as out_map
  const float* bias;    // [Cout] fp32 (BN folded), may be null
  float acc_scale;      // 2^-w_shift
  int has_res;
  int relu;
  int pad_[11];
};
static_assert(sizeof(ConvChainLayer) % 64 == 0, "tensor maps of a layer table must stay 64-byte aligned");

#ifdef __CUDACC__
namespace chain {

__device__ __forceinline__ int ld_acquire_gpu(const int* ptr) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;" ::: "memory");
}

// Has `*cnt` reached `need`? With block = false the answer may be no; otherwise spins (2 s
// watchdog). One lane calls it - the lane that issues the TMA loads that depend on the answer.
__device__ __forceinline__ bool wait_count(const int* cnt, int need, bool block, int* err_flag, int code) {
  if (ld_acquire_gpu(cnt) < need) {
    if (!block) return false;
    const uint64_t t0 = ptx::global_timer_ns();
    uint32_t spins = 0;
    while (ld_acquire_gpu(cnt) < need) {
      if ((++spins & 0x3FF) == 0 && ptx::global_timer_ns() - t0 > 2000000000ull) {
        if (err_flag) atomicExch(err_flag, code);
        __threadfence_system();
        asm volatile("trap;");
      }
    }
  }
  fence_proxy_async_all();  // the TMA loads that follow read what other CTAs' TMA stores wrote
  return true;
}

// Called by the thread whose bulk-async groups hold the TMA stores of the finished item.
__device__ __forceinline__ void signal_stored(int* cnt) {
  ptx::bulk_wait_all<0>();  // stores complete (not only read from shared memory)
  fence_proxy_async_all();
  __threadfence();
  atomicAdd(cnt, 1);
}

}  // namespace chain
#endif

}  // namespace cerb
