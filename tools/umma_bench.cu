// Microbenchmark: issue rate of tcgen05.mma (kind::f16, cta_group::1, M = 128, SS operands) as a
// function of N and of the A-descriptor alignment. One CTA per SM, one issuing thread, operands
// are whatever is in shared memory (values do not matter). Evidence for DESIGN.md section 3.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_bench tools/umma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../cerberus_b200/csrc/ptx.cuh"

using namespace cerb;

struct Args {
  int n;          // MMA N
  int iters;      // MMAs per measurement
  int a_step;     // bytes added to the A start address between consecutive MMAs (mod window)
  int a_window;   // A addresses cycle inside this many bytes
  int b_step;
  int b_window;
  int a_off;      // constant byte offset of A (alignment experiment)
  int sbo;        // SBO bytes
  int n_acc;      // round-robin over this many accumulators
  long long* out; // [grid] cycles
};

__global__ void __launch_bounds__(128, 1) umma_rate(Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // deterministic small values so nothing overflows to inf/nan (does not matter for timing)
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x2C002C00u;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&holder, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = holder;
  if (warp == 1) {
    const bool leader = ptx::elect_one();
    const uint32_t idesc = ptx::umma_idesc_f16(128, a.n);
    const uint32_t a_base = ptx::smem_u32(smem) + a.a_off;
    const uint32_t b_base = ptx::smem_u32(smem) + 100 * 1024;
    const uint64_t ad0 = ptx::umma_desc_sw128(a_base, a.sbo);
    const uint64_t bd0 = ptx::umma_desc_sw128(b_base, 1024);
    // warm-up
    for (int i = 0; i < 64; ++i) if (leader) ptx::umma_f16(tmem, ad0, bd0, idesc, 1);
    if (leader) ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0, nullptr, 0);
    const long long t0 = clock64();
    if (a.a_step == 0) {
      // unrolled: 16 MMAs per trip, descriptors = base + compile-time constants (k-walk of 4 x 32 B)
      for (int i = 0; i < a.iters; i += 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (leader) ptx::umma_f16(tmem + (j & 1) * 256, ad0 + 2 * (j & 3), bd0 + 2 * (j & 3), idesc, 1);
      }
    } else if (a.a_step == 1) {
      // conv64 style: 9 taps x 4 k-steps, A = halo view (r*10+s)*128 B, B = tap*8192 B
      for (int i = 0; i < a.iters; i += 36) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (leader)
              ptx::umma_f16(tmem, ad0 + (((t / 3) * 10 + (t % 3)) * 128 >> 4) + 2 * k, bd0 + t * 512 + 2 * k, idesc, 1);
      }
    } else {
      int ao = 0, bo = 0, acc = 0;
      for (int i = 0; i < a.iters; ++i) {
        if (leader) ptx::umma_f16(tmem + acc * a.n, ad0 + (ao >> 4), bd0 + (bo >> 4), idesc, 1);
        ao += a.a_step; if (ao >= a.a_window) ao = 0;
        bo += a.b_step; if (bo >= a.b_window) bo = 0;
        if (++acc == a.n_acc) acc = 0;
      }
    }
    if (leader) ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 1, nullptr, 0);
    const long long t1 = clock64();
    if (leader) a.out[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

static double run(Args a, int grid, const char* label) {
  long long* d;
  cudaMalloc(&d, sizeof(long long) * grid);
  a.out = d;
  cudaFuncSetAttribute(umma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  umma_rate<<<grid, 128, 210 * 1024>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: CUDA error %s\n", label, cudaGetErrorString(e));
    exit(1);
  }
  long long* h = (long long*)malloc(sizeof(long long) * grid);
  cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double mx = 0, sum = 0;
  for (int i = 0; i < grid; ++i) { sum += h[i]; if (h[i] > mx) mx = (double)h[i]; }
  const double cyc = sum / grid / a.iters;
  printf("%-58s N=%3d  %7.1f cyc/MMA (max-CTA %7.1f)  -> %6.0f MAC/cyc/SM  smem %5.1f B/cyc\n", label, a.n,
         cyc, mx / a.iters, 128.0 * a.n * 16 / cyc, (4096.0 + a.n * 32.0) / cyc);
  cudaFree(d);
  free(h);
  return cyc;
}

int main() {
  const int grid = 148;
  const int ns[] = {16, 32, 64, 96, 128, 192, 256};
  for (int n : ns) {
    Args a{n, 4096, 0, 1, 0, 1, 0, 1024, 1, nullptr};
    run(a, grid, "unrolled x16, constant descriptors, 2 accumulators");
  }
  for (int n : {64}) {
    Args a{n, 4608, 1, 1, 0, 1, 0, 1280, 1, nullptr};
    run(a, grid, "conv64 pattern: 9 taps x 4 k, halo views, resident W");
  }
  for (int n : {64, 128}) {
    Args a{n, 4096, 32, 128, 32, 128, 0, 1024, 1, nullptr};
    run(a, grid, "rolled loop with loop-carried uniform address chain");
  }
  for (int n : {64, 128, 256}) {
    Args a{n, 4096, 0, 1, 0, 1, 0, 1024, 1, nullptr};
    run(a, 1, "ONE CTA only, unrolled x16");
  }
  return 0;
}
