// 3x3 stride-1 64->64 convolution in the split-precision (parity) mode: activations and weights are
// hi+lo fp16 pairs, every product is three MMAs (hi*hi + lo*hi + hi*lo), fp32 accumulation in three
// TMEM accumulators per tile that the epilogue adds in fp32 (conv_tc.cu explains why three).
//
// The generic kernel runs these layers - 53 % of the parity-mode forward (models/backbone/
// resnet.py:202 layer1, the decoder stages u2 / u1 of models/utils/net_layers.py:23-28) - at the
// L2 -> SM delivery limit: per filter tap it fetches a 128-pixel activation slab (hi and lo) and
// the tap's weights (hi and lo), 48 KB for 576 cycles of MMA. Here
//   * the 18x10 halo of a 16x8-pixel tile is loaded ONCE (hi and lo, 46 KB) and the nine taps are
//     nine shared-memory descriptors into it (as in conv64.cu / conv3x3.cu);
//   * the hi weights of all nine taps stay resident in shared memory for the life of the
//     persistent CTA (72 KB); only the lo weights (needed by one MMA in three) stream, 8 KB per tap;
//   * two accumulator stages (2 x 3 x 64 TMEM columns): the epilogue of a tile (one group of four
//     warps, ~2000 cycles) overlaps the MMAs of the next (5184 cycles); the generic split kernel has
//     a single stage.
// 118 KB from L2 per 5184 cycles of MMA instead of 432 KB: the kernel is MMA-bound.
// Epilogue: a thread owns a pixel; (main0 + main1 + cross) * 2^-shift + bias (+ residual hi + lo) ->
// ReLU -> hi = fp16(v), lo = fp16(v - hi) -> two swizzled staging tiles -> two TMA stores.
#include "conv64s.cuh"
#include "ptx.cuh"

namespace cerb {

namespace {

#define CERB_PROF_T0(var) const long long var = p.prof != nullptr ? clock64() : 0
#define CERB_PROF_ADD(acc, var) \
  do { if (p.prof != nullptr) acc += clock64() - var; } while (0)

constexpr int kTH = 16, kTW = 8;                    // tile: 16 rows x 8 columns = 128 pixels
constexpr int kHaloH = 18, kHaloW = 10;
constexpr int kHaloTx = kHaloH * kHaloW * 128;      // 23040 bytes per plane
constexpr int kHaloBytes = 23 * 1024;               // padded to the swizzle period
constexpr int kStageBytes = 2 * kHaloBytes;         // hi | lo
constexpr int kTapBytes = 64 * 128;                 // one tap: 64 output channels x 64 fp16
constexpr int kWBytes = 9 * kTapBytes;              // resident hi weights
constexpr int kLoStages = 3;                        // streamed lo-weight taps
constexpr int kOutBytes = 128 * 128;                // staging: 128 pixels x 64 fp16 (one plane)
constexpr int kAStages = 2;
constexpr int kAccStages = 2;
constexpr int kAccStride = 256;                     // TMEM columns per stage: main0 | main1 | cross | -
constexpr int kTmemCols = 512;

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kConv64sThreads, 1)
conv64s_kernel(const __grid_constant__ Conv64sParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint8_t* sW = smem;                               // 9 resident hi taps
  uint8_t* sWlo = sW + kWBytes;                     // kLoStages streamed lo taps
  uint8_t* sA = sWlo + kLoStages * kTapBytes;       // kAStages x (hi halo | lo halo)
  uint8_t* sOut = sA + kAStages * kStageBytes;      // hi | lo staging tiles
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sOut + 2 * kOutBytes);
  uint64_t* a_empty = a_full + kAStages;
  uint64_t* l_full = a_empty + kAStages;
  uint64_t* l_empty = l_full + kLoStages;
  uint64_t* tfull_bar = l_empty + kLoStages;
  uint64_t* tempty_bar = tfull_bar + kAccStages;
  uint64_t* w_bar = tempty_bar + kAccStages;
  uint64_t* res_bar = w_bar + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(res_bar + 1);
  volatile int* s_ring = reinterpret_cast<volatile int*>(tmem_holder + 2);
  volatile int* s_nt = s_ring + 8;   // tiles fetched so far by the halo producer
  volatile int* s_done = s_ring + 9; // set after the last tile was fetched

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.in_hi);
    ptx::prefetch_tmap(&p.in_lo);
    ptx::prefetch_tmap(&p.w_hi);
    ptx::prefetch_tmap(&p.w_lo);
    ptx::prefetch_tmap(&p.out_hi);
    ptx::prefetch_tmap(&p.out_lo);
    if (p.has_res) {
      ptx::prefetch_tmap(&p.res_hi);
      ptx::prefetch_tmap(&p.res_lo);
    }
    for (int s = 0; s < kAStages; ++s) {
      ptx::mbar_init(&a_full[s], 1);
      ptx::mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kLoStages; ++s) {
      ptx::mbar_init(&l_full[s], 1);
      ptx::mbar_init(&l_empty[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      ptx::mbar_init(&tfull_bar[s], 1);
      ptx::mbar_init(&tempty_bar[s], 4);  // the four epilogue warps
    }
    ptx::mbar_init(w_bar, 1);
    ptx::mbar_init(res_bar, 1);
    *s_nt = 0;
    *s_done = 0;
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_holder, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  ptx::grid_dep_launch();

  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    const bool leader = ptx::elect_one() != 0;
    const int leader_lane = __ffs(__ballot_sync(0xffffffffu, leader)) - 1;
    if (leader) {  // resident hi weights: do not depend on the previous kernel
      ptx::mbar_arrive_expect_tx(w_bar, kWBytes);
#pragma unroll
      for (int t = 0; t < 9; ++t) ptx::tma_load_2d(sW + t * kTapBytes, &p.w_hi, w_bar, t * 64, 0);
    }
    ptx::grid_dep_wait();
    int a_issued = 0;
    for (int i = 0;; ++i) {
      int tile = 0;
      if (p.tile_counter != nullptr) {
        if (leader) tile = atomicAdd(p.tile_counter, 1);
        tile = __shfl_sync(0xffffffffu, tile, leader_lane);
      } else {
        tile = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      }
      const bool done = tile >= p.n_tiles;
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int st = a_issued % kAStages;
      ptx::mbar_wait(&a_empty[st], ((a_issued / kAStages) & 1) ^ 1, p.err_flag, 71);
      if (leader) {
        s_ring[i & 7] = done ? -1 : tile;
        if (done) {
          ptx::mbar_arrive(&a_full[st]);
        } else {
          uint8_t* dst = sA + st * kStageBytes;
          ptx::mbar_arrive_expect_tx(&a_full[st], 2 * kHaloTx);
          ptx::tma_load_4d(dst, &p.in_hi, &a_full[st], 0, tx * kTW - 1, ty * kTH - 1, img);
          ptx::tma_load_4d(dst + kHaloBytes, &p.in_lo, &a_full[st], 0, tx * kTW - 1, ty * kTH - 1, img);
        }
      }
      __syncwarp();
      ++a_issued;
      if (leader) {  // tell the lo-weight producer (warp 6) how many taps it may stream
        if (done) {
          __threadfence_block();
          *s_done = 1;
        } else {
          *s_nt = i + 1;
        }
      }
      __syncwarp();
      if (done) break;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = ptx::elect_one() != 0;
    const uint32_t idesc = ptx::umma_idesc_f16(128, 64);
    ptx::mbar_wait(w_bar, 0, p.err_flag, 73);
    ptx::tc_fence_after();
    const uint64_t a_d0 = ptx::umma_desc_sw128(ptx::smem_u32(sA), kHaloW * 128);
    const uint64_t w_d0 = ptx::umma_desc_sw128(ptx::smem_u32(sW), 1024);
    const uint64_t l_d0 = ptx::umma_desc_sw128(ptx::smem_u32(sWlo), 1024);
    constexpr uint32_t kLoA = static_cast<uint32_t>(kHaloBytes >> 4);
    long long prof_a = 0, prof_b = 0, prof_c = 0, prof_d = 0;
    CERB_PROF_T0(t_all);
    int l_cnt = 0;
    for (int i = 0;; ++i) {
      const int acc = i & 1;
      const int ast = i % kAStages;
      CERB_PROF_T0(t_m0);
      ptx::mbar_wait(&tempty_bar[acc], ((i >> 1) & 1) ^ 1, p.err_flag, 74);
      CERB_PROF_ADD(prof_a, t_m0);
      CERB_PROF_T0(t_m1);
      ptx::mbar_wait(&a_full[ast], (i / kAStages) & 1, p.err_flag, 75);
      CERB_PROF_ADD(prof_b, t_m1);
      if (s_ring[i & 7] < 0) {  // end marker: pass it on to the epilogue
        if (leader) ptx::mbar_arrive(&tfull_bar[acc]);
        __syncwarp();
        break;
      }
      ptx::tc_fence_after();
      const uint32_t d_main0 = tmem_base + acc * kAccStride;
      const uint32_t d_main1 = d_main0 + 64;
      const uint32_t d_cross = d_main0 + 128;
      const uint64_t a_hi = a_d0 + static_cast<uint32_t>((ast * kStageBytes) >> 4);
      // The hi*lo MMAs of tap t are issued one tap LATE (after the eight hi-weight MMAs of tap
      // t + 1): the streamed lo weights of a tap then have two more MMA groups' time to arrive.
      auto issue_lo = [&](int t) {
        const uint32_t tap_u = static_cast<uint32_t>((((t / 3) * kHaloW + (t % 3)) * 128) >> 4);
        const int ls = l_cnt % kLoStages;
        CERB_PROF_T0(t_m2);
        ptx::mbar_wait(&l_full[ls], (l_cnt / kLoStages) & 1, p.err_flag, 76);
        CERB_PROF_ADD(prof_c, t_m2);
        ptx::tc_fence_after();
        CERB_PROF_T0(t_m4);
        if (leader) {
          const uint64_t ld = l_d0 + static_cast<uint32_t>((ls * kTapBytes) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // hi(activation) * lo(weight)
            ptx::umma_f16(d_cross, a_hi + tap_u + 2 * k, ld + 2 * k, idesc, 1);
          ptx::umma_commit(&l_empty[ls]);
          if (t == 8) {
            ptx::umma_commit(&a_empty[ast]);
            ptx::umma_commit(&tfull_bar[acc]);
          }
        }
        __syncwarp();
        CERB_PROF_ADD(prof_d, t_m4);
        ++l_cnt;
      };
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const uint32_t tap_u = static_cast<uint32_t>((((t / 3) * kHaloW + (t % 3)) * 128) >> 4);
        const uint32_t w_u = static_cast<uint32_t>((t * kTapBytes) >> 4);
        CERB_PROF_T0(t_m3);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // hi * hi alternates between two accumulators (halves the per-MMA truncation error)
            ptx::umma_f16((k & 1) ? d_main1 : d_main0, a_hi + tap_u + 2 * k, w_d0 + w_u + 2 * k, idesc,
                          (t != 0) || (k >= 2));
            // lo(activation) * hi(weight)
            ptx::umma_f16(d_cross, a_hi + kLoA + tap_u + 2 * k, w_d0 + w_u + 2 * k, idesc, (t | k) != 0);
          }
        }
        __syncwarp();
        CERB_PROF_ADD(prof_d, t_m3);
        if (t > 0) issue_lo(t - 1);
      }
      issue_lo(8);
    }
    if (p.prof != nullptr && lane == 0) {
      long long* o = p.prof + blockIdx.x * 16;
      o[1] = prof_a; o[2] = prof_b + prof_c; o[3] = prof_d; o[8] = clock64() - t_all; o[9] = prof_b;
    }
  } else if (warp == 6) {
    // ------------------------------------------------------------------ lo-weight producer
    // The same nine 8 KB taps for every tile, streamed through kLoStages buffers. A warp of its
    // own: interleaved with the halo loads in warp 0 (both block on their "stage free" barriers)
    // either the next halo or the next taps were requested late (measured: 10 % + 19 % of the
    // kernel waiting for operands).
    const bool leader = ptx::elect_one() != 0;
    int issued = 0;
    for (;;) {
      const int d = *s_done;
      const int n = *s_nt;
      if (issued >= 9 * n) {
        if (d) break;
        __nanosleep(64);
        continue;
      }
      const int ls = issued % kLoStages;
      ptx::mbar_wait(&l_empty[ls], ((issued / kLoStages) & 1) ^ 1, p.err_flag, 72);
      if (leader) {
        ptx::mbar_arrive_expect_tx(&l_full[ls], kTapBytes);
        ptx::tma_load_2d(sWlo + ls * kTapBytes, &p.w_lo, &l_full[ls], (issued % 9) * 64, 0);
      }
      __syncwarp();
      ++issued;
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    ptx::grid_dep_wait();
    const int q = warp & 3;
    const int m = q * 32 + lane;  // TMEM lane = pixel (y, x): y = m >> 3, x = m & 7
    const int sw = m & 7;
    const bool store_warp = q == 2;
    uint8_t* sHi = sOut;
    uint8_t* sLo = sHi + kOutBytes;
    uint8_t* row_hi = sHi + m * 128;
    uint8_t* row_lo = sLo + m * 128;
    for (int i = 0;; ++i) {
      const int acc = i & 1;
      ptx::mbar_wait(&tfull_bar[acc], (i >> 1) & 1, p.err_flag, 77);
      const int tile = s_ring[i & 7];
      if (tile < 0) break;
      ptx::tc_fence_after();
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int x0 = tx * kTW, y0 = ty * kTH;
      if (store_warp && ptx::elect_one()) {
        ptx::bulk_wait_read<0>();  // the previous stores have drained the staging tiles
        if (p.has_res) {
          ptx::mbar_arrive_expect_tx(res_bar, 2 * kOutBytes);
          ptx::tma_load_4d(sHi, &p.res_hi, res_bar, 0, x0, y0, img);
          ptx::tma_load_4d(sLo, &p.res_lo, res_bar, 0, x0, y0, img);
        }
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccStride;
      bool waited = false;
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t r0[32], r1[32], r2[32];
        ptx::tmem_ld32(taddr + c2 * 32, r0);
        ptx::tmem_ld32(taddr + 64 + c2 * 32, r1);
        ptx::tmem_ld32(taddr + 128 + c2 * 32, r2);
        ptx::tmem_ld_wait();
        if (c2 == 1) {  // the three accumulators are in registers: hand the stage back
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
        }
        if (!waited) {
          if (p.has_res) {
            ptx::mbar_wait(res_bar, i & 1, p.err_flag, 78);
          } else {
            ptx::named_bar_sync(1, 128);  // the elected lane has seen the staging tiles drained
          }
          waited = true;
        }
        float v[32];
        const float sc = p.acc_scale;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          v[j] = ((__uint_as_float(r0[j]) + __uint_as_float(r1[j])) + __uint_as_float(r2[j])) * sc;
        if (p.bias != nullptr) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + c2 * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
          }
        }
        if (p.has_res) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int off = ((c2 * 4 + j) ^ sw) << 4;
            const uint4 uh = *reinterpret_cast<const uint4*>(row_hi + off);
            const uint4 ul = *reinterpret_cast<const uint4*>(row_lo + off);
            const __half2* hh = reinterpret_cast<const __half2*>(&uh);
            const __half2* hl = reinterpret_cast<const __half2*>(&ul);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 fh = __half22float2(hh[e]), fl = __half22float2(hl[e]);
              v[8 * j + 2 * e] += fh.x + fl.x;
              v[8 * j + 2 * e + 1] += fh.y + fl.y;
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = v[8 * j + 2 * e], b = v[8 * j + 2 * e + 1];
            const __half2 h = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(h);
            hi[e] = *reinterpret_cast<const uint32_t*>(&h);
            lo[e] = pack_half2(a - hf.x, b - hf.y);
          }
          const int off = ((c2 * 4 + j) ^ sw) << 4;
          *reinterpret_cast<uint4*>(row_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(row_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(1, 128);
      if (store_warp && ptx::elect_one()) {
        ptx::tma_store_4d(&p.out_hi, sHi, 0, x0, y0, img);
        ptx::tma_store_4d(&p.out_lo, sLo, 0, x0, y0, img);
        ptx::bulk_commit_group();
      }
    }
    if (store_warp) ptx::bulk_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

void conv64s_plan(Conv64sParams& p) {
  p.tiles_x = (p.W + kTW - 1) / kTW;
  p.tiles_y = (p.H + kTH - 1) / kTH;
  p.n_tiles = p.n_img * p.tiles_x * p.tiles_y;
}

size_t conv64s_smem_bytes(const Conv64sParams&) {
  return static_cast<size_t>(kWBytes) + kLoStages * kTapBytes + kAStages * kStageBytes + 2 * kOutBytes +
         512 + 1024;
}

cudaError_t conv64s_launch(const Conv64sParams& p, int num_sms, cudaStream_t stream, bool pdl) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv64s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kConv64sThreads);
  cfg.dynamicSmemBytes = conv64s_smem_bytes(p);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, conv64s_kernel, p);
}

}  // namespace cerb
