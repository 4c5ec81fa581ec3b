// tcgen05 / TMA implicit-GEMM convolution kernel. See conv_tc.cuh for the scheme.
//
// Replaces, for the hot path, every nn.Conv2d + BatchNorm2d(+ReLU)(+residual add) the
// reference issues through cuDNN/ATen as separate kernels:
//   models/backbone/resnet.py:32-48,81-97,195-200   (encoder convs, BN, ReLU, residual)
//   models/utils/conv_layers.py:24-60               (decoder conv+bias -> BN -> ReLU)
//   models/net_desc.py:52                           (conv_map 1x1)
// BatchNorm is folded into weights/bias on load; bias, residual
// add, ReLU and the fp16 (or hi/lo split) store are fused in the TMEM epilogue.
#include "conv_tc.cuh"
#include "head_tail.cuh"
#include "ptx.cuh"

namespace cerb {

namespace {

// In-kernel attribution (ConvKParams::prof != nullptr): cycles a role spends in each phase.
#define CERB_PROF_T0(var) const long long var = p.prof != nullptr ? clock64() : 0
#define CERB_PROF_ADD(acc, var) \
  do { if (p.prof != nullptr) acc += clock64() - var; } while (0)

constexpr int kATileBytes = 128 * 128;  // 128 pixels x 64 fp16 channels
constexpr int kTmemCols = 512;
constexpr int kMaxAcc = 4;  // accumulator stages (= epilogue warp groups): 512 TMEM columns / n_acc each
// Split-precision mode (BN <= 128): ONE accumulator stage made of three TMEM accumulators.
// The tensor core truncates the fp32 accumulator on every MMA, an error proportional to
// |acc| per instruction; keeping the small cross terms (lo*hi + hi*lo) out of the big hi*hi
// accumulator and alternating hi*hi between two accumulators cuts that error ~6x. The
// epilogue adds the three in fp32 (round-to-nearest).
// split mode: the hi*hi products and the cross terms (hi*lo + lo*hi) accumulate separately (the lo
// weights carry their own power-of-two scale), and the hi*hi products of even / odd k16 steps go to
// two accumulators (halves the accumulation error, see conv_tc_plan_pipeline). Layouts: one stage
// of 3 x 128 columns (wide tiles, long K); or two stages of 256 columns so that the epilogue of a
// tile overlaps the MMAs of the next - 3 x 64 columns for tiles of at most 64 output channels
// (the 7-k-step stem, 128 -> 64), or ONE hi*hi accumulator + cross terms for wide tiles with at
// most two k-steps (the heads' 64 -> 96: four k16 products per accumulator either way).
__device__ __forceinline__ int split_main1(int n_acc, int bn) { return n_acc == 2 ? (bn > 64 ? 0 : 64) : 128; }
__device__ __forceinline__ int split_cross(int n_acc) { return n_acc == 2 ? 128 : 256; }
constexpr int kHeadA2Bytes = 2 * kATileBytes;  // hidden tile of one epilogue group: 2 slabs x 16 KB

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// max(x, 0) fused into the fp32 -> fp16x2 conversion (one instruction instead of two FMNMX + F2FP)
__device__ __forceinline__ uint32_t pack_half2_relu(float a, float b) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}

template <bool SPLIT, int MAX_THREADS>
__global__ void __launch_bounds__(MAX_THREADS, 1)
conv_tc_kernel(const __grid_constant__ ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment; the dynamic smem base only promises 16.
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_stages = p.n_stages;
  const int stage_bytes = p.stage_bytes;
  const int b_tile_bytes = p.BN * 128;

  // tile ids handed from the producer to the MMA / epilogue warps (dynamic scheduling, see
  // conv64x.cu); -1 ends the kernel. Lives behind the head-weight area.
  volatile int* s_ring = nullptr;
  // MMA tail (fp16 mode, fused classification head): per epilogue group a [128 x 96] fp16 hidden
  // tile (two 128-byte-swizzled 64-channel slabs) and the [16 x 96] fp16 head weights, both
  // K-major UMMA operands, sit between the pipeline stages and the barriers.
  const bool mma_tail = !SPLIT && p.mma_tail != 0;
  uint8_t* sA2 = smem + n_stages * stage_bytes;
  uint8_t* sB2 = sA2 + kMaxAcc * kHeadA2Bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(mma_tail ? sB2 + 4096 : sA2);
  uint64_t* empty_bar = full_bar + n_stages;
  uint64_t* tfull_bar = empty_bar + n_stages;
  uint64_t* tempty_bar = tfull_bar + kMaxAcc;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + kMaxAcc);
  uint64_t* head_bar = reinterpret_cast<uint64_t*>(tmem_holder + 2);  // [kMaxAcc], MMA-tail mode
  // [C][96] fp32 when the head tail is fused (16-byte aligned for float4 reads)
  float* s_head_w = reinterpret_cast<float*>(
      (reinterpret_cast<uintptr_t>(head_bar + kMaxAcc) + 15) & ~static_cast<uintptr_t>(15));
  s_ring = reinterpret_cast<volatile int*>(s_head_w + kHeadMaxC * 96);
  if (p.head_classes > 0) {
    if (!mma_tail) {
      for (int i = threadIdx.x; i < p.head_classes * 96; i += blockDim.x) s_head_w[i] = p.head_w[i];
    } else {
      // hidden-layer bias and head bias in shared memory (broadcast reads in the epilogue)
      for (int i = threadIdx.x; i < 96; i += blockDim.x) s_head_w[i] = p.bias != nullptr ? p.bias[i] : 0.0f;
      for (int i = threadIdx.x; i < kHeadMaxC; i += blockDim.x)
        s_head_w[96 + i] = i < p.head_classes ? p.head_b[i] : 0.0f;
      // B2[n][k] = fp16(head_w[n][k]) for n < C, zero rows up to N = 16
      for (int i = threadIdx.x; i < 16 * 96; i += blockDim.x) {
        const int n = i / 96, k = i - n * 96;
        const float w = n < p.head_classes ? p.head_w[n * 96 + k] : 0.0f;
        const int slab = k >> 6, kk = k & 63;
        *reinterpret_cast<__half*>(sB2 + slab * 2048 + n * 128 + ((((kk >> 3) ^ (n & 7))) << 4) +
                                   (kk & 7) * 2) = __float2half_rn(w);
      }
      ptx::fence_proxy_async_smem();
    }
  }

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) ptx::prefetch_tmap(&p.in_hi[i]);
    ptx::prefetch_tmap(&p.w_hi);
    if (SPLIT) {
      for (int i = 0; i < 4; ++i) ptx::prefetch_tmap(&p.in_lo[i]);
      ptx::prefetch_tmap(&p.w_lo);
    }
    for (int s = 0; s < n_stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kMaxAcc; ++s) {
      ptx::mbar_init(&tfull_bar[s], 1);
      ptx::mbar_init(&tempty_bar[s], 4);  // one arrival per epilogue warp
      ptx::mbar_init(&head_bar[s], 1);
    }
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
  ptx::grid_dep_wait();  // prologue above overlaps the previous kernel's tail (PDL)

  const int kAccStages = p.n_acc;  // split mode: 1, or 2 for BN <= 64 (conv_tc_plan_pipeline)
  const int kAccStride = 512 / p.n_acc;
  const int kSplitMain1 = split_main1(p.n_acc, p.BN), kSplitCross = split_cross(p.n_acc);
  const bool split_merged = SPLIT && kSplitMain1 == 0;  // one hi*hi accumulator
  const int n_ksteps = p.n_taps * p.n_chunks;
  const int bw_mask = (1 << p.bw_log2) - 1;
  const int BW = 1 << p.bw_log2;
  const int BH = 128 >> p.bw_log2;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp walks the loops and ONE elected lane issues: a branch on `lane == 0` makes
    // ptxas wrap every uniform-datapath instruction (UTMALDG, UTCHMMA) in an ELECT/BRA.U.ANY
    // serialisation loop (~78 cycles per instruction; tools/umma_bench.cu, profiles/).
    const bool leader = ptx::elect_one() != 0;
    int stage = 0;
    uint32_t phase = 0;
    const bool a_lo_zero = SPLIT && p.a_lo_zero != 0;
    const uint32_t tx_bytes =
        (SPLIT ? 2u : 1u) * (kATileBytes + b_tile_bytes) - (a_lo_zero ? static_cast<uint32_t>(kATileBytes) : 0u);
    long long prof_a = 0;
    const int leader_lane = __ffs(__ballot_sync(0xffffffffu, leader)) - 1;
    // Dynamic scheduling: the atomic that draws tile i + 1 is issued BEFORE the loads of tile i and
    // its result is first touched one iteration later, so its L2 round trip (the producer's whole
    // per-tile critical path when a tile is one k-step, as in the heads) overlaps useful work:
    // heads 0.630 -> 0.588 ms per step. Every CTA draws one index more than it uses. (Drawing two
    // tiles ahead and asking L2 for the second one's activation box, as conv64x.cu does for its
    // halos, did not help here: 0.619 ms.)
    int drawn = 0;
    if (p.tile_counter != nullptr && leader) drawn = atomicAdd(p.tile_counter, 1);
    for (int it = 0;; ++it) {
      int tile = 0;
      if (p.tile_counter != nullptr) {
        tile = __shfl_sync(0xffffffffu, drawn, leader_lane);
        if (leader && tile < p.n_tiles) drawn = atomicAdd(p.tile_counter, 1);
      } else {
        tile = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      }
      if (tile >= p.n_tiles) {
        // end marker for the MMA warp and for every epilogue group (each waits for its own next tile)
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 1);
        if (leader) {
          for (int g = 0; g < kMaxAcc; ++g) s_ring[(it + g) & 15] = -1;
          ptx::mbar_arrive(&full_bar[stage]);
        }
        __syncwarp();
        break;
      }
      if (leader) s_ring[it & 15] = tile;  // published by the arrive of the tile's first k-step
      const int nt = tile % p.n_ntiles;
      const int mt = tile / p.n_ntiles;
      const int tx = mt % p.tiles_x;
      const int ty = (mt / p.tiles_x) % p.tiles_y;
      const int img = mt / (p.tiles_x * p.tiles_y);
      const int x0 = tx * BW, y0 = ty * BH;
      for (int t = 0; t < p.n_taps; ++t) {
        const ConvTap tap = p.taps[t];
        for (int c = 0; c < p.n_chunks; ++c) {
          CERB_PROF_T0(t_p0);
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 1);
          CERB_PROF_ADD(prof_a, t_p0);
          if (leader) {
            uint8_t* sA = smem + stage * stage_bytes;
            uint8_t* sB = sA + (SPLIT ? 2 : 1) * kATileBytes;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            ptx::tma_load_4d(sA, &p.in_hi[tap.map], &full_bar[stage], c * 64, x0 + tap.dx,
                             y0 + tap.dy, img);
            ptx::tma_load_2d(sB, &p.w_hi, &full_bar[stage], (t * p.n_chunks + c) * 64,
                             nt * p.BN);
            if (SPLIT) {
              if (!a_lo_zero)
                ptx::tma_load_4d(sA + kATileBytes, &p.in_lo[tap.map], &full_bar[stage], c * 64,
                                 x0 + tap.dx, y0 + tap.dy, img);
              ptx::tma_load_2d(sB + b_tile_bytes, &p.w_lo, &full_bar[stage],
                               (t * p.n_chunks + c) * 64, nt * p.BN);
            }
          }
          __syncwarp();
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    if (p.prof != nullptr && lane == 0) p.prof[blockIdx.x * 16 + 0] = prof_a;
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const bool leader = ptx::elect_one() != 0;
    const uint32_t idesc = ptx::umma_idesc_f16(128, p.BN);
    // descriptor of stage 0, k-step 0; only the 14-bit start-address field (16-byte units) moves
    const uint64_t a_d0 = ptx::umma_desc_sw128(ptx::smem_u32(smem), 1024);
    const uint32_t stage_u = static_cast<uint32_t>(stage_bytes >> 4);
    const uint32_t b_off_u = static_cast<uint32_t>(((SPLIT ? 2 : 1) * kATileBytes) >> 4);
    const uint32_t lo_a_u = static_cast<uint32_t>(kATileBytes >> 4);
    const uint32_t lo_b_u = static_cast<uint32_t>(b_tile_bytes >> 4);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long prof_a = 0, prof_b = 0, prof_c = 0;
    CERB_PROF_T0(t_all);
    bool done = false;
    for (int it = 0; !done; ++it) {
      CERB_PROF_T0(t_m0);
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1, p.err_flag, 2);
      CERB_PROF_ADD(prof_a, t_m0);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * kAccStride;
      for (int ks = 0; ks < n_ksteps; ++ks) {
        CERB_PROF_T0(t_m1);
        ptx::mbar_wait(&full_bar[stage], phase, p.err_flag, 3);
        CERB_PROF_ADD(prof_b, t_m1);
        if (ks == 0 && s_ring[it & 15] < 0) {
          // end marker: wake every epilogue group on the accumulator stage it waits for next.
          // A stage is signalled only after its last real tile was drained (tempty), so the plain
          // arrive can never overtake a pending tcgen05.commit on the same barrier.
          for (int g = 0; g < kAccStages; ++g) {
            if (g > 0) ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1, p.err_flag, 2);
            if (leader) ptx::mbar_arrive(&tfull_bar[acc]);
            __syncwarp();
            acc = (acc + 1) % kAccStages;
            if (acc == 0) acc_phase ^= 1;
          }
          done = true;
          break;
        }
        ptx::tc_fence_after();
        CERB_PROF_T0(t_m2);
        if (leader) {
          const uint64_t a0 = a_d0 + static_cast<uint32_t>(stage) * stage_u;
          const uint64_t b0 = a0 + b_off_u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x (K = 16) per 64-channel slab
            if (!SPLIT) {
              ptx::umma_f16(tmem_d, a0 + 2 * k, b0 + 2 * k, idesc, (ks | k) != 0);
            } else {
              ptx::umma_f16(tmem_d + ((k & 1) ? kSplitMain1 : 0), a0 + 2 * k, b0 + 2 * k, idesc,
                            (ks != 0) || (k >= (split_merged ? 1 : 2)));
              if (!p.a_lo_zero)
                ptx::umma_f16(tmem_d + kSplitCross, a0 + lo_a_u + 2 * k, b0 + 2 * k, idesc, (ks | k) != 0);
              ptx::umma_f16(tmem_d + kSplitCross, a0 + 2 * k, b0 + lo_b_u + 2 * k, idesc,
                            p.a_lo_zero ? static_cast<int>((ks | k) != 0) : 1);
            }
          }
          ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot once the MMAs retire
        }
        __syncwarp();
        CERB_PROF_ADD(prof_c, t_m2);
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
      if (done) break;
      if (leader) ptx::umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
      __syncwarp();
      acc = (acc + 1) % kAccStages;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.prof != nullptr && lane == 0) {
      long long* o = p.prof + blockIdx.x * 16;
      o[1] = prof_a; o[2] = prof_b; o[3] = prof_c; o[8] = clock64() - t_all;
    }
  } else {
    // ------------------------------------------------------------ epilogue (2 groups x 4 warps)
    // Group g owns accumulator stage g and every second tile of this CTA, so two epilogues run
    // concurrently (layers with a short K loop are epilogue-bound with a single group).
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int grp = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int py = m >> p.bw_log2, px = m & bw_mask;
    const int acc = grp;
    uint32_t use = 0;
    long long prof_a = 0, prof_b = 0, prof_c = 0, prof_d = 0;
    for (int it = grp; grp < kAccStages; it += kAccStages, ++use) {
      const uint32_t acc_phase = use & 1;
      CERB_PROF_T0(t_e0);
      ptx::mbar_wait(&tfull_bar[acc], acc_phase, p.err_flag, 4);
      CERB_PROF_ADD(prof_a, t_e0);
      const int tile = s_ring[it & 15];
      if (tile < 0) break;
      const int nt = tile % p.n_ntiles;
      const int mt = tile / p.n_ntiles;
      const int tx = mt % p.tiles_x;
      const int ty = (mt / p.tiles_x) % p.tiles_y;
      const int img = mt / (p.tiles_x * p.tiles_y);
      const int ox = tx * BW + px, oy = ty * BH + py;
      const bool valid = (ox < p.W) && (oy < p.H);
      const size_t pix = (static_cast<size_t>(img) * p.H + oy) * p.W + ox;
      const int n0 = nt * p.BN;

      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccStride;
      CERB_PROF_T0(t_e1);
      if (mma_tail) {
        // ---- classification head on the tensor core: hidden = ReLU(acc + bias) -> fp16 tile in
        // shared memory (A operand), logits = hidden x W2^T as six N = 16 MMAs into TMEM columns
        // 96..111 of this accumulator stage (the main MMA, N = 96, never touches them).
        uint8_t* a2 = sA2 + grp * kHeadA2Bytes;
        const int sw = m & 7;
#pragma unroll
        for (int j = 0; j < 96; j += 32) {
          uint32_t r[32];
          ptx::tmem_ld32(taddr + j, r);
          ptx::tmem_ld_wait();
          float v[32];
          const float sc = p.acc_scale;
#pragma unroll
          for (int i = 0; i < 32; ++i)  // j and i are compile-time: the bias is a constant-bank operand
            v[i] = fmaf(__uint_as_float(r[i]), sc, p.head_hbias[j + i]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int cj = (j >> 3) + i;  // 16-byte chunk index inside the 96-channel row
            uint4 u;
            if (p.relu) {  // ReLU rides on the conversion
              u.x = pack_half2_relu(v[8 * i + 0], v[8 * i + 1]);
              u.y = pack_half2_relu(v[8 * i + 2], v[8 * i + 3]);
              u.z = pack_half2_relu(v[8 * i + 4], v[8 * i + 5]);
              u.w = pack_half2_relu(v[8 * i + 6], v[8 * i + 7]);
            } else {
              u.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
              u.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
              u.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
              u.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
            }
            *reinterpret_cast<uint4*>(a2 + (cj >> 3) * kATileBytes + m * 128 + (((cj & 7) ^ sw) << 4)) = u;
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);  // main accumulator drained
        CERB_PROF_ADD(prof_b, t_e1);
        CERB_PROF_T0(t_e2);
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(1 + grp, 128);
        if (q == 2) {  // first warp of the group (warps 2, 6, 10, 14 have warp & 3 == 2)
          if (ptx::elect_one()) {
            ptx::tc_fence_after();
            const uint32_t idesc2 = ptx::umma_idesc_f16(128, 16);
            const uint64_t ad = ptx::umma_desc_sw128(ptx::smem_u32(a2), 1024);
            const uint64_t bd = ptx::umma_desc_sw128(ptx::smem_u32(sB2), 1024);
            const uint32_t d2 = tmem_base + acc * kAccStride + 96;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
              const uint32_t a_off = static_cast<uint32_t>(((k >> 2) * kATileBytes + (k & 3) * 32) >> 4);
              const uint32_t b_off = static_cast<uint32_t>(((k >> 2) * 2048 + (k & 3) * 32) >> 4);
              ptx::umma_f16(d2, ad + a_off, bd + b_off, idesc2, k != 0);
            }
            ptx::umma_commit(&head_bar[grp]);
          }
          __syncwarp();
        }
        ptx::mbar_wait(&head_bar[grp], acc_phase, p.err_flag, 5);
        CERB_PROF_ADD(prof_c, t_e2);
        CERB_PROF_T0(t_e3);
        ptx::tc_fence_after();
        uint32_t r8[8];
        ptx::tmem_ld8(taddr + 96, r8);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        if (valid) {
          float hacc[kHeadMaxC];
#pragma unroll
          for (int c = 0; c < kHeadMaxC; ++c)
            hacc[c] = c < p.head_classes ? __uint_as_float(r8[c]) + p.head_obias[c] : 0.0f;
          const int y_off = static_cast<int>((p.H - p.oh) * 0.5), x_off = static_cast<int>((p.W - p.ow) * 0.5);
          const int cy = oy - y_off, cx = ox - x_off;
          float* dst = nullptr;
          if (cy >= 0 && cy < p.oh && cx >= 0 && cx < p.ow)
            dst = p.canvas + ((static_cast<size_t>(img) * p.oh + cy) * p.ow + cx) * p.canvas_c + p.canvas_coff;
          head_tail(hacc, p.head_classes, p.head_mode,
                    p.logits != nullptr ? p.logits + pix * p.head_classes : nullptr, dst);
        }
        CERB_PROF_ADD(prof_d, t_e3);
        continue;
      }
      float hacc[kHeadMaxC];
#pragma unroll
      for (int c = 0; c < kHeadMaxC; ++c) hacc[c] = 0.0f;
      for (int j = 0; j < p.BN; j += 32) {
        uint32_t r[32];
        float v[32];
        ptx::tmem_ld32(taddr + j, r);
        ptx::tmem_ld_wait();
        if (SPLIT) {
          uint32_t r1[32], r2[32];
          if (!split_merged) ptx::tmem_ld32(taddr + kSplitMain1 + j, r1);
          ptx::tmem_ld32(taddr + kSplitCross + j, r2);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float main1 = split_merged ? 0.0f : __uint_as_float(r1[i]);
            v[i] = ((__uint_as_float(r[i]) + main1) + __uint_as_float(r2[i])) * p.acc_scale;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.acc_scale;
        }
        if (valid) {
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + j);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = __ldg(b4 + i);
              v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
            }
          }
          if (p.res_hi != nullptr) {
            const size_t roff = pix * p.res_cs + p.res_coff + n0 + j;
            const uint4* rh = reinterpret_cast<const uint4*>(p.res_hi + roff);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 u = __ldg(rh + i);
              const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h[e]);
                v[8 * i + 2 * e] += f.x; v[8 * i + 2 * e + 1] += f.y;
              }
            }
            if (SPLIT) {
              const uint4* rl = reinterpret_cast<const uint4*>(p.res_lo + roff);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = __ldg(rl + i);
                const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __half22float2(h[e]);
                  v[8 * i + 2 * e] += f.x; v[8 * i + 2 * e + 1] += f.y;
                }
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
          }
          if (p.head_classes > 0) {
            // fused 1x1 96 -> C of the classification head: this thread owns the whole pixel
#pragma unroll
            for (int c = 0; c < kHeadMaxC; ++c) {
              if (c < p.head_classes) {
                const float4* w4 = reinterpret_cast<const float4*>(s_head_w + c * 96 + j);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 w = w4[i];
                  hacc[c] = fmaf(v[4 * i + 0], w.x, hacc[c]);
                  hacc[c] = fmaf(v[4 * i + 1], w.y, hacc[c]);
                  hacc[c] = fmaf(v[4 * i + 2], w.z, hacc[c]);
                  hacc[c] = fmaf(v[4 * i + 3], w.w, hacc[c]);
                }
              }
            }
            continue;
          }
          const size_t ooff = pix * p.out_cs + p.out_coff + n0 + j;
          uint4* oh = reinterpret_cast<uint4*>(p.out_hi + ooff);
          if (!SPLIT) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u;
              u.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
              u.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
              u.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
              u.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
              oh[i] = u;
            }
          } else {
            uint4* ol = reinterpret_cast<uint4*>(p.out_lo + ooff);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = v[8 * i + 2 * e], b = v[8 * i + 2 * e + 1];
                const __half2 h = __floats2half2_rn(a, b);
                const float2 hf = __half22float2(h);
                hi[e] = *reinterpret_cast<const uint32_t*>(&h);
                lo[e] = pack_half2(a - hf.x, b - hf.y);
              }
              oh[i] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              ol[i] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      CERB_PROF_ADD(prof_b, t_e1);
      if (p.head_classes > 0 && valid) {
#pragma unroll
        for (int c = 0; c < kHeadMaxC; ++c)
          if (c < p.head_classes) hacc[c] += __ldg(p.head_b + c);
        const int y_off = static_cast<int>((p.H - p.oh) * 0.5), x_off = static_cast<int>((p.W - p.ow) * 0.5);
        const int cy = oy - y_off, cx = ox - x_off;
        float* dst = nullptr;
        if (cy >= 0 && cy < p.oh && cx >= 0 && cx < p.ow)
          dst = p.canvas + ((static_cast<size_t>(img) * p.oh + cy) * p.ow + cx) * p.canvas_c + p.canvas_coff;
        head_tail(hacc, p.head_classes, p.head_mode,
                  p.logits != nullptr ? p.logits + pix * p.head_classes : nullptr, dst);
      }
    }
    if (p.prof != nullptr && (warp & 3) == 2 && lane == 0 && grp < 2) {
      long long* o = p.prof + blockIdx.x * 16 + (grp == 0 ? 4 : 10);
      o[0] = prof_a; o[1] = prof_b; o[2] = prof_c; o[3] = prof_d;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

void conv_tc_plan_pipeline(ConvKParams& p, bool split) {
  // epilogue-bound layers (one or two K steps per tile) get four accumulator stages / epilogue
  // groups when the tile is narrow enough for 4 x BN TMEM columns
  const bool short_k = p.n_taps * p.n_chunks <= 2;
  // split mode: two stages of 256 TMEM columns for narrow tiles and for wide tiles with a short K
  // loop. Long-K wide tiles keep ONE stage with the even / odd hi*hi accumulators: merging them
  // made the parity mode 3.4 % faster (15.70 -> 15.17 ms) but doubled its logit error (3.3e-4 ->
  // 6.5e-4 max-abs against the 1e-3 gate) - the tensor core's fp32 accumulation error grows with
  // the number of products per accumulator, and it is the dominant error of this mode.
  p.n_acc = split ? ((p.BN <= 64 || (p.BN <= 128 && short_k)) ? 2 : 1) : ((p.BN <= 128 && short_k) ? 4 : 2);
  p.stage_bytes = (split ? 2 : 1) * (kATileBytes + p.BN * 128);
  int budget = 192 * 1024;
  p.mma_tail = (!split && p.head_classes > 0 && p.n_acc == 4 && p.BN == 96) ? 1 : 0;
  if (p.mma_tail) budget = 224 * 1024 - kMaxAcc * kHeadA2Bytes - 4096;  // MMA tail staging
  int n = budget / p.stage_bytes;
  if (n > 8) n = 8;
  if (n < 2) n = 2;
  p.n_stages = n;
}

size_t conv_tc_smem_bytes(const ConvKParams& p) {
  // tiles + barriers (+ tmem holder) + slack for the manual 1024-byte alignment.
  // Always above half the SM's shared memory so that exactly one CTA (and one 512-column
  // TMEM allocation) lives on an SM at a time.
  size_t b = static_cast<size_t>(p.n_stages) * p.stage_bytes + 256 + 1024 + 3328;
  if (p.mma_tail) b += kMaxAcc * kHeadA2Bytes + 4096;
  if (b < 120 * 1024) b = 120 * 1024;
  return b;
}

cudaError_t conv_tc_launch(const ConvKParams& p, bool split, int num_sms, cudaStream_t stream, bool pdl) {
  static bool attr_set[4] = {false, false, false, false};
  const int variant = split ? (p.n_acc == 2 ? 3 : 2) : (p.n_acc == 4 ? 1 : 0);
  void (*kern)(ConvKParams) = variant == 3   ? conv_tc_kernel<true, 320>
                              : variant == 2 ? conv_tc_kernel<true, 192>
                              : variant == 1 ? conv_tc_kernel<false, 576>
                                             : conv_tc_kernel<false, 320>;
  if (!attr_set[variant]) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[variant] = true;
  }
  const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
  const int threads = 64 + 128 * p.n_acc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = conv_tc_smem_bytes(p);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

}  // namespace cerb
