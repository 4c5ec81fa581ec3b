// 3x3 stride-1 64->64 convolution specialised for the full-resolution layers.
//
// These layers (layer1 of the encoder and the u2/u1 stages of the five decoders:
// models/backbone/resnet.py:202, models/utils/net_layers.py:25-26) are 40 % of all FLOPs
//  and have N = 64, K = 576: the generic kernel re-reads every
// activation pixel nine times (once per tap) and the 72 KB of weights once per 128-pixel
// tile from L2, which caps it at the L2 bandwidth (~1/3 of the tensor pipe, measured). Here
//   * the 9 x [64 x 64] weights are loaded ONCE per persistent CTA and stay in shared memory,
//   * each 16 x 8 pixel tile loads its input halo once and the nine taps are nine shared-
//     memory matrix descriptors into that halo (no im2col, no re-read).
// Three halo layouts are implemented because the descriptor semantics for a start address
// that is not aligned to the 1024-byte swizzle period could only be settled on hardware:
//   mode 0: three x-shifted copies of an 18 x 8 pixel slab (every descriptor 1024-aligned);
//   mode 1: one 18 x 10 slab, taps are byte offsets into it (swizzle on absolute address bits);
//   mode 2: one 18 x 16 slab (1024-byte row-group pitch), x taps via the descriptor's
//           base_offset field.
// tests/test_gpu_conv.py runs all three against the fp32 reference; capi.cu uses the mode
// selected by cerb_ctx_set_option(ctx, "conv64_mode", m).
#include "conv64.cuh"
#include "head_tail.cuh"
#include "ptx.cuh"
#include "upadd_math.cuh"

namespace cerb {

namespace {

// In-kernel attribution (Conv64Params::prof != nullptr): cycles a role spends in each wait.
#define CERB_PROF_T0(var) const long long var = p.prof != nullptr ? clock64() : 0
#define CERB_PROF_ADD(acc, var) \
  do { if (p.prof != nullptr) acc += clock64() - var; } while (0)

constexpr int kWBytes = 9 * 64 * 128;  // resident weights: 9 taps x 64 rows x 128 B
constexpr int kTileW = 8, kTileH = 16;
constexpr int kTmemCols = 128;      // 2 accumulator stages x 64 columns
constexpr int kTmemColsTail = 512;  // + per epilogue group 96 (hidden) + 16 (logits) columns at 128 + 128 * group
constexpr int kTailW1Bytes = 96 * 128;   // head hidden weights [96][64] fp16
constexpr int kTailW2Bytes = 2 * 2048;   // head output weights [16][96] fp16, two 64-channel slabs
constexpr int kTailBytes = kTailW1Bytes + kTailW2Bytes + 2 * 128 * 128;  // + second hidden slab per group
constexpr int kOutTileBytes = 128 * 128;  // 128 pixels x 64 fp16 channels
constexpr int kUpGroupThreads = 128;
constexpr int kUpStageBytes = (18 * 10 + 10 * 6) * 128;  // skip halo + prev patch, per producer group

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kConv64ThreadsUp, 1)
conv64_kernel(const __grid_constant__ Conv64Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_stages = p.n_stages;
  const int stage_bytes = p.stage_bytes;

  uint8_t* sW = smem;
  uint8_t* sOut = smem + kWBytes;            // 2 x 16 KB output / residual staging (epilogue groups)
  uint8_t* sTail = sOut + 2 * kOutTileBytes;  // fused head: W1 | W2 | second hidden slab per group
  uint8_t* sW1 = sTail;
  uint8_t* sW2 = sW1 + kTailW1Bytes;
  uint8_t* sH1 = sW2 + kTailW2Bytes;
  uint8_t* sA = sTail + (p.has_tail ? kTailBytes : 0);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sA + n_stages * stage_bytes);
  uint64_t* empty_bar = full_bar + n_stages;
  uint64_t* tfull_bar = empty_bar + n_stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* w_bar = tempty_bar + 2;
  uint64_t* res_bar = w_bar + 1;
  uint64_t* tail_bar = res_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tail_bar + 2);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.in_map);
    ptx::prefetch_tmap(&p.w_map);
    ptx::prefetch_tmap(&p.out_map);
    if (p.has_tail) ptx::prefetch_tmap(&p.w1_map);
    if (p.has_res) ptx::prefetch_tmap(&p.res_map);
    for (int s = 0; s < n_stages; ++s) {
      // fused upsample+add: the 128 threads of one producer group arrive instead of a TMA
      ptx::mbar_init(&full_bar[s], p.up_prev != nullptr ? kUpGroupThreads : 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull_bar[s], 1);
      ptx::mbar_init(&tempty_bar[s], 4);
    }
    ptx::mbar_init(w_bar, 1);
    ptx::mbar_init(&res_bar[0], 1);
    ptx::mbar_init(&res_bar[1], 1);
    ptx::mbar_init(&tail_bar[0], 1);
    ptx::mbar_init(&tail_bar[1], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_holder, p.has_tail ? kTmemColsTail : kTmemCols);
    ptx::tmem_relinquish();
  }
  if (p.has_tail) {
    // W2[n][k] = fp16(tail_w2[n][k]) for n < C, zero rows up to N = 16 (K-major, 128-byte swizzle)
    for (int i = threadIdx.x; i < 16 * 96; i += blockDim.x) {
      const int n = i / 96, k = i - n * 96;
      const float w = n < p.tail_classes ? p.tail_w2[n * 96 + k] : 0.0f;
      const int slab = k >> 6, kk = k & 63;
      *reinterpret_cast<__half*>(sW2 + slab * 2048 + n * 128 + (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2) =
          __float2half_rn(w);
    }
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  ptx::grid_dep_launch();  // the next kernel of the stream may start its own prologue

  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // The whole warp walks the loop and ONE elected lane issues: a branch on `lane == 0` makes
    // ptxas wrap every uniform-datapath instruction (UTMALDG, UTCHMMA) in an ELECT/BRA.U.ANY
    // serialisation loop, which costs ~78 cycles per MMA (tools/umma_bench.cu, profiles/).
    const bool leader = ptx::elect_one() != 0;
    if (leader) {
      // resident weights: one 2-D box per tap
      ptx::mbar_arrive_expect_tx(w_bar, p.n_taps * 8192 + (p.has_tail ? kTailW1Bytes : 0));
#pragma unroll
      for (int t = 0; t < 9; ++t)
        if (t < p.n_taps) ptx::tma_load_2d(sW + t * 8192, &p.w_map, w_bar, t * 64, 0);
      if (p.has_tail) ptx::tma_load_2d(sW1, &p.w1_map, w_bar, 0, 0);
    }
    ptx::grid_dep_wait();  // weights do not depend on the previous kernel, activations do
    int stage = 0;
    uint32_t phase = 0;
    long long prof_a = 0;
    for (int tile = blockIdx.x; p.up_prev == nullptr && tile < p.n_tiles; tile += gridDim.x) {
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int x0 = tx * kTileW - p.halo_x, y0 = ty * kTileH - p.halo_y;
      CERB_PROF_T0(t_pe);
      ptx::mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 21);
      CERB_PROF_ADD(prof_a, t_pe);
      uint8_t* dst = sA + stage * stage_bytes;
      if (leader) {
        ptx::mbar_arrive_expect_tx(&full_bar[stage], p.tx_bytes);
        if (p.mode == 0) {
          for (int s = 0; s < 3; ++s)
            ptx::tma_load_4d(dst + s * p.copy_bytes, &p.in_map, &full_bar[stage], 0, x0 + s, y0, img);
        } else {
          ptx::tma_load_4d(dst, &p.in_map, &full_bar[stage], 0, x0, y0, img);
        }
      }
      __syncwarp();
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    if (p.prof != nullptr && lane == 0) p.prof[blockIdx.x * 16 + 0] = prof_a;
  } else if (warp == 1) {
    const bool leader = ptx::elect_one() != 0;
    const uint32_t idesc = ptx::umma_idesc_f16(128, 64);
    ptx::mbar_wait(w_bar, 0, p.err_flag, 22);
    ptx::tc_fence_after();
    const uint64_t a_d0 = ptx::umma_desc_sw128(ptx::smem_u32(sA), p.sbo_bytes, 0);
    const uint64_t b_d0 = ptx::umma_desc_sw128(ptx::smem_u32(sW), 1024, 0);
    const uint32_t a_hi = static_cast<uint32_t>(a_d0 >> 32), a_lo0 = static_cast<uint32_t>(a_d0);
    const uint32_t b_hi = static_cast<uint32_t>(b_d0 >> 32), b_lo0 = static_cast<uint32_t>(b_d0);
    uint32_t tap_off[9];  // byte offset of tap (r, s) inside a stage, in 16-byte units
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int r = t / 3, s = t - 3 * r;
      tap_off[t] = static_cast<uint32_t>(
          (p.mode == 4   ? t * (kTileW * 128)  // stem: seven vertical taps, no horizontal shift
           : p.mode == 0 ? s * p.copy_bytes + r * (kTileW * 128)
                         : (r * p.pitch_px + s) * 128) >> 4);
    }
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long prof_a = 0, prof_b = 0, prof_c = 0;
    CERB_PROF_T0(t_all);
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      CERB_PROF_T0(t_m0);
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1, p.err_flag, 23);
      CERB_PROF_ADD(prof_a, t_m0);
      CERB_PROF_T0(t_m1);
      ptx::mbar_wait(&full_bar[stage], phase, p.err_flag, 24);
      CERB_PROF_ADD(prof_b, t_m1);
      ptx::tc_fence_after();
      CERB_PROF_T0(t_m2);
      if (leader) {
        const uint32_t tmem_d = tmem_base + acc * 64;
        // descriptor low words: only the 14-bit start-address field changes between MMAs
        const uint32_t a_lo = a_lo0 + static_cast<uint32_t>((stage * stage_bytes) >> 4);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          if (t >= p.n_taps) break;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + tap_off[t] + 2 * k);
            const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (b_lo0 + t * 512 + 2 * k);
            ptx::umma_f16(tmem_d, ad, bd, idesc, (t | k) != 0);
          }
        }
        ptx::umma_commit(&empty_bar[stage]);
        ptx::umma_commit(&tfull_bar[acc]);
      }
      __syncwarp();
      CERB_PROF_ADD(prof_c, t_m2);
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.prof != nullptr && lane == 0) {
      long long* o = p.prof + blockIdx.x * 16;
      o[1] = prof_a; o[2] = prof_b; o[3] = prof_c; o[8] = clock64() - t_all;
    }
  } else if (warp >= 6 && p.up_prev != nullptr) {
    // ---------------------------------------------------- fused `skip + bilinear_x2(prev)` producer
    // (models/net_desc.py:185-188). Two groups of four warps build alternate tiles' halos
    // directly in the swizzled shared-memory layout the MMA descriptors expect, so the summed
    // tensor never exists in HBM. Arithmetic and fp16 rounding are those of upadd_kernel.
    ptx::grid_dep_wait();
    const int grp = (warp - 6) >> 2;
    const int gtid = threadIdx.x - (6 + 4 * grp) * 32;
    const int PH = p.H >> 1, PW = p.W >> 1;
    // per-group staging: the raw skip halo (18x10 px) and the prev patch (10x6 px), fetched with
    // cp.async so that every 16-byte request of a tile is in flight at once
    uint8_t* stS = sA + n_stages * stage_bytes + 256 + grp * kUpStageBytes;
    uint8_t* stP = stS + 18 * 10 * 128;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const int stage = it % n_stages;
      const uint32_t phase = (it / n_stages) & 1;
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int x0 = tx * kTileW - 1, y0 = ty * kTileH - 1;
      const int pby = ty * (kTileH / 2) - 1, pbx = tx * (kTileW / 2) - 1;  // prev patch origin
      const __half* skip_img = p.up_skip + static_cast<size_t>(img) * p.H * p.W * p.up_skip_cs;
      const __half* prev_img = p.up_prev + static_cast<size_t>(img) * PH * PW * p.up_prev_cs;
      constexpr int kSkipTasks = 18 * 10 * 8;  // halo pixels x 16-byte channel chunks
      constexpr int kPrevTasks = 10 * 6 * 8;
      for (int t = gtid; t < kSkipTasks; t += kUpGroupThreads) {
        const int h = t >> 3, c = t & 7;
        const int hy = h / 10, hx = h - hy * 10;
        const int Y = y0 + hy, X = x0 + hx;
        if (Y >= 0 && Y < p.H && X >= 0 && X < p.W)
          ptx::cp_async16(stS + t * 16, skip_img + (static_cast<size_t>(Y) * p.W + X) * p.up_skip_cs + c * 8);
      }
      for (int t = gtid; t < kPrevTasks; t += kUpGroupThreads) {
        const int h = t >> 3, c = t & 7;
        const int ly_ = h / 6, lx_ = h - ly_ * 6;
        const int py = min(max(pby + ly_, 0), PH - 1), px = min(max(pbx + lx_, 0), PW - 1);
        ptx::cp_async16(stP + t * 16, prev_img + (static_cast<size_t>(py) * PW + px) * p.up_prev_cs + c * 8);
      }
      ptx::cp_async_wait_all();
      ptx::named_bar_sync(1 + grp, kUpGroupThreads);  // the whole group's copies have landed
      ptx::mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 26);
      uint8_t* dst = sA + stage * stage_bytes;
      for (int t = gtid; t < kSkipTasks; t += kUpGroupThreads) {
        const int h = t >> 3, c = t & 7;
        const int hy = h / 10, hx = h - hy * 10;
        const int Y = y0 + hy, X = x0 + hx;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (Y >= 0 && Y < p.H && X >= 0 && X < p.W) {
          const float sy = fmaxf((Y + 0.5f) * 0.5f - 0.5f, 0.0f);
          const float sx = fmaxf((X + 0.5f) * 0.5f - 0.5f, 0.0f);
          const int py0 = static_cast<int>(sy), px0 = static_cast<int>(sx);
          const int py1 = min(py0 + 1, PH - 1), px1 = min(px0 + 1, PW - 1);
          const float ly = sy - py0, lx = sx - px0;
          const uint4 sk = *reinterpret_cast<const uint4*>(stS + t * 16);
          const uint4 q00 = *reinterpret_cast<const uint4*>(stP + (((py0 - pby) * 6 + (px0 - pbx)) * 8 + c) * 16);
          const uint4 q01 = *reinterpret_cast<const uint4*>(stP + (((py0 - pby) * 6 + (px1 - pbx)) * 8 + c) * 16);
          const uint4 q10 = *reinterpret_cast<const uint4*>(stP + (((py1 - pby) * 6 + (px0 - pbx)) * 8 + c) * 16);
          const uint4 q11 = *reinterpret_cast<const uint4*>(stP + (((py1 - pby) * 6 + (px1 - pbx)) * 8 + c) * 16);
          o = upadd_pixel8_h2(q00, q01, q10, q11, sk, ly, lx);  // the arithmetic of the stand-alone pass
        }
        // 128-byte swizzle: 16-byte chunk c of row h lives at chunk (c ^ (h & 7))
        *reinterpret_cast<uint4*>(dst + h * 128 + ((c ^ (h & 7)) << 4)) = o;
      }
      ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor-core proxy
      ptx::mbar_arrive(&full_bar[stage]);
      ptx::named_bar_sync(1 + grp, kUpGroupThreads);  // staging may be overwritten by the next tile
    }
  } else {
    // epilogue: warps 2-5 drain accumulator stage 0 (even tiles of this CTA); without the fused
    // producer, warps 6-9 drain stage 1 (odd tiles) so that two tiles' epilogues overlap.
    // A thread owns one pixel (= one TMEM lane = one 128-byte row of the output tile). Writing
    // that row straight to global memory makes every warp store touch 32 different lines; the
    // rows go to a swizzled shared-memory tile instead and ONE TMA store per tile writes whole
    // lines (and clips partial tiles at the image border). The residual tile arrives in the same
    // staging buffer by TMA while the MMAs of the tile are still running.
    ptx::grid_dep_wait();  // residual reads / output writes must follow the previous kernel
    const int egrp = (warp - 2) >> 2;
    const bool two_groups = p.up_prev == nullptr;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int gtid = static_cast<int>(threadIdx.x) - (2 + 4 * egrp) * 32;
    uint8_t* sO = sOut + egrp * kOutTileBytes;
    uint8_t* my_row = sO + m * 128;
    const int sw = m & 7;  // 128-byte swizzle: 16-byte chunk c of row m lives at chunk c ^ (m & 7)
    // first warp of the group issues the TMA traffic through an elected lane (uniform branch:
    // no ELECT/BRA.U.ANY serialisation loop around UTMALDG / UTMASTG)
    const bool store_warp = q == 2;
    uint32_t res_phase = 0, tail_phase = 0;
    int it = 0;
    long long prof_a = 0, prof_b = 0, prof_c = 0, prof_d = 0, prof_e = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      if (two_groups && (it & 1) != egrp) continue;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int x0 = tx * kTileW, y0 = ty * kTileH;
      CERB_PROF_T0(t_e0);
      if (!p.has_tail && store_warp && ptx::elect_one()) {
        ptx::bulk_wait_read<0>();  // the previous store of this group has drained the staging tile
        if (p.has_res) {
          ptx::mbar_arrive_expect_tx(&res_bar[egrp], kOutTileBytes);
          ptx::tma_load_4d(sO, &p.res_map, &res_bar[egrp], 0, x0, y0, img);
        }
      }
      CERB_PROF_ADD(prof_e, t_e0);
      CERB_PROF_T0(t_e1);
      ptx::mbar_wait(&tfull_bar[acc], acc_phase, p.err_flag, 25);
      CERB_PROF_ADD(prof_a, t_e1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 64;
      uint32_t r0[32], r1[32];
      ptx::tmem_ld32(taddr, r0);
      ptx::tmem_ld32(taddr + 32, r1);
      ptx::tmem_ld_wait();
      // the accumulator is in registers: release it before the rest of the epilogue
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      if (p.debug & 4) continue;
      CERB_PROF_T0(t_e2);
      if (p.has_res) {
        ptx::mbar_wait(&res_bar[egrp], res_phase, p.err_flag, 27);
        res_phase ^= 1;
      } else {
        ptx::named_bar_sync(3 + egrp, 128);  // thread 0 has seen the staging tile drained
      }
      CERB_PROF_ADD(prof_b, t_e2);
      CERB_PROF_T0(t_e3);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i)
          v[i] = __uint_as_float(half == 0 ? r0[i] : r1[i]) * p.acc_scale;
        const int j = half * 32;
        if (p.bias != nullptr) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + j);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b = __ldg(b4 + i);
            v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
          }
        }
        if (p.has_res) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 u = *reinterpret_cast<const uint4*>(my_row + (((half * 4 + i) ^ sw) << 4));
            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(h[e]);
              v[8 * i + 2 * e] += f.x; v[8 * i + 2 * e + 1] += f.y;
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
          u.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
          u.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
          u.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
          *reinterpret_cast<uint4*>(my_row + (((half * 4 + i) ^ sw) << 4)) = u;
        }
      }
      CERB_PROF_ADD(prof_c, t_e3);
      CERB_PROF_T0(t_e4);
      ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA store / the MMA
      ptx::named_bar_sync(3 + egrp, 128);
      if (!p.has_tail) {
        if (store_warp && !(p.debug & 1) && ptx::elect_one()) {
          ptx::tma_store_4d(&p.out_map, sO, 0, x0, y0, img);
          ptx::bulk_commit_group();
        }
        CERB_PROF_ADD(prof_d, t_e4);
        continue;
      }
      // ---- fused classification head (models/utils/net_layers.py:31-38, run_desc.py:451-491):
      // the tile just written is the A operand of the hidden 1x1 64 -> 96 (four N = 96 MMAs into
      // this group's TMEM columns), whose ReLU output is the A operand of the 1x1 96 -> C (six
      // N = 16 MMAs); the 64- and 96-channel tensors never exist in HBM.
      const uint32_t d2 = tmem_base + 128 + egrp * 128;
      const uint32_t t2addr = d2 + (static_cast<uint32_t>(q * 32) << 16);
      if (store_warp) {
        if (ptx::elect_one()) {
          ptx::tc_fence_after();
          const uint32_t idesc1 = ptx::umma_idesc_f16(128, 96);
          const uint64_t ad = ptx::umma_desc_sw128(ptx::smem_u32(sO), 1024);
          const uint64_t bd = ptx::umma_desc_sw128(ptx::smem_u32(sW1), 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_f16(d2, ad + 2 * k, bd + 2 * k, idesc1, k != 0);
          ptx::umma_commit(&tail_bar[egrp]);
        }
        __syncwarp();
      }
      ptx::mbar_wait(&tail_bar[egrp], tail_phase, p.err_flag, 28);
      tail_phase ^= 1;
      ptx::tc_fence_after();
      uint8_t* sH = sH1 + egrp * kOutTileBytes;  // channels 64..95 of the hidden tile
#pragma unroll
      for (int j = 0; j < 96; j += 32) {
        uint32_t r[32];
        ptx::tmem_ld32(t2addr + j, r);
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.tail_scale;
        const float4* b4 = reinterpret_cast<const float4*>(p.tail_b1 + j);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = __ldg(b4 + i);
          v[4 * i + 0] = fmaxf(v[4 * i + 0] + b.x, 0.0f);
          v[4 * i + 1] = fmaxf(v[4 * i + 1] + b.y, 0.0f);
          v[4 * i + 2] = fmaxf(v[4 * i + 2] + b.z, 0.0f);
          v[4 * i + 3] = fmaxf(v[4 * i + 3] + b.w, 0.0f);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cj = (j >> 3) + i;  // 16-byte chunk inside the 96-channel row
          uint4 u;
          u.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
          u.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
          u.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
          u.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
          uint8_t* base = cj < 8 ? sO : sH;
          *reinterpret_cast<uint4*>(base + m * 128 + (((cj & 7) ^ sw) << 4)) = u;
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(3 + egrp, 128);
      if (store_warp) {
        if (ptx::elect_one()) {
          ptx::tc_fence_after();
          const uint32_t idesc2 = ptx::umma_idesc_f16(128, 16);
          const uint64_t a0 = ptx::umma_desc_sw128(ptx::smem_u32(sO), 1024);
          const uint64_t a1 = ptx::umma_desc_sw128(ptx::smem_u32(sH), 1024);
          const uint64_t bd = ptx::umma_desc_sw128(ptx::smem_u32(sW2), 1024);
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const uint64_t ad = (k < 4 ? a0 : a1) + 2 * (k & 3);
            ptx::umma_f16(d2 + 96, ad, bd + static_cast<uint32_t>(((k >> 2) * 2048 + (k & 3) * 32) >> 4), idesc2, k != 0);
          }
          ptx::umma_commit(&tail_bar[egrp]);
        }
        __syncwarp();
      }
      ptx::mbar_wait(&tail_bar[egrp], tail_phase, p.err_flag, 29);
      tail_phase ^= 1;
      ptx::tc_fence_after();
      uint32_t r8[8];
      ptx::tmem_ld8(t2addr + 96, r8);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      const int ox = x0 + (m & 7), oy = y0 + (m >> 3);
      if (ox < p.W && oy < p.H) {
        float hacc[kHeadMaxC];
#pragma unroll
        for (int c = 0; c < kHeadMaxC; ++c)
          hacc[c] = c < p.tail_classes ? __uint_as_float(r8[c]) + __ldg(p.tail_b2 + c) : 0.0f;
        const int y_off = static_cast<int>((p.H - p.oh) * 0.5), x_off = static_cast<int>((p.W - p.ow) * 0.5);
        const int cy = oy - y_off, cx = ox - x_off;
        float* dst = nullptr;
        if (cy >= 0 && cy < p.oh && cx >= 0 && cx < p.ow)
          dst = p.canvas + ((static_cast<size_t>(img) * p.oh + cy) * p.ow + cx) * p.canvas_c + p.canvas_coff;
        const size_t pix = (static_cast<size_t>(img) * p.H + oy) * p.W + ox;
        head_tail(hacc, p.tail_classes, p.tail_mode,
                  p.logits != nullptr ? p.logits + pix * p.tail_classes : nullptr, dst);
      }
      CERB_PROF_ADD(prof_d, t_e4);
    }
    if (store_warp) ptx::bulk_wait_all<0>();  // only the elected lane has groups pending
    if (p.prof != nullptr && gtid == 0) {
      long long* o = p.prof + blockIdx.x * 16 + (egrp == 0 ? 4 : 10);
      o[0] = prof_a; o[1] = prof_b; o[2] = prof_c; o[3] = prof_d; o[5] = prof_e;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.has_tail ? kTmemColsTail : kTmemCols);
  }
}

}  // namespace

void conv64_plan(Conv64Params& p) {
  // halo rows = kTileH + 2 = 18
  if (p.mode == 0) {
    p.pitch_px = kTileW;
    p.copy_bytes = 18 * kTileW * 128;  // 18 KB, a multiple of 1024
    p.stage_bytes = 3 * p.copy_bytes;
    p.tx_bytes = 3 * p.copy_bytes;
    p.sbo_bytes = 1024;
  } else if (p.mode == 4) {
    // 7x7 stem (models/backbone/resnet.py:195-200) through the overlapping-window view of the
    // PREP tensor: a "pixel" of the halo is an 8-pixel x 8-channel window (one filter row = one
    // K chunk), so only the seven vertical taps shift the view: 22 rows x 8 windows, all
    // descriptors 1024-byte aligned
    p.pitch_px = kTileW;
    p.copy_bytes = (kTileH + 6) * kTileW * 128;  // 22528
    p.stage_bytes = p.copy_bytes;
    p.tx_bytes = p.copy_bytes;
    p.sbo_bytes = 1024;
  } else if (p.mode == 1) {
    p.pitch_px = kTileW + 2;
    p.copy_bytes = 18 * p.pitch_px * 128;  // 23040
    p.stage_bytes = 23 * 1024;             // padded to the swizzle period
    p.tx_bytes = p.copy_bytes;
    p.sbo_bytes = p.pitch_px * 128;
  } else {
    p.pitch_px = 16;
    p.copy_bytes = 18 * 16 * 128;
    p.stage_bytes = p.copy_bytes;
    p.tx_bytes = p.copy_bytes;
    p.sbo_bytes = 16 * 128;
  }
  int n = (224 * 1024 - kWBytes - 2 * kOutTileBytes - (p.has_tail ? kTailBytes : 0) -
           (p.up_prev != nullptr ? 2 * kUpStageBytes : 0)) / p.stage_bytes;
  if (n > 4) n = 4;
  if (n < 2) n = 2;
  p.n_stages = n;
}

int conv64_box_w(int mode) { return (mode == 0 || mode == 4) ? kTileW : (mode == 1 ? kTileW + 2 : 16); }
int conv64_tile_w() { return kTileW; }
int conv64_tile_h() { return kTileH; }

size_t conv64_smem_bytes(const Conv64Params& p) {
  return static_cast<size_t>(kWBytes) + 2 * kOutTileBytes + (p.has_tail ? kTailBytes : 0) +
         static_cast<size_t>(p.n_stages) * p.stage_bytes + 256 + 1024 +
         (p.up_prev != nullptr ? 2 * kUpStageBytes : 0);
}

cudaError_t conv64_launch(const Conv64Params& p, int num_sms, cudaStream_t stream, bool pdl) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
  const int threads = p.up_prev != nullptr ? kConv64ThreadsUp : kConv64Threads;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = conv64_smem_bytes(p);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, conv64_kernel, p);
}

}  // namespace cerb
