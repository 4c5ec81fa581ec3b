// 3x3 stride-1 64->64 convolution, even/odd formulation (conv64_mode 3).
//
// conv64.cu is bound by shared-memory bandwidth: a 128 x 64 x 16 MMA reads 4 KB of A and 2 KB of
// B for 32 cycles of math, i.e. 48 cycles at 128 B/clk - two thirds of the
// tensor pipe at best. N cannot grow with 64 output channels ... unless the second half of N is a
// SECOND OUTPUT PIXEL. Here a TMEM lane is the pixel pair (y, 2j), (y, 2j+1) of a 16 x 16 region:
// M = 16 rows x 8 even columns, N = [64 channels of the even pixel | 64 channels of the odd one].
// For an A view "input pixel (y + r - 1, 2j + a - 1)", a = 0..3, the even output needs tap
// (r, s = a) and the odd output, which sits one column to the right, tap (r, s = a - 1):
//     a = 0 : B = W(r,0)            -> N =  64 MMA into the even half
//     a = 1 : B = [W(r,1); W(r,0)]  -> N = 128 MMA
//     a = 2 : B = [W(r,2); W(r,1)]  -> N = 128 MMA
//     a = 3 : B = W(r,2)            -> N =  64 MMA into the odd half
// 24 N=128 + 24 N=64 MMAs per 256 output pixels: 2688 cycles at the shared-memory floor instead
// of 2 x 36 x 48 = 3456. The stride-2 pixel access of the A views comes for free from TMA: the
// 18 x 18 halo is loaded as two planes (odd and even input columns, tensor-map element stride 2),
// so a view is again "plane + constant offset". The six stacked weight tiles are assembled in
// shared memory from per-tap TMA boxes and stay resident. Epilogue: a thread owns two adjacent
// pixels = 256 contiguous bytes of the NHWC output; rows go through a swizzled staging tile and
// one TMA store per region (which also clips partial regions); the residual arrives by TMA.
//
// conv64x_kernel<true> (Conv64xParams::fuse_up): the input is skip + bilinear_x2(low)
// (models/net_desc.py:185-188). The halo planes are loaded from the SKIP tensor, a third TMA box
// brings the 10 x 10 low-resolution pixels under the halo, and six extra warps add the
// interpolated term to the planes in shared memory before the MMA warp may read them
// (upadd_math.cuh: the arithmetic of the stand-alone pass, so both round identically).
//
// Both variants: while the stage a halo will land in is still busy, the producer asks L2 for its
// boxes (cp.async.bulk.prefetch.tensor) - with two stages the region period of a latency-bound
// launch is load latency (+ fix-up), and the later load then hits L2.
#include "conv64x.cuh"
#include "ptx.cuh"
#include "upadd_math.cuh"

namespace cerb {

namespace {

#define CERB_PROF_T0(var) const long long var = p.prof != nullptr ? clock64() : 0
#define CERB_PROF_ADD(acc, var) \
  do { if (p.prof != nullptr) acc += clock64() - var; } while (0)

constexpr int kTapBytes = 64 * 128;            // one tap: 64 rows x 128 B
constexpr int kWBytes = 6 * 2 * kTapBytes;     // six stacked [128 x 64] tiles
constexpr int kPlanePx = 18 * 9;               // 18 rows x 9 columns of one parity
constexpr int kPlaneTx = kPlanePx * 128;       // 20736 bytes delivered per plane
constexpr int kPlaneBytes = 21 * 1024;         // padded to the swizzle period
constexpr int kStageBytes = 2 * kPlaneBytes;   // odd plane | even plane
constexpr int kOutBytes = 256 * 128;           // staging: 16 x 16 pixels x 64 fp16 channels
constexpr int kAccStages = 3;                  // accumulator stages of 128 columns
constexpr int kTmemCols = 512;                 // power of two >= 3 x 128
// fused upsample+add: the 10 x 10 low-resolution pixels under a halo (one buffer: the fix-up of a
// region is over long before the next halo can be requested)
constexpr int kLowPx = 10;
constexpr int kLowTx = kLowPx * kLowPx * 128;  // 12800
constexpr int kLowBytes = 13 * 1024;
constexpr int kFixWarps = 6;
constexpr int kFixItems = 9 * 9 * 8;           // 2x2-pixel blocks of the 18 x 18 halo x 8 channel groups

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <bool FUSE>
__global__ void __launch_bounds__(FUSE ? kConv64xFuseThreads : kConv64xThreads, 1)
conv64x_kernel(const __grid_constant__ Conv64xParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_stages = p.n_stages;

  uint8_t* sW = smem;                   // tile (r, q): sW + (2r + q) * 16 KB; q = 0: [W(r,1); W(r,0)], q = 1: [W(r,2); W(r,1)]
  uint8_t* sOut = sW + kWBytes;         // output / residual staging
  uint8_t* sA = sOut + kOutBytes;       // halo stages
  uint8_t* sLow = sA + n_stages * kStageBytes;  // FUSE only
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sLow + (FUSE ? kLowBytes : 0));
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* tfull_bar = empty_bar + 4;
  uint64_t* tempty_bar = tfull_bar + kAccStages;
  uint64_t* w_bar = tempty_bar + kAccStages;
  uint64_t* res_bar = w_bar + 1;
  uint64_t* fixed_bar = res_bar + 1;   // FUSE: "halo stage s holds skip + up(low)" (one arrival per fix-up warp)
  uint64_t* low_empty = fixed_bar + 4;  // FUSE: the fix-up warps are done with the low buffer
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(low_empty + 1);
  // region ids handed from the producer to the MMA / epilogue warps (dynamic scheduling)
  volatile int* s_ring = reinterpret_cast<volatile int*>(tmem_holder + 2);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.in_map);
    ptx::prefetch_tmap(&p.w_map);
    ptx::prefetch_tmap(&p.out_map);
    if (p.has_res) ptx::prefetch_tmap(&p.res_map);
    if (FUSE) ptx::prefetch_tmap(&p.low_map);
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
      ptx::mbar_init(&fixed_bar[s], kFixWarps);
    }
    ptx::mbar_init(low_empty, kFixWarps);
    for (int s = 0; s < kAccStages; ++s) {
      ptx::mbar_init(&tfull_bar[s], 1);
      ptx::mbar_init(&tempty_bar[s], 8);  // one arrival per epilogue warp
    }
    ptx::mbar_init(w_bar, 1);
    ptx::mbar_init(res_bar, 1);
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

  const int regions_per_img = p.regions_x * p.regions_y;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    const bool leader = ptx::elect_one() != 0;
    if (leader) {
      ptx::mbar_arrive_expect_tx(w_bar, kWBytes);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        uint8_t* t0 = sW + (2 * r) * 2 * kTapBytes;
        uint8_t* t1 = t0 + 2 * kTapBytes;
        ptx::tma_load_2d(t0, &p.w_map, w_bar, (3 * r + 1) * 64, 0);              // W(r,1)
        ptx::tma_load_2d(t0 + kTapBytes, &p.w_map, w_bar, (3 * r + 0) * 64, 0);  // W(r,0)
        ptx::tma_load_2d(t1, &p.w_map, w_bar, (3 * r + 2) * 64, 0);              // W(r,2)
        ptx::tma_load_2d(t1 + kTapBytes, &p.w_map, w_bar, (3 * r + 1) * 64, 0);  // W(r,1)
      }
    }
    ptx::grid_dep_wait();  // weights do not depend on the previous kernel, activations do
    long long prof_a = 0;
    int stage = 0;
    uint32_t phase = 0, low_phase = 0;
    const int leader_lane = __ffs(__ballot_sync(0xffffffffu, leader)) - 1;
    // Region ids come from a global counter (dynamic scheduling): a CTA that starts late - its SM
    // was busy with a block of another stream - simply finds less work, instead of forcing a
    // second wave as a static blockIdx-strided split would. The id travels to the other warps
    // through s_ring (slot i & 7), published before the arrive on full_bar; -1 ends the kernel.
    for (int i = 0;; ++i) {
      int rg = 0;
      if (p.tile_counter != nullptr) {
        if (leader) rg = atomicAdd(p.tile_counter, 1);
        rg = __shfl_sync(0xffffffffu, rg, leader_lane);
      } else {
        rg = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      }
      const bool done = rg >= p.n_regions;
      const int img = rg / regions_per_img;
      const int rem = rg - img * regions_per_img;
      const int ry = rem / p.regions_x, rx = rem - ry * p.regions_x;
      const int x0 = rx * 16, y0 = ry * 16 - 1;
      // The stage this halo will land in is still being read by the MMAs of the region before
      // last: ask L2 for the boxes now, so that the load issued after the wait below is an L2 hit
      // (with two stages the region period is load latency (+ fix-up), not MMA time).
      if (leader && !done) {
        ptx::tma_prefetch_4d(&p.in_map, 0, x0 - 1, y0, img);
        ptx::tma_prefetch_4d(&p.in_map, 0, x0, y0, img);
        if (FUSE) ptx::tma_prefetch_4d(&p.low_map, 0, (x0 >> 1) - 1, ry * 8 - 1, img);
      }
      CERB_PROF_T0(t_p);
      ptx::mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 41);
      if (FUSE) {
        ptx::mbar_wait(low_empty, low_phase ^ 1, p.err_flag, 48);
        low_phase ^= 1;
      }
      CERB_PROF_ADD(prof_a, t_p);
      if (leader) {
        s_ring[i & 7] = done ? -1 : rg;
        if (done) {
          ptx::mbar_arrive(&full_bar[stage]);
        } else {
          uint8_t* dst = sA + stage * kStageBytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * kPlaneTx + (FUSE ? kLowTx : 0));
          ptx::tma_load_4d(dst, &p.in_map, &full_bar[stage], 0, x0 - 1, y0, img);            // odd columns
          ptx::tma_load_4d(dst + kPlaneBytes, &p.in_map, &full_bar[stage], 0, x0, y0, img);  // even columns
          if (FUSE) ptx::tma_load_4d(sLow, &p.low_map, &full_bar[stage], 0, (x0 >> 1) - 1, ry * 8 - 1, img);
        }
      }
      __syncwarp();
      if (done) break;
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    if (p.prof != nullptr && lane == 0) p.prof[blockIdx.x * 16 + 0] = prof_a;
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = ptx::elect_one() != 0;
    const uint32_t idesc128 = ptx::umma_idesc_f16(128, 128);
    const uint32_t idesc64 = ptx::umma_idesc_f16(128, 64);
    ptx::mbar_wait(w_bar, 0, p.err_flag, 42);
    ptx::tc_fence_after();
    // A: 8-row groups are the 8 column pairs of one image row (9-pixel pitch inside a plane)
    const uint64_t a_d0 = ptx::umma_desc_sw128(ptx::smem_u32(sA), 9 * 128);
    const uint64_t b_d0 = ptx::umma_desc_sw128(ptx::smem_u32(sW), 1024);
    long long prof_a = 0, prof_b = 0, prof_c = 0;
    CERB_PROF_T0(t_all);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int i = 0;; ++i) {
      CERB_PROF_T0(t_m0);
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1, p.err_flag, 43);
      CERB_PROF_ADD(prof_a, t_m0);
      CERB_PROF_T0(t_m1);
      ptx::mbar_wait(FUSE ? &fixed_bar[stage] : &full_bar[stage], phase, p.err_flag, 44);
      CERB_PROF_ADD(prof_b, t_m1);
      if (s_ring[i & 7] < 0) {  // no more regions: pass the end marker on to the epilogue
        if (leader) ptx::mbar_arrive(&tfull_bar[acc]);
        __syncwarp();
        break;
      }
      ptx::tc_fence_after();
      CERB_PROF_T0(t_m2);
      if (leader) {
        const uint32_t d_even = tmem_base + acc * 128;
        const uint32_t d_odd = d_even + 64;
        const uint64_t a_st = a_d0 + static_cast<uint32_t>((stage * kStageBytes) >> 4);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int ai = 0; ai < 4; ++ai) {
            // Issue order 0, 3, 1, 2: an N = 128 MMA has ONE accumulate flag for both halves, so
            // each half is first written by its own N = 64 MMA (a = 0 -> even, a = 3 -> odd; these
            // carry the overwrite flag at r = 0, k = 0) and the N = 128 MMAs always accumulate.
            const int a = ai == 0 ? 0 : ai == 1 ? 3 : ai - 1;
            // a = 0: odd plane col j, a = 1: even plane col j, a = 2: odd plane col j+1, a = 3: even plane col j+1
            const uint32_t a_off = static_cast<uint32_t>(
                (((a & 1) ? kPlaneBytes : 0) + (r * 9 + (a >> 1)) * 128) >> 4);
            // B: a = 0 -> W(r,0) = second half of tile (r,0); a = 1 -> tile (r,0); a = 2 -> tile (r,1);
            //    a = 3 -> W(r,2) = first half of tile (r,1)
            const uint32_t b_off = static_cast<uint32_t>(
                ((2 * r + (a >> 1)) * 2 * kTapBytes + (a == 0 ? kTapBytes : 0)) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = a_st + a_off + static_cast<uint32_t>(2 * k);
              const uint64_t bd = b_d0 + b_off + static_cast<uint32_t>(2 * k);
              if (a == 0) {
                ptx::umma_f16(d_even, ad, bd, idesc64, (r | k) != 0);
              } else if (a == 3) {
                ptx::umma_f16(d_odd, ad, bd, idesc64, (r | k) != 0);
              } else {
                ptx::umma_f16(d_even, ad, bd, idesc128, 1);
              }
            }
          }
        }
        ptx::umma_commit(&empty_bar[stage]);
        ptx::umma_commit(&tfull_bar[acc]);
      }
      __syncwarp();
      CERB_PROF_ADD(prof_c, t_m2);
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if (p.prof != nullptr && lane == 0) {
      long long* o = p.prof + blockIdx.x * 16;
      o[1] = prof_a; o[2] = prof_b; o[3] = prof_c; o[8] = clock64() - t_all;
    }
  } else if (FUSE && warp >= 10) {
    // ------------------------------------------------------------------ halo fix-up (warps 10-15)
    // The planes hold the SKIP halo; add bilinear_x2(low) in place. A work item is a 2x2 block of
    // halo pixels (rows 2 bj, 2 bj + 1; columns 2 bi = odd plane column bi, 2 bi + 1 = even plane
    // column bi) x one 16-byte channel group: exactly the block below / right of low pixel (j, i)
    // of the stand-alone pass (ops_misc.cu), whose arithmetic is shared (upadd_math.cuh). Pixels
    // outside the image stay zero (TMA fill = the convolution's padding).
    const int tid_f = (warp - 10) * 32 + lane;
    const int PH = p.H >> 1, PW = p.W >> 1;
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0;; ++it) {
      ptx::mbar_wait(&full_bar[stage], phase, p.err_flag, 47);
      const int rg = s_ring[it & 7];
      if (rg < 0) {  // pass the end marker on to the MMA warp
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&fixed_bar[stage]);
        break;
      }
      const int img = rg / regions_per_img;
      const int rem = rg - img * regions_per_img;
      const int ry = rem / p.regions_x, rx = rem - ry * p.regions_x;
      const int j0 = ry * 8 - 1, i0 = rx * 8 - 1;  // low pixel of block (0, 0) = origin of the low box
      uint8_t* plane0 = sA + stage * kStageBytes;  // odd columns | even columns
      for (int idx = tid_f; idx < kFixItems; idx += kFixWarps * 32) {
        const int g = idx & 7;
        const int blk = idx >> 3;
        const int bj = blk / 9, bi = blk - bj * 9;
        const int j = j0 + bj, i = i0 + bi;
        const int Y = 2 * j + 1, X = 2 * i + 1;  // top-left output pixel of the block
        const bool oky[2] = {Y >= 0 && Y < p.H, Y + 1 < p.H};
        const bool okx[2] = {X >= 0 && X < p.W, X + 1 < p.W};
        if (!((oky[0] || oky[1]) && (okx[0] || okx[1]))) continue;
        const int by0 = max(j, 0) - j0, by1 = min(j + 1, PH - 1) - j0;
        const int bx0 = max(i, 0) - i0, bx1 = min(i + 1, PW - 1) - i0;
        // all eight loads first: the four stores below would otherwise order every later load
        // behind them (the compiler cannot tell the cells apart)
        const int q00 = by0 * kLowPx + bx0, q01 = by0 * kLowPx + bx1;
        const int q10 = by1 * kLowPx + bx0, q11 = by1 * kLowPx + bx1;
        const uint4 l00 = *reinterpret_cast<const uint4*>(sLow + q00 * 128 + ((g ^ (q00 & 7)) << 4));
        const uint4 l01 = *reinterpret_cast<const uint4*>(sLow + q01 * 128 + ((g ^ (q01 & 7)) << 4));
        const uint4 l10 = *reinterpret_cast<const uint4*>(sLow + q10 * 128 + ((g ^ (q10 & 7)) << 4));
        const uint4 l11 = *reinterpret_cast<const uint4*>(sLow + q11 * 128 + ((g ^ (q11 & 7)) << 4));
        uint4* cell[4];
        uint4 sk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int pix = (2 * bj + (k >> 1)) * 9 + bi;
          cell[k] = reinterpret_cast<uint4*>(plane0 + (k & 1) * kPlaneBytes + pix * 128 + ((g ^ (pix & 7)) << 4));
          sk[k] = *cell[k];
        }
        uint4 o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int dy = k >> 1, dx = k & 1;
          const float ly = (j < 0) ? 0.0f : (dy == 0 ? 0.25f : 0.75f);
          const float lx = (i < 0) ? 0.0f : (dx == 0 ? 0.25f : 0.75f);
          o[k] = upadd_pixel8_h2(l00, l01, l10, l11, sk[k], ly, lx);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (oky[k >> 1] && okx[k & 1]) *cell[k] = o[k];
      }
      ptx::fence_proxy_async_smem();  // the tensor core reads these planes through the async proxy
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&fixed_bar[stage]);
        ptx::mbar_arrive(low_empty);
      }
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2-9)
    // warps 2-5 drain the even-pixel half of the accumulator (TMEM columns 0..63), warps 6-9 the
    // odd-pixel half (64..127); both write rows of the SAME staging tile, one TMA store per region.
    ptx::grid_dep_wait();
    const int half = (warp - 2) >> 2;        // pixel parity handled by this warp
    const int q = warp & 3;
    const int m = q * 32 + lane;             // TMEM lane = (y, j): y = m >> 3, j = m & 7
    // Staging: one [16 y][8 j] plane per pixel parity (16 KB each), row = TMEM lane. Consecutive
    // lanes then write consecutive 128-byte rows, which the 128-byte swizzle spreads over all banks
    // (interleaving the parities in one tile would put a warp on rows of one parity only: 8-way
    // bank conflicts). The planes go out / come in through tensor maps with element stride 2 in x.
    uint8_t* my_row = sOut + half * (kOutBytes / 2) + m * 128;
    const int sw = m & 7;
    const bool store_warp = warp == 2;
    long long prof_a = 0, prof_b = 0, prof_c = 0, prof_d = 0;
    uint32_t res_phase = 0;
    int it = 0;
    for (;; ++it) {
      const int acc = it % kAccStages;
      const uint32_t acc_phase = (it / kAccStages) & 1;
      CERB_PROF_T0(t_e0);
      ptx::mbar_wait(&tfull_bar[acc], acc_phase, p.err_flag, 45);
      CERB_PROF_ADD(prof_a, t_e0);
      const int rg = s_ring[it & 7];
      if (rg < 0) break;
      const int img = rg / regions_per_img;
      const int rem = rg - img * regions_per_img;
      const int ry = rem / p.regions_x, rx = rem - ry * p.regions_x;
      const int x0 = rx * 16, y0 = ry * 16;
      if (store_warp && ptx::elect_one()) {
        ptx::bulk_wait_read<0>();  // the previous store has drained the staging tile
        if (p.has_res) {
          ptx::mbar_arrive_expect_tx(res_bar, kOutBytes);
          ptx::tma_load_4d(sOut, &p.res_map, res_bar, 0, x0, y0, img);
          ptx::tma_load_4d(sOut + kOutBytes / 2, &p.res_map, res_bar, 0, x0 + 1, y0, img);
        }
      }
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 128 + half * 64;
      uint32_t r0[32], r1[32];
      ptx::tmem_ld32(taddr, r0);
      ptx::tmem_ld32(taddr + 32, r1);
      ptx::tmem_ld_wait();
      // this warp's share of the accumulator is in registers: hand it back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      CERB_PROF_T0(t_e1);
      if (p.has_res) {
        ptx::mbar_wait(res_bar, res_phase, p.err_flag, 46);
        res_phase ^= 1;
      } else {
        ptx::named_bar_sync(1, 256);  // the elected lane has seen the staging tile drained
      }
      CERB_PROF_ADD(prof_b, t_e1);
      CERB_PROF_T0(t_e2);
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        float v[32];
        const float sc = p.acc_scale;
        if (p.bias != nullptr) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + c2 * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b = __ldg(b4 + i);
            v[4 * i + 0] = fmaf(__uint_as_float(c2 == 0 ? r0[4 * i + 0] : r1[4 * i + 0]), sc, b.x);
            v[4 * i + 1] = fmaf(__uint_as_float(c2 == 0 ? r0[4 * i + 1] : r1[4 * i + 1]), sc, b.y);
            v[4 * i + 2] = fmaf(__uint_as_float(c2 == 0 ? r0[4 * i + 2] : r1[4 * i + 2]), sc, b.z);
            v[4 * i + 3] = fmaf(__uint_as_float(c2 == 0 ? r0[4 * i + 3] : r1[4 * i + 3]), sc, b.w);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(c2 == 0 ? r0[i] : r1[i]) * sc;
        }
        if (p.has_res) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 u = *reinterpret_cast<const uint4*>(my_row + (((c2 * 4 + i) ^ sw) << 4));
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
          *reinterpret_cast<uint4*>(my_row + (((c2 * 4 + i) ^ sw) << 4)) = u;
        }
      }
      CERB_PROF_ADD(prof_c, t_e2);
      CERB_PROF_T0(t_e3);
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(1, 256);
      if (store_warp && ptx::elect_one()) {
        ptx::tma_store_4d(&p.out_map, sOut, 0, x0, y0, img);
        ptx::tma_store_4d(&p.out_map, sOut + kOutBytes / 2, 0, x0 + 1, y0, img);
        ptx::bulk_commit_group();
      }
      CERB_PROF_ADD(prof_d, t_e3);
    }
    if (store_warp) ptx::bulk_wait_all<0>();
    if (p.prof != nullptr && q == 2 && lane == 0) {
      long long* o = p.prof + blockIdx.x * 16 + (half == 0 ? 4 : 10);
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

void conv64x_plan(Conv64xParams& p) {
  p.regions_x = (p.W + 15) / 16;
  p.regions_y = (p.H + 15) / 16;
  p.n_regions = p.n_img * p.regions_x * p.regions_y;
  int n = (224 * 1024 - kWBytes - kOutBytes - 1024 - (p.fuse_up ? kLowBytes : 0)) / kStageBytes;
  if (n > 4) n = 4;
  if (n < 2) n = 2;
  p.n_stages = n;
}

size_t conv64x_smem_bytes(const Conv64xParams& p) {
  return static_cast<size_t>(kWBytes) + kOutBytes + static_cast<size_t>(p.n_stages) * kStageBytes +
         (p.fuse_up ? kLowBytes : 0) + 256 + 1024;
}

cudaError_t conv64x_launch(const Conv64xParams& p, int num_sms, cudaStream_t stream, bool pdl) {
  static bool attr_set[2] = {false, false};
  const int fuse = p.fuse_up ? 1 : 0;
  void (*kern)(Conv64xParams) = fuse ? conv64x_kernel<true> : conv64x_kernel<false>;
  if (!attr_set[fuse]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[fuse] = true;
  }
  const int grid = p.n_regions < num_sms ? p.n_regions : num_sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(fuse ? kConv64xFuseThreads : kConv64xThreads);
  cfg.dynamicSmemBytes = conv64x_smem_bytes(p);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

}  // namespace cerb
