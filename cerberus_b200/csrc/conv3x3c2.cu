// 3x3 stride-1 pad-1 convolution, Cin = 128 / 256 / 512, Cout = 64, 128 or a multiple of 256, on CTA
// PAIRS: encoder layer2-4 (models/backbone/resnet.py:203-211) and the decoder stages u4 / u3
// (models/utils/net_layers.py:23-28) - the layers conv3x3.cu runs at half
// of the tensor peak because an N = 128 single-CTA MMA needs the whole 128 B/clk of shared-memory
// bandwidth for its operands while TMA is filling the same memory.
//
// One work item = a 16x16-pixel region x BN output channels (BN = 256, or the layer's 128 / 64),
// executed by the two CTAs of a cluster (the two SMs of a TPC) as ONE tcgen05.mma.cta_group::2 of
// M = 256, N = BN per k-step (described below for BN = 256):
//   * CTA r owns the left / right 16x8 half of the region: its 18x10 input halo (one 64-channel
//     chunk, 23 KB) is its 128 rows of A, the nine taps being nine descriptors into the halo;
//   * CTA r holds rows [128 r, 128 r + 128) of the [256 x 64] weight slab of a (tap, chunk): half
//     of B. The hardware feeds both halves to both tensor cores, so a CTA streams 16 KB of
//     weights per 512 cycles of MMA instead of 16 KB per 256, and reads 8 KB of operands per
//     128-cycle MMA (64 B/clk) instead of 8 KB per 64 cycles;
//   * the accumulator is 128 lanes x 256 columns of TMEM in EACH CTA (two stages = all 512
//     columns); each CTA's epilogue warps drain their own half region.
// Only the leader CTA (cluster rank 0) issues MMAs. Operand barriers live in the leader: its
// producer posts the expected byte count of BOTH CTAs, the peer's TMA loads complete on the
// leader's barrier (cp.async.bulk.tensor.cta_group::2); tcgen05.commit multicasts the "stage
// free" / "accumulator full" arrivals to both CTAs; the peer's epilogue warps release an
// accumulator with a remote mbarrier arrive. Work items come from a global counter (leader) and
// reach the peer through distributed shared memory + a remote arrive.
#include "conv3x3c2.cuh"
#include "ptx.cuh"

namespace cerb {

namespace {

#define CERB_PROF_T0(var) const long long var = p.prof != nullptr ? clock64() : 0
#define CERB_PROF_ADD(acc, var) \
  do { if (p.prof != nullptr) acc += clock64() - var; } while (0)

constexpr int kRegion = 16;
constexpr int kHaloH = 18, kHaloW = 10;            // halo of a 16 (rows) x 8 (columns) half region
constexpr int kATxBytes = kHaloH * kHaloW * 128;   // 23040
constexpr int kAStageBytes = 23 * 1024;            // padded to the swizzle period
constexpr int kAStages = 2;                        // halo pipeline depth
constexpr int kSlabBytes = 128 * 128;              // staging: 128 pixels x 64 fp16 channels
constexpr int kMaxBStages = 8;
constexpr int kTmemCols = 512;                     // 2 accumulator stages x 256 columns

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(ptx::smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int* err_flag, int code) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const uint64_t t0 = ptx::global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 0x3FF) == 0 && ptx::global_timer_ns() - t0 > 2000000000ull) {
      if (err_flag) atomicExch(err_flag, code);
      __threadfence_system();
      asm volatile("trap;");
    }
  }
}

struct Item {
  int layer, nt, rx, ry, img;
};

__device__ __forceinline__ Item decode(const Conv3c2Params& p, int item) {
  Item it;
  it.layer = item / p.n_items_layer;  // chains number their items layer-major
  item -= it.layer * p.n_items_layer;
  it.nt = item % p.n_ntiles;
  const int rg = item / p.n_ntiles;
  it.rx = rg % p.regions_x;
  const int t = rg / p.regions_x;
  it.ry = t % p.regions_y;
  it.img = t / p.regions_y;
  return it;
}

__device__ __forceinline__ const Conv3c2Layer& layer_of(const Conv3c2Params& p, int layer) {
  return p.layers != nullptr ? p.layers[layer] : p.l0;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConv3c2Threads, 1)
conv3x3c2_kernel(const __grid_constant__ Conv3c2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_bstages = p.n_bstages;
  const int BN = p.BN;                       // output channels per work item: 64, 128 or 256
  const int b_stage_bytes = (BN >> 1) * 128;  // this CTA's half of a weight slab
  const uint32_t rank = ptx::cluster_ctarank();  // 0 = leader (left half region), 1 = peer (right half)
  const bool is_leader = rank == 0;
  const int pair = static_cast<int>(blockIdx.x >> 1), n_pairs = static_cast<int>(gridDim.x >> 1);

  // identical layout in both CTAs: the MMA descriptors and the multicast commits address the same
  // offsets in the leader's and the peer's shared memory
  uint8_t* sA = smem;                                // halo stages
  uint8_t* sOut = sA + kAStages * kAStageBytes;             // 2 output / residual slabs
  uint8_t* sB = sOut + 2 * kSlabBytes;               // weight half-slab pipeline
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + n_bstages * b_stage_bytes);
  uint64_t* a_empty = a_full + kAStages;
  uint64_t* b_full = a_empty + kAStages;
  uint64_t* b_empty = b_full + kMaxBStages;
  uint64_t* tfull_bar = b_empty + kMaxBStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;
  uint64_t* item_bar = res_bar + 2;  // peer only: "work item i has been published" (slot i & 7)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(item_bar + 8);
  volatile int* s_ring = reinterpret_cast<volatile int*>(tmem_holder + 2);

  if (warp == 0 && lane == 0) {
    if (p.layers == nullptr) {
      ptx::prefetch_tmap(&p.l0.in_map);
      ptx::prefetch_tmap(&p.l0.w_map);
      ptx::prefetch_tmap(&p.l0.out_map);
      if (p.l0.has_res) ptx::prefetch_tmap(&p.l0.res_map);
    }
    for (int s = 0; s < kAStages; ++s) {
      ptx::mbar_init(&a_full[s], 1);   // leader's producer (expect_tx of both CTAs)
      ptx::mbar_init(&a_empty[s], 1);  // multicast commit
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull_bar[s], 1);
      ptx::mbar_init(&tempty_bar[s], 8);  // four epilogue warps in each CTA of the pair
      ptx::mbar_init(&res_bar[s], 1);
    }
    for (int s = 0; s < 8; ++s) ptx::mbar_init(&item_bar[s], 1);
    for (int s = 0; s < kMaxBStages; ++s) {
      ptx::mbar_init(&b_full[s], 1);
      ptx::mbar_init(&b_empty[s], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(tmem_holder, kTmemCols);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();  // the peer's barriers exist before anything arrives on them remotely
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  ptx::grid_dep_launch();
  ptx::grid_dep_wait();
  const int n_chunks = p.n_chunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    const bool leader_lane_p = ptx::elect_one() != 0;
    const int elected = __ffs(__ballot_sync(0xffffffffu, leader_lane_p)) - 1;
    // operand barriers of the LEADER, as shared::cluster addresses usable from both CTAs
    const uint32_t a_full_l0 = ptx::mapa_u32(ptx::smem_u32(&a_full[0]), 0);  // barriers are 8 bytes apart
    const uint32_t b_full_l0 = ptx::mapa_u32(ptx::smem_u32(&b_full[0]), 0);
    long long prof_a = 0;
    int a_issued = 0, b_cnt = 0, n_fetched = 0;
    auto fetch = [&]() -> int {
      int item = 0;
      const int slot = n_fetched & 7;
      if (is_leader) {
        if (p.tile_counter != nullptr) {
          if (leader_lane_p) item = atomicAdd(p.tile_counter, 1);
          item = __shfl_sync(0xffffffffu, item, elected);
        } else {
          item = pair + n_fetched * n_pairs;
        }
        if (item >= p.n_items) item = -1;
        if (leader_lane_p) {
          s_ring[slot] = item;
          ptx::st_cluster_u32(ptx::mapa_u32(ptx::smem_u32(const_cast<int*>(&s_ring[slot])), 1),
                              static_cast<uint32_t>(item));
          ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&item_bar[slot]), 1));  // release.cluster
        }
        __syncwarp();
      } else {
        mbar_wait_cluster(&item_bar[slot], (n_fetched >> 3) & 1, p.err_flag, 61);
        item = s_ring[slot];
      }
      ++n_fetched;
      return item;
    };
    // The (item, chunk) pairs form one flat stream; the halo of the NEXT pair is requested while
    // tap 4 of the current pair is being queued, i.e. as soon as its stage can have been released.
    // (A decoupled halo stream polling for free stages on every tap, with two or three stages,
    // measured slower: 0.048-0.050 ms against 0.044 ms on 256->256 at 32x32, batch 32.)
    // Chains: the first halo of an item of layer > 0 may only be requested once the previous layer
    // of its image is stored. `block` = false while loads of the CURRENT item are still to be
    // issued: a producer that blocked there could hold back the very items others wait for
    // (its own included); the caller then retries after the tap loop.
    auto issue_a = [&](int item, int c, bool block) -> bool {
      const Item it = decode(p, item);
      if (c == 0 && it.layer > 0) {
        int ok = 1;
        if (leader_lane_p)
          ok = chain::wait_count(p.done + (it.layer - 1) * p.n_img + it.img, 2 * p.items_per_img, block,
                                 p.err_flag, 69) ? 1 : 0;
        ok = __shfl_sync(0xffffffffu, ok, elected);
        if (!ok) return false;
      }
      const int st = a_issued % kAStages;
      const uint32_t ph = (a_issued / kAStages) & 1;
      ptx::mbar_wait(&a_empty[st], ph ^ 1, p.err_flag, 62);
      if (leader_lane_p) {
        if (is_leader) ptx::mbar_arrive_expect_tx(&a_full[st], 2 * kATxBytes);
        ptx::tma_load_4d_2sm(sA + st * kAStageBytes, &layer_of(p, it.layer).in_map, a_full_l0 + 8u * st, c * 64,
                             it.rx * kRegion + 8 * static_cast<int>(rank) - 1, it.ry * kRegion - 1, it.img);
      }
      __syncwarp();
      ++a_issued;
      return true;
    };
    int cur_item = fetch(), cur_c = 0;
    if (cur_item >= 0) issue_a(cur_item, 0, true);
    int nxt_item = cur_item, nxt_c = 1;
    if (cur_item >= 0 && nxt_c == n_chunks) { nxt_c = 0; nxt_item = fetch(); }
    while (cur_item >= 0) {
      const Item cur = decode(p, cur_item);
      const int nt = cur.nt;
      const CUtensorMap* w_map = &layer_of(p, cur.layer).w_map;
      bool a_deferred = false;
      for (int t = 0; t < 9; ++t) {
        if (t == 4 && nxt_item >= 0) a_deferred = !issue_a(nxt_item, nxt_c, false);
        const int bs = b_cnt % n_bstages;
        const uint32_t bph = (b_cnt / n_bstages) & 1;
        CERB_PROF_T0(t_p);
        ptx::mbar_wait(&b_empty[bs], bph ^ 1, p.err_flag, 63);
        CERB_PROF_ADD(prof_a, t_p);
        if (leader_lane_p) {
          if (is_leader) ptx::mbar_arrive_expect_tx(&b_full[bs], 2 * b_stage_bytes);
          ptx::tma_load_2d_2sm(sB + bs * b_stage_bytes, w_map, b_full_l0 + 8u * bs, (t * n_chunks + cur_c) * 64,
                               nt * BN + (BN >> 1) * static_cast<int>(rank));
        }
        __syncwarp();
        ++b_cnt;
      }
      if (a_deferred) issue_a(nxt_item, nxt_c, true);
      cur_item = nxt_item;
      cur_c = nxt_c;
      if (nxt_item >= 0 && ++nxt_c == n_chunks) { nxt_c = 0; nxt_item = fetch(); }
    }
    if (is_leader) {  // end marker: wake the MMA warp on the halo barrier it will wait on next
      const int st = a_issued % kAStages;
      ptx::mbar_wait(&a_empty[st], ((a_issued / kAStages) & 1) ^ 1, p.err_flag, 62);
      if (leader_lane_p) ptx::mbar_arrive(&a_full[st]);
      __syncwarp();
    }
    if (p.prof != nullptr && lane == 0) p.prof[blockIdx.x * 16 + 0] = prof_a;
  } else if (warp == 1) {
    if (is_leader) {
      // ---------------------------------------------------------------- MMA issuer (leader CTA)
      const bool leader_lane_m = ptx::elect_one() != 0;
      const uint32_t idesc = ptx::umma_idesc_f16(256, BN);
      const uint32_t b_stage_u = static_cast<uint32_t>(b_stage_bytes >> 4);
      // A: rows = the 16x8 pixels of a half region inside its 10-pixel-pitch halo; an 8-row group
      // is one image row (8 x 128 B), groups are one halo row (10 x 128 B) apart.
      const uint64_t a_d0 = ptx::umma_desc_sw128(ptx::smem_u32(sA), kHaloW * 128);
      const uint64_t b_d0 = ptx::umma_desc_sw128(ptx::smem_u32(sB), 1024);
      uint32_t tap_u[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_u[t] = static_cast<uint32_t>((((t / 3) * kHaloW + (t % 3)) * 128) >> 4);
      long long prof_a = 0, prof_b = 0, prof_c = 0, prof_d = 0;
      CERB_PROF_T0(t_all);
      int a_idx = 0, b_cnt = 0, it = 0;
      bool done = false;
      for (;; ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        CERB_PROF_T0(t_m0);
        mbar_wait_cluster(&tempty_bar[acc], acc_phase ^ 1, p.err_flag, 64);
        CERB_PROF_ADD(prof_a, t_m0);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 256;
        for (int c = 0; c < n_chunks; ++c, ++a_idx) {
          const int ast = a_idx % kAStages;
          CERB_PROF_T0(t_m1);
          ptx::mbar_wait(&a_full[ast], (a_idx / kAStages) & 1, p.err_flag, 65);
          CERB_PROF_ADD(prof_b, t_m1);
          if (c == 0 && s_ring[it & 7] < 0) {  // end marker: pass it on to the epilogue of both CTAs
            if (leader_lane_m) {
              ptx::mbar_arrive(&tfull_bar[acc]);
              ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&tfull_bar[acc]), 1));
            }
            __syncwarp();
            done = true;
            break;
          }
          const uint64_t a_st = a_d0 + static_cast<uint32_t>((ast * kAStageBytes) >> 4);
#pragma unroll
          for (int t = 0; t < 9; ++t, ++b_cnt) {
            const int bs = b_cnt % n_bstages;
            CERB_PROF_T0(t_m2);
            ptx::mbar_wait(&b_full[bs], (b_cnt / n_bstages) & 1, p.err_flag, 66);
            CERB_PROF_ADD(prof_c, t_m2);
            ptx::tc_fence_after();
            CERB_PROF_T0(t_m3);
            if (leader_lane_m) {
              const uint64_t bd = b_d0 + static_cast<uint32_t>(bs) * b_stage_u;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                ptx::umma_f16_2sm(tmem_d, a_st + tap_u[t] + static_cast<uint32_t>(2 * k),
                                  bd + static_cast<uint32_t>(2 * k), idesc, (c | t | k) != 0);
              }
              ptx::umma_commit_2sm(&b_empty[bs], 3);
              if (t == 8) ptx::umma_commit_2sm(&a_empty[ast], 3);
            }
            __syncwarp();
            CERB_PROF_ADD(prof_d, t_m3);
          }
        }
        if (done) break;
        if (leader_lane_m) ptx::umma_commit_2sm(&tfull_bar[acc], 3);
        __syncwarp();
      }
      if (p.prof != nullptr && lane == 0) {
        long long* o = p.prof + blockIdx.x * 16;
        o[1] = prof_a; o[2] = prof_b + prof_c; o[3] = prof_d; o[8] = clock64() - t_all; o[9] = prof_b;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs)
    const int q = warp & 3;
    const int m = q * 32 + lane;  // TMEM lane = pixel (y, x) of the half region: y = m >> 3, x = m & 7
    const int sw = m & 7;
    const bool store_warp = q == 2;
    const uint32_t tempty_l0 = ptx::mapa_u32(ptx::smem_u32(&tempty_bar[0]), 0);
    const uint32_t tempty_l1 = ptx::mapa_u32(ptx::smem_u32(&tempty_bar[1]), 0);
    long long prof_a = 0, prof_b = 0, prof_c = 0, prof_d = 0;
    int sidx = 0, it = 0;
    uint32_t res_phase = 0;  // bit b: parity of the next residual arrival in slab buffer b
    for (;; ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      CERB_PROF_T0(t_e0);
      mbar_wait_cluster(&tfull_bar[acc], acc_phase, p.err_flag, 67);
      CERB_PROF_ADD(prof_a, t_e0);
      const int item = s_ring[it & 7];
      if (item < 0) break;
      ptx::tc_fence_after();
      const Item im = decode(p, item);
      const Conv3c2Layer& L = layer_of(p, im.layer);
      const int has_res = L.has_res, relu = L.relu;
      const float acc_scale = L.acc_scale;
      const float* bias = L.bias;
      const int x0 = im.rx * kRegion + 8 * static_cast<int>(rank), y0 = im.ry * kRegion;
      const int n0 = im.nt * BN;
      const int n_slabs = BN >> 6;
      for (int slab = 0; slab < n_slabs; ++slab, ++sidx) {
        const int buf = sidx & 1;
        uint8_t* sO = sOut + buf * kSlabBytes;
        uint8_t* my_row = sO + m * 128;
        uint64_t* rbar = &res_bar[buf];
        if (store_warp && ptx::elect_one()) {
          ptx::bulk_wait_read<1>();  // the store that last used this slab buffer has drained it
          if (has_res) {
            ptx::mbar_arrive_expect_tx(rbar, kSlabBytes);
            ptx::tma_load_4d(sO, &L.res_map, rbar, n0 + slab * 64, x0, y0, im.img);
          }
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + slab * 64;
        uint32_t r0[32], r1[32];
        ptx::tmem_ld32(taddr, r0);
        ptx::tmem_ld32(taddr + 32, r1);
        ptx::tmem_ld_wait();
        if (slab == n_slabs - 1) {  // accumulator fully in registers: hand it back to the leader's MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(acc == 0 ? tempty_l0 : tempty_l1);
        }
        CERB_PROF_T0(t_e1);
        if (has_res) {
          ptx::mbar_wait(rbar, (res_phase >> buf) & 1, p.err_flag, 68);
          res_phase ^= 1u << buf;
        } else {
          ptx::named_bar_sync(1, 128);  // the elected lane has seen the slab buffer drained
        }
        CERB_PROF_ADD(prof_b, t_e1);
        CERB_PROF_T0(t_e2);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i)
            v[i] = __uint_as_float(half == 0 ? r0[i] : r1[i]) * acc_scale;
          if (bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + slab * 64 + half * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = __ldg(b4 + i);
              v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
            }
          }
          if (has_res) {
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
          if (relu) {
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
        CERB_PROF_ADD(prof_c, t_e2);
        CERB_PROF_T0(t_e3);
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(1, 128);
        if (store_warp && ptx::elect_one()) {
          ptx::tma_store_4d(&L.out_map, sO, n0 + slab * 64, x0, y0, im.img);
          ptx::bulk_commit_group();
          if (slab == n_slabs - 1 && im.layer + 1 < p.n_layers) {
            // chain: this CTA's half of the item is in global memory -> release it to the next layer
            chain::signal_stored(p.done + im.layer * p.n_img + im.img);
          }
        }
        CERB_PROF_ADD(prof_d, t_e3);
      }
    }
    if (store_warp) ptx::bulk_wait_all<0>();
    if (p.prof != nullptr && q == 2 && lane == 0) {
      long long* o = p.prof + blockIdx.x * 16 + 4;
      o[0] = prof_a; o[1] = prof_b; o[2] = prof_c; o[3] = prof_d;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();  // both CTAs are done with the pair's TMEM and with remote arrivals
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

}  // namespace

void conv3x3c2_plan(Conv3c2Params& p) {
  const int budget = 224 * 1024 - kAStages * kAStageBytes - 2 * kSlabBytes - 1024;
  int n = budget / ((p.BN >> 1) * 128);
  if (n > kMaxBStages) n = kMaxBStages;
  if (n < 2) n = 2;
  p.n_bstages = n;
  p.regions_x = (p.W + kRegion - 1) / kRegion;
  p.regions_y = (p.H + kRegion - 1) / kRegion;
  p.items_per_img = p.regions_x * p.regions_y * p.n_ntiles;
  p.n_items_layer = p.n_img * p.items_per_img;
  if (p.n_layers < 1) p.n_layers = 1;
  p.n_items = p.n_layers * p.n_items_layer;
}

size_t conv3x3c2_smem_bytes(const Conv3c2Params& p) {
  return static_cast<size_t>(kAStages * kAStageBytes + 2 * kSlabBytes) +
         static_cast<size_t>(p.n_bstages) * (p.BN >> 1) * 128 + 512 + 1024;
}

cudaError_t conv3x3c2_launch(const Conv3c2Params& p, int num_sms, cudaStream_t stream, bool pdl) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3c2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int pairs = num_sms / 2;
  if (p.n_items < pairs) pairs = p.n_items;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kConv3c2Threads);
  cfg.dynamicSmemBytes = conv3x3c2_smem_bytes(p);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, conv3x3c2_kernel, p);
}

}  // namespace cerb
