// 3x3 stride-1 pad-1 convolution for the wide layers (Cin = 128 / 256 / 512): encoder layer2-4
// (models/backbone/resnet.py:203-211) and the decoder stages u4..u2 (models/utils/net_layers.py:
// 23-28). fp16 operands, fp32 accumulation in TMEM, bias / residual / ReLU fused.
//
// Why a second kernel next to conv_tc.cu: measured with in-kernel role counters
// (tools/one_conv.py), the generic kernel spends half of its time waiting for operands. It
// streams a 16 KB activation slab AND a 16 KB weight slab from L2 for every 256 cycles of MMA
// (128 B/clk per SM); the L2 -> SM fabric delivers about 43 B/clk per SM when all 148 SMs pull.
// This kernel cuts the traffic to ~40 B/clk:
//   * one work item = a 16x16-pixel region x BN output channels. Its 18x18 input halo is loaded
//     ONCE per 64-channel chunk (41 KB) and the nine taps are nine shared-memory descriptors into
//     it (as in conv64.cu) instead of nine separately loaded 128-pixel slabs;
//   * the region is two 16x8 M tiles with their own TMEM accumulators, so every weight slab
//     fetched from L2 feeds eight MMAs (512 cycles) instead of four.
// Epilogue as in conv64.cu: rows are staged in swizzled shared memory and written by TMA stores
// (which also clip partial regions); the residual tile is TMA-loaded into the same staging slab.
#include "conv3x3.cuh"
#include "ptx.cuh"

namespace cerb {

namespace {

#define CERB_PROF_T0(var) const long long var = p.prof != nullptr ? clock64() : 0
#define CERB_PROF_ADD(acc, var) \
  do { if (p.prof != nullptr) acc += clock64() - var; } while (0)

constexpr int kRegion = 16;                       // region edge in pixels
constexpr int kHalo = kRegion + 2;                // 18
constexpr int kAStageBytes = 42 * 1024;           // 18*18*128 = 41472, padded to the swizzle period
constexpr int kATxBytes = kHalo * kHalo * 128;
constexpr int kSlabBytes = 128 * 128;             // 128 pixels x 64 fp16 channels
constexpr int kMaxBStages = 8;
constexpr int kTmemCols = 512;                    // 2 accumulator stages x 2 M tiles x 128 columns

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

struct Item {
  int layer, nt, rx, ry, img;
};

__device__ __forceinline__ Item decode(const Conv3Params& p, int item) {
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

__device__ __forceinline__ const ConvChainLayer& layer_of(const Conv3Params& p, int layer) {
  return p.layers != nullptr ? p.layers[layer] : p.l0;
}

__global__ void __launch_bounds__(kConv3Threads, 1)
conv3x3_kernel(const __grid_constant__ Conv3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_bstages = p.n_bstages;
  const int b_stage_bytes = p.BN * 128;

  uint8_t* sA = smem;                                // 2 halo stages
  uint8_t* sOut = sA + 2 * kAStageBytes;             // [group][buffer] output / residual slabs
  uint8_t* sB = sOut + 4 * kSlabBytes;               // weight-slab pipeline
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + n_bstages * b_stage_bytes);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + kMaxBStages;
  uint64_t* tfull_bar = b_empty + kMaxBStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;  // [group * 2 + buffer]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(res_bar + 4);
  // work-item ids handed from the producer to the MMA / epilogue warps (dynamic scheduling,
  // see conv64x.cu); -1 ends the kernel
  volatile int* s_ring = reinterpret_cast<volatile int*>(tmem_holder + 2);

  if (warp == 0 && lane == 0) {
    if (p.layers == nullptr) {
      ptx::prefetch_tmap(&p.l0.in_map);
      ptx::prefetch_tmap(&p.l0.w_map);
      ptx::prefetch_tmap(&p.l0.out_map);
      if (p.l0.has_res) ptx::prefetch_tmap(&p.l0.res_map);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&a_full[s], 1);
      ptx::mbar_init(&a_empty[s], 1);
      ptx::mbar_init(&tfull_bar[s], 1);
      ptx::mbar_init(&tempty_bar[s], 8);  // one arrival per epilogue warp
    }
    for (int s = 0; s < kMaxBStages; ++s) {
      ptx::mbar_init(&b_full[s], 1);
      ptx::mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 4; ++s) ptx::mbar_init(&res_bar[s], 1);
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
  ptx::grid_dep_wait();  // every global access below depends on the previous kernel (weights
                         // are streamed together with the activations here)
  const int n_chunks = p.n_chunks;
  // Every CTA needs the SAME weight slabs; walking K in the same order makes all 148 SMs hit the
  // same L2 lines in the same window (measured: the MMA warp waits ~17 % of the time for weight
  // slabs although the aggregate L2 traffic is low). Each CTA therefore starts its K walk at a
  // different (tap, chunk); fp32 accumulation order differs per CTA but is fixed per launch.
  const int rot_t = p.rotate ? static_cast<int>(blockIdx.x % 9) : 0;
  const int rot_c = p.rotate ? static_cast<int>((blockIdx.x / 9) % n_chunks) : 0;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // One elected lane issues (see conv64.cu for why not `lane == 0`). The (item, chunk) pairs
    // form one flat stream; the halo of the NEXT pair is requested while tap 4 of the current
    // pair is being queued, i.e. as soon as its stage can have been released.
    const bool leader = ptx::elect_one() != 0;
    long long prof_a = 0;
    int a_issued = 0, b_cnt = 0;
    const int leader_lane = __ffs(__ballot_sync(0xffffffffu, leader)) - 1;
    // Chains (conv_chain.cuh): the first halo of an item of layer > 0 waits for the previous layer
    // of its image. `block` = false while loads of the CURRENT item are still to be issued (a
    // producer blocked there could hold back the very items others wait for); the caller retries
    // after the tap loop.
    auto issue_a = [&](int item, int c_seq, bool block) -> bool {
      const int c = (c_seq + rot_c) % n_chunks;
      const Item it = decode(p, item);
      if (c_seq == 0 && it.layer > 0) {
        int ok = 1;
        if (leader)
          ok = chain::wait_count(p.done + (it.layer - 1) * p.n_img + it.img, 2 * p.items_per_img, block,
                                 p.err_flag, 39) ? 1 : 0;
        ok = __shfl_sync(0xffffffffu, ok, leader_lane);
        if (!ok) return false;
      }
      const int st = a_issued & 1;
      const uint32_t ph = (a_issued >> 1) & 1;
      ptx::mbar_wait(&a_empty[st], ph ^ 1, p.err_flag, 31);
      if (leader) {
        ptx::mbar_arrive_expect_tx(&a_full[st], kATxBytes);
        ptx::tma_load_4d(sA + st * kAStageBytes, &layer_of(p, it.layer).in_map, &a_full[st], c * 64,
                         it.rx * kRegion - 1, it.ry * kRegion - 1, it.img);
      }
      __syncwarp();
      ++a_issued;
      return true;
    };
    int n_fetched = 0;
    auto fetch = [&]() -> int {
      int item = 0;
      if (p.tile_counter != nullptr) {
        if (leader) item = atomicAdd(p.tile_counter, 1);
        item = __shfl_sync(0xffffffffu, item, leader_lane);
      } else {
        item = static_cast<int>(blockIdx.x) + n_fetched * static_cast<int>(gridDim.x);
      }
      if (item >= p.n_items) item = -1;
      if (leader) s_ring[n_fetched & 7] = item;  // before the arrive that announces its first halo
      ++n_fetched;
      return item;
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
        ptx::mbar_wait(&b_empty[bs], bph ^ 1, p.err_flag, 32);
        CERB_PROF_ADD(prof_a, t_p);
        if (leader) {
          ptx::mbar_arrive_expect_tx(&b_full[bs], b_stage_bytes);
          const int tap = (t + rot_t) % 9, chunk = (cur_c + rot_c) % n_chunks;
          ptx::tma_load_2d(sB + bs * b_stage_bytes, w_map, &b_full[bs], (tap * n_chunks + chunk) * 64,
                           nt * p.BN);
        }
        __syncwarp();
        ++b_cnt;
      }
      if (a_deferred) issue_a(nxt_item, nxt_c, true);
      cur_item = nxt_item;
      cur_c = nxt_c;
      if (nxt_item >= 0 && ++nxt_c == n_chunks) { nxt_c = 0; nxt_item = fetch(); }
    }
    {  // end marker: wake the MMA warp on the halo barrier it will wait on next
      const int st = a_issued & 1;
      ptx::mbar_wait(&a_empty[st], ((a_issued >> 1) & 1) ^ 1, p.err_flag, 31);
      if (leader) ptx::mbar_arrive(&a_full[st]);
      __syncwarp();
    }
    if (p.prof != nullptr && lane == 0) p.prof[blockIdx.x * 16 + 0] = prof_a;
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = ptx::elect_one() != 0;
    const uint32_t idesc = ptx::umma_idesc_f16(128, p.BN);
    // A: rows = the 16x8 pixels of one half region inside the 18-pixel-pitch halo; an 8-row
    // group is one image row (8 x 128 B), groups are one halo row (18 x 128 B) apart.
    const uint64_t a_d0 = ptx::umma_desc_sw128(ptx::smem_u32(sA), kHalo * 128);
    const uint64_t b_d0 = ptx::umma_desc_sw128(ptx::smem_u32(sB), 1024);
    const uint32_t b_stage_u = static_cast<uint32_t>(b_stage_bytes >> 4);
    uint32_t tap_u[9];  // halo offset (16-byte units) of the tap used at K-walk position t
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int tap = (t + rot_t) % 9;
      tap_u[t] = static_cast<uint32_t>((((tap / 3) * kHalo + (tap % 3)) * 128) >> 4);
    }
    long long prof_a = 0, prof_b = 0, prof_c = 0, prof_d = 0;
    CERB_PROF_T0(t_all);
    int a_idx = 0, b_cnt = 0, it = 0;
    bool done = false;
    for (;; ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      CERB_PROF_T0(t_m0);
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1, p.err_flag, 33);
      CERB_PROF_ADD(prof_a, t_m0);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * 256;
      for (int c = 0; c < n_chunks; ++c, ++a_idx) {
        const int ast = a_idx & 1;
        CERB_PROF_T0(t_m1);
        ptx::mbar_wait(&a_full[ast], (a_idx >> 1) & 1, p.err_flag, 34);
        CERB_PROF_ADD(prof_b, t_m1);
        if (c == 0 && s_ring[it & 7] < 0) {  // end marker: pass it on to the epilogue
          if (leader) ptx::mbar_arrive(&tfull_bar[acc]);
          __syncwarp();
          done = true;
          break;
        }
        const uint64_t a_st = a_d0 + static_cast<uint32_t>((ast * kAStageBytes) >> 4);
#pragma unroll
        for (int t = 0; t < 9; ++t, ++b_cnt) {
          const int bs = b_cnt % n_bstages;
          CERB_PROF_T0(t_m2);
          ptx::mbar_wait(&b_full[bs], (b_cnt / n_bstages) & 1, p.err_flag, 35);
          CERB_PROF_ADD(prof_c, t_m2);
          ptx::tc_fence_after();
          CERB_PROF_T0(t_m3);
          if (leader) {
            const uint64_t bd = b_d0 + static_cast<uint32_t>(bs) * b_stage_u;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                ptx::umma_f16(tmem_d + j * 128, a_st + tap_u[t] + static_cast<uint32_t>(j * 64 + 2 * k),
                              bd + static_cast<uint32_t>(2 * k), idesc, (c | t | k) != 0);
              }
            }
            ptx::umma_commit(&b_empty[bs]);
            if (t == 8) ptx::umma_commit(&a_empty[ast]);
          }
          __syncwarp();
          CERB_PROF_ADD(prof_d, t_m3);
        }
      }
      if (done) break;
      if (leader) ptx::umma_commit(&tfull_bar[acc]);
      __syncwarp();
    }
    if (p.prof != nullptr && lane == 0) {
      long long* o = p.prof + blockIdx.x * 16;
      o[1] = prof_a; o[2] = prof_b + prof_c; o[3] = prof_d; o[8] = clock64() - t_all; o[9] = prof_b;
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    // group g (warps 2-5 / 6-9) owns the left / right 16x8 half of every region of this CTA
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int gtid = static_cast<int>(threadIdx.x) - (2 + 4 * g) * 32;
    const int sw = m & 7;
    // first warp of the group issues the TMA traffic through an elected lane (uniform branch:
    // no ELECT/BRA.U.ANY serialisation loop around UTMALDG / UTMASTG)
    const bool store_warp = q == 2;
    const int n_slabs = p.BN >> 6;
    long long prof_a = 0, prof_b = 0, prof_c = 0, prof_d = 0;
    int sidx = 0, it = 0;
    uint32_t res_phase = 0;  // bit b: parity of the next residual arrival in slab buffer b of this group
    for (;; ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      CERB_PROF_T0(t_e0);
      ptx::mbar_wait(&tfull_bar[acc], acc_phase, p.err_flag, 36);
      CERB_PROF_ADD(prof_a, t_e0);
      const int item = s_ring[it & 7];
      if (item < 0) break;
      ptx::tc_fence_after();
      const Item im = decode(p, item);
      const ConvChainLayer& L = layer_of(p, im.layer);
      const int has_res = L.has_res, relu = L.relu;
      const float acc_scale = L.acc_scale;
      const float* bias = L.bias;
      const int x0 = im.rx * kRegion + 8 * g, y0 = im.ry * kRegion;
      const int n0 = im.nt * p.BN;
      for (int slab = 0; slab < n_slabs; ++slab, ++sidx) {
        const int buf = sidx & 1;
        uint8_t* sO = sOut + (g * 2 + buf) * kSlabBytes;
        uint8_t* my_row = sO + m * 128;
        uint64_t* rbar = &res_bar[g * 2 + buf];
        if (store_warp && ptx::elect_one()) {
          ptx::bulk_wait_read<1>();  // the store that last used this slab buffer has drained it
          if (has_res) {
            ptx::mbar_arrive_expect_tx(rbar, kSlabBytes);
            ptx::tma_load_4d(sO, &L.res_map, rbar, n0 + slab * 64, x0, y0, im.img);
          }
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + g * 128 + slab * 64;
        uint32_t r0[32], r1[32];
        ptx::tmem_ld32(taddr, r0);
        ptx::tmem_ld32(taddr + 32, r1);
        ptx::tmem_ld_wait();
        if (slab == n_slabs - 1) {  // accumulator fully in registers: hand it back to the MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
        }
        CERB_PROF_T0(t_e1);
        if (has_res) {
          ptx::mbar_wait(rbar, (res_phase >> buf) & 1, p.err_flag, 37);
          res_phase ^= 1u << buf;
        } else {
          ptx::named_bar_sync(1 + g, 128);  // thread 0 has seen the slab buffer drained
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
        ptx::named_bar_sync(1 + g, 128);
        if (store_warp && ptx::elect_one()) {
          ptx::tma_store_4d(&L.out_map, sO, n0 + slab * 64, x0, y0, im.img);
          ptx::bulk_commit_group();
          // chain: this group's half of the item is in global memory -> release it to the next layer
          if (slab == n_slabs - 1 && im.layer + 1 < p.n_layers)
            chain::signal_stored(p.done + im.layer * p.n_img + im.img);
        }
        CERB_PROF_ADD(prof_d, t_e3);
      }
    }
    if (store_warp) ptx::bulk_wait_all<0>();  // only the elected lane has groups pending
    if (p.prof != nullptr && gtid == 0) {
      long long* o = p.prof + blockIdx.x * 16 + (g == 0 ? 4 : 10);
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

void conv3x3_plan(Conv3Params& p) {
  const int budget = 224 * 1024 - 2 * kAStageBytes - 4 * kSlabBytes - 1024;
  int n = budget / (p.BN * 128);
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

size_t conv3x3_smem_bytes(const Conv3Params& p) {
  return static_cast<size_t>(2 * kAStageBytes + 4 * kSlabBytes) +
         static_cast<size_t>(p.n_bstages) * p.BN * 128 + 512 + 1024;
}

cudaError_t conv3x3_launch(const Conv3Params& p, int num_sms, cudaStream_t stream, bool pdl) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = p.n_items < num_sms ? p.n_items : num_sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kConv3Threads);
  cfg.dynamicSmemBytes = conv3x3_smem_bytes(p);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, conv3x3_kernel, p);
}

}  // namespace cerb
