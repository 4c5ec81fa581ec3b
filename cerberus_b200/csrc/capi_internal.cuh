// Shared between the translation units that implement include/cerberus_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/cerberus_b200.h"

struct cerb_ctx {
  int device = 0;
  int precision = 0;
  int num_sms = 148;
  int conv_sms = 148;  // CTAs of the persistent convolution kernels (option "conv_sms")
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // D2H copies overlapped with compute
  cudaStream_t up_stream = nullptr;    // H2D copies overlapped with compute
  cudaStream_t side_stream = nullptr;  // graph branch for ops flagged `side` (Patch-Class)
  cudaEvent_t side_event = nullptr;
  cudaEvent_t order_event = nullptr;
  cudaEvent_t slot_event[8] = {};  // download-stream marks (cerb_copy_mark / cerb_copy_wait)
  cudaEvent_t mark_event[8] = {};  // compute-stream marks (cerb_ctx_mark / cerb_ctx_wait_mark)
  int* err_flag_host = nullptr;  // mapped pinned memory: readable even after a trapped kernel
  int* err_flag_dev = nullptr;
  void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled, resolved at run time (no libcuda link)
  int64_t launches = 0;
  bool use_pdl = true;     // convolution kernels: programmatic dependent launch (prologue overlap)
  bool use_graphs = true;  // replay the forward op list as a CUDA graph
  unsigned long long* stat_dev = nullptr;  // [4] device counters of the small-tile watershed (cerb_ctx_stat)
  int64_t stat_ws_large = 0, stat_ws_fallback = 0;  // large-image watershed calls / exact fallbacks
  int64_t stat_ws_tied = 0;  // components of large images whose marker ties were settled by enumeration
  int conv64_debug = 0;
  long long* prof_dev = nullptr;  // [kProfSlots] in-kernel attribution counters (option "kernel_prof")
  int ws_mode = 0;  // 0: component-parallel watershed with exact fallback; 1: whole-tile emulation only;
                    // 2 (test hook): image 0 of every batch is redone by the exact emulation
  bool dyn_sched = true;  // dynamic tile scheduling in the persistent conv kernels
  int stem_mode = 1;    // 1: 7x7 stem on conv64.cu (mode 4); 0: generic kernel
  int k_rotate = 1;     // per-CTA rotated K walk in conv3x3.cu (de-synchronises weight-slab reads)
  bool conv64s = true;  // split-precision mode: 64->64 3x3 layers on csrc/conv64s.cu (0: generic kernel)
  int conv3_pair = 1;  // wide 3x3 stride-1 layers on CTA pairs (csrc/conv3x3c2.cu): 0 never, 1 Cout % 256 == 0, 2 also Cout 128 / 64
  bool fuse_upadd = true;   // fp16 mode: UPADD ops read only by a 64->64 3x3 convolution run inside its producer (conv64x.cu)
  bool conv3_chain = true;  // consecutive pair-kernel layers of one geometry run as ONE launch (layer chains)
  int conv3_mode = 1;   // 0: generic kernel for the wide 3x3 layers; 1: conv3x3.cu (cout <= 512); 2: always
  int conv64_mode = -1;  // -1: generic kernel for every conv; 0/1/2: conv64.cu halo layout
  std::vector<void*> scratch;  // device allocations owned by the ctx (post-proc workspaces)
  void* postproc_ws = nullptr;             // csrc/postproc.cu Workspace, created on first use
  void (*postproc_ws_free)(void*) = nullptr;
  // ring of small parameter blocks (patch top-left tables): pinned host slot -> device slot by an
  // asynchronous copy, so that device-resident tile plumbing calls return without synchronising
  static constexpr int kParamSlots = 64;
  static constexpr size_t kParamBytes = 8192;
  char* param_host = nullptr;
  char* param_dev = nullptr;
  cudaEvent_t param_event[kParamSlots] = {};
  int param_next = 0;
  void* instinfo_ws = nullptr;             // csrc/instinfo.cu Workspace, created on first use
  void (*instinfo_ws_free)(void*) = nullptr;
};

namespace cerb {

constexpr int kProfSlots = 256 * 16;

extern thread_local std::string g_last_error;
int fail(int code, const char* fmt, ...);
// cerb_plan_create with an optional device-resident weight blob owned by someone else (a
// cerb_model shares one upload between the plans of all its batch shapes)
int plan_create_impl(cerb_ctx* ctx, const cerb_tensor_desc* tensors, int n_tensors, const cerb_op* ops,
                     int n_ops, const void* weight_blob, size_t blob_bytes, uint8_t* shared_dev_blob,
                     cerb_plan** out);

}  // namespace cerb

#define CERB_CUDA(call)                                                                     \
  do {                                                                                      \
    cudaError_t cerb_e__ = (call);                                                          \
    if (cerb_e__ != cudaSuccess)                                                            \
      return cerb::fail(CERB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(cerb_e__)); \
  } while (0)
