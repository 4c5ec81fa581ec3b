// Model-level C ABI: builds the forward op graph of one batch shape inside the library and runs
// it (cerb_model_* / cerb_forward), plus the NCCL weight broadcast (cerb_bcast_weights).
//
// The graph is the reference's NetDesc.forward (models/net_desc.py:144-200) over the ResNet34
// encoder (models/backbone/resnet.py:202-211,273-286) restructured for the device exactly as
// cerberus_b200/plan.py::PlanSpec documents: BN folded, bias / residual / ReLU in the conv
// epilogue, the decoders' common first stage computed once with concatenated output channels,
// the 64->96->C head fused into one kernel that writes the per-patch canvas
// (models/run_desc.py:451-491, infer/tile.py:116-134). tests/test_model_spec.py checks on the CPU
// that cerb_model_spec returns the same tensors and ops as PlanSpec for every shape it tries.
#include <dlfcn.h>

#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "capi_internal.cuh"

using namespace cerb;

namespace {

// BasicBlock encoders (models/backbone/resnet.py:292-313): resnet34 has 3, 4, 6, 3 blocks per stage,
// resnet18 2, 2, 2, 2 - the builder takes the count from the layer table (consecutive block indices)
constexpr int kFilters[5] = {64, 64, 128, 256, 512};

struct Spec {
  std::vector<cerb_tensor_desc> tensors;
  std::vector<cerb_op> ops;
  int canvas = -1;
  int logits[CERB_MAX_DECODERS + 1];
};

struct Builder {
  const cerb_model_desc& d;
  const cerb_layer* layers;
  int n_layers;
  Spec& s;

  const cerb_layer* find(int role, int a = 0, int b = 0, int c = 0) const {
    for (int i = 0; i < n_layers; ++i) {
      const cerb_layer& l = layers[i];
      if (l.role == role && l.a == a && l.b == b && l.c == c) return &l;
    }
    return nullptr;
  }
  int T(int n, int h, int w, int c, int dt = CERB_F16) {
    s.tensors.push_back(cerb_tensor_desc{n, h, w, c, dt});
    return static_cast<int>(s.tensors.size()) - 1;
  }
  static cerb_op blank(int kind) {
    cerb_op o;
    std::memset(&o, 0, sizeof(o));
    o.kind = kind;
    o.in0 = o.in1 = o.out = -1;
    o.logits_out = -1;
    o.w_off = o.w_lo_off = o.b_off = -1;
    o.aux_w_off = o.aux_b_off = -1;
    o.tail_w_off = o.tail_b_off = -1;
    return o;
  }
  cerb_op conv(const cerb_layer& l, int src, int dst, int relu, int stride = 1, int residual = -1,
               int stem = 0, int in_coff = 0, int out_coff = 0) {
    cerb_op o = blank(CERB_OP_CONV);
    o.in0 = src;
    o.in1 = residual;
    o.out = dst;
    o.in_coff = in_coff;
    o.in_c = stem ? 8 : l.cin;
    o.out_coff = out_coff;
    o.cout = l.cout;
    o.kh = o.kw = l.kh;
    o.stride = stride;
    o.pad = l.kh / 2;
    o.relu = relu;
    o.stem = stem;
    o.w_off = l.w_off;
    o.w_lo_off = l.w_lo_off;
    o.b_off = l.b_off;
    o.w_shift = l.w_shift;
    return o;
  }
};

#define NEED(var, ...)                                                                     \
  const cerb_layer* var = b.find(__VA_ARGS__);                                             \
  if (!var) return fail(CERB_ERR_ARG, "cerb_model: layer table lacks %s", #__VA_ARGS__)

int build_spec(const cerb_model_desc& d, const cerb_layer* layers, int n_layers, int n, int h, int w,
               int out_h, int out_w, bool want_logits, Spec& s) {
  if (n <= 0 || h <= 0 || w <= 0 || (h % 16) || (w % 16))
    return fail(CERB_ERR_ARG, "input size must be a multiple of 16 (got %dx%d, n=%d)", h, w, n);
  if (out_h > h || out_w > w || out_h <= 0 || out_w <= 0)
    return fail(CERB_ERR_ARG, "output shape %dx%d exceeds the input shape %dx%d", out_h, out_w, h, w);
  if (d.n_decoders < 0 || d.n_decoders > CERB_MAX_DECODERS)
    return fail(CERB_ERR_ARG, "cerb_model_desc: n_decoders %d out of range", d.n_decoders);
  Builder b{d, layers, n_layers, s};
  for (int i = 0; i <= CERB_MAX_DECODERS; ++i) s.logits[i] = -1;

  const int t_in = b.T(n, h, w, 3, CERB_U8);
  const int t_prep = b.T(n, h, w + 8, 8);
  {
    cerb_op o = Builder::blank(CERB_OP_PREP);
    o.in0 = t_in;
    o.out = t_prep;
    s.ops.push_back(o);
  }
  NEED(stem, CERB_L_STEM);
  const int x0 = b.T(n, h, w, 64);
  s.ops.push_back(b.conv(*stem, t_prep, x0, 1, 1, -1, 1));
  const int hs[5] = {h, h / 2, h / 4, h / 8, h / 16};
  const int ws[5] = {w, w / 2, w / 4, w / 8, w / 16};
  const int pool = b.T(n, hs[1], ws[1], 64);
  {
    cerb_op o = Builder::blank(CERB_OP_MAXPOOL);
    o.in0 = x0;
    o.out = pool;
    s.ops.push_back(o);
  }
  int feats[5] = {x0, -1, -1, -1, -1};
  int cur = pool;
  for (int li = 1; li <= 4; ++li) {
    const int c = kFilters[li];
    const int mid = b.T(n, hs[li], ws[li], c);
    const int o_t = b.T(n, hs[li], ws[li], c);
    const int ds = li > 1 ? b.T(n, hs[li], ws[li], c) : -1;
    const int pair0 = li == 1 ? pool : ds, pair1 = o_t;
    int n_blocks = 0;
    while (b.find(CERB_L_BLOCK_CONV1, li, n_blocks) != nullptr) ++n_blocks;
    if (n_blocks == 0) return fail(CERB_ERR_ARG, "cerb_model: layer table lacks encoder stage %d", li);
    for (int bi = 0; bi < n_blocks; ++bi) {
      const int stride = (li > 1 && bi == 0) ? 2 : 1;
      NEED(c1, CERB_L_BLOCK_CONV1, li, bi);
      NEED(c2, CERB_L_BLOCK_CONV2, li, bi);
      s.ops.push_back(b.conv(*c1, cur, mid, 1, stride));
      int res, dst;
      const cerb_layer* down = b.find(CERB_L_BLOCK_DOWN, li, bi);
      if (down) {
        s.ops.push_back(b.conv(*down, cur, ds, 0, 2));
        res = ds;
        dst = o_t;
      } else {
        res = cur;  // never in place: the residual (= block input) is read by the epilogue
        dst = cur == pair1 ? pair0 : pair1;
      }
      s.ops.push_back(b.conv(*c2, mid, dst, 1, 1, res));
      cur = dst;
    }
    feats[li] = cur;
  }
  const int x1 = feats[1], x2 = feats[2], x3 = feats[3], x4 = feats[4];

  const int canvas = b.T(n, out_h, out_w, d.canvas_c, CERB_F32);
  s.canvas = canvas;
  if (d.has_pclass) {
    NEED(pc, CERB_L_PCLASS);
    int lg = -1;
    if (want_logits) lg = b.T(n, 1, 1, pc->classes, CERB_F32);
    cerb_op o = Builder::blank(CERB_OP_PCLASS);
    o.in0 = x4;
    o.out = canvas;
    o.out_coff = d.pclass_canvas_coff;
    o.cout = pc->classes;
    o.logits_out = lg;
    o.w_off = pc->w_off;
    o.side = 1;
    s.ops.push_back(o);
    s.logits[CERB_MAX_DECODERS] = lg;
  }
  const int D = d.n_decoders;
  if (D > 0) {
    NEED(cmap, CERB_L_CONV_MAP);
    NEED(first, CERB_L_DEC_FIRST);
    const int f4 = b.T(n, hs[4], ws[4], 256);
    s.ops.push_back(b.conv(*cmap, x4, f4, 0));
    auto upadd = [&](int skip, int prev, int out) {
      cerb_op o = Builder::blank(CERB_OP_UPADD);
      o.in0 = skip;
      o.in1 = prev;
      o.out = out;
      s.ops.push_back(o);
    };
    const int s3 = b.T(n, hs[3], ws[3], 256);
    upadd(x3, f4, s3);
    const int u4a = b.T(n, hs[3], ws[3], 256 * D);
    s.ops.push_back(b.conv(*first, s3, u4a, 1));
    const int u4b = b.T(n, hs[3], ws[3], 128);
    const int s2 = b.T(n, hs[2], ws[2], 128);
    const int a2 = b.T(n, hs[2], ws[2], 128);
    const int b2 = b.T(n, hs[2], ws[2], 64);
    const int s1 = b.T(n, hs[1], ws[1], 64);
    const int a1 = b.T(n, hs[1], ws[1], 64);
    const int b1 = b.T(n, hs[1], ws[1], 64);
    const int s0 = b.T(n, h, w, 64);
    const int a0 = b.T(n, h, w, 64);
    const int b0 = b.T(n, h, w, 64);
    for (int di = 0; di < D; ++di) {
      NEED(d01, CERB_L_DEC_CONV, di, 0, 1);
      NEED(d10, CERB_L_DEC_CONV, di, 1, 0);
      NEED(d11, CERB_L_DEC_CONV, di, 1, 1);
      NEED(d20, CERB_L_DEC_CONV, di, 2, 0);
      NEED(d21, CERB_L_DEC_CONV, di, 2, 1);
      NEED(d30, CERB_L_DEC_CONV, di, 3, 0);
      NEED(d31, CERB_L_DEC_CONV, di, 3, 1);
      NEED(hh, CERB_L_HEAD_HIDDEN, di);
      NEED(ho, CERB_L_HEAD_OUT, di);
      s.ops.push_back(b.conv(*d01, u4a, u4b, 1, 1, -1, 0, di * 256));
      upadd(x2, u4b, s2);
      s.ops.push_back(b.conv(*d10, s2, a2, 1));
      s.ops.push_back(b.conv(*d11, a2, b2, 1));
      upadd(x1, b2, s1);
      s.ops.push_back(b.conv(*d20, s1, a1, 1));
      s.ops.push_back(b.conv(*d21, a1, b1, 1));
      upadd(x0, b1, s0);
      s.ops.push_back(b.conv(*d30, s0, a0, 1));
      s.ops.push_back(b.conv(*d31, a0, b0, 1));
      int lg = -1;
      if (want_logits) {
        lg = b.T(n, h, w, ho->classes, CERB_F32);
        s.logits[di] = lg;
      }
      // 1x1 64->96 + BN + ReLU + 1x1 96->C + softmax / argmax / crop in ONE kernel
      cerb_op o = b.conv(*hh, b0, canvas, 1, 1, -1, 0, 0, d.canvas_coff[di]);
      o.aux_classes = ho->classes;
      o.aux_w_off = ho->w_off;
      o.aux_b_off = ho->b_off;
      o.head_mode = d.head_mode[di];
      o.logits_out = lg;
      s.ops.push_back(o);
    }
  }
  return CERB_OK;
}

// ---------------------------------------------------------------- NCCL, resolved at run time
struct Nccl {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, char[128], int) = nullptr;  // ncclUniqueId passed by value = 128 bytes
  int (*CommDestroy)(void*) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

struct UniqueId {
  char b[128];
};

int nccl_load(Nccl** out) {
  static Nccl nc;
  if (!nc.h) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nme : names) {
      nc.h = dlopen(nme, RTLD_NOW | RTLD_GLOBAL);
      if (nc.h) break;
    }
    if (!nc.h) return fail(CERB_ERR_ARG, "NCCL not found (dlopen libnccl.so.2): %s", dlerror());
    nc.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(nc.h, "ncclGetUniqueId"));
    nc.CommInitRank = reinterpret_cast<int (*)(void**, int, char[128], int)>(dlsym(nc.h, "ncclCommInitRank"));
    nc.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(nc.h, "ncclCommDestroy"));
    nc.Broadcast = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(
        dlsym(nc.h, "ncclBroadcast"));
    nc.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(nc.h, "ncclGetErrorString"));
    if (!nc.GetUniqueId || !nc.CommInitRank || !nc.CommDestroy || !nc.Broadcast)
      return fail(CERB_ERR_ARG, "libnccl lacks a required symbol");
  }
  *out = &nc;
  return CERB_OK;
}

int nccl_fail(Nccl* nc, const char* what, int rc) {
  return fail(CERB_ERR_CUDA, "%s failed: %s", what, nc->GetErrorString ? nc->GetErrorString(rc) : "?");
}

}  // namespace

struct cerb_model {
  cerb_ctx* ctx = nullptr;
  cerb_model_desc desc;
  std::vector<cerb_layer> layers;
  uint8_t* blob = nullptr;  // device
  size_t blob_bytes = 0;
  struct Entry {
    cerb_plan* plan;
    int canvas;
  };
  std::map<std::tuple<int, int, int, int, int, int>, Entry> plans;
};

extern "C" int cerb_model_spec(const cerb_model_desc* desc, const cerb_layer* layers, int n_layers, int n,
                               int h, int w, int out_h, int out_w, int want_logits,
                               cerb_tensor_desc* tensors, int* n_tensors, cerb_op* ops, int* n_ops,
                               int32_t* canvas_tensor, int32_t* logit_tensors) {
  if (!desc || !layers || n_layers <= 0 || !n_tensors || !n_ops)
    return fail(CERB_ERR_ARG, "cerb_model_spec: bad arguments");
  Spec s;
  const int rc = build_spec(*desc, layers, n_layers, n, h, w, out_h, out_w, want_logits != 0, s);
  if (rc) return rc;
  const int nt = static_cast<int>(s.tensors.size()), no = static_cast<int>(s.ops.size());
  if (tensors && ops) {
    if (*n_tensors < nt || *n_ops < no)
      return fail(CERB_ERR_ARG, "cerb_model_spec: need room for %d tensors and %d ops", nt, no);
    std::memcpy(tensors, s.tensors.data(), sizeof(cerb_tensor_desc) * nt);
    std::memcpy(ops, s.ops.data(), sizeof(cerb_op) * no);
  }
  *n_tensors = nt;
  *n_ops = no;
  if (canvas_tensor) *canvas_tensor = s.canvas;
  if (logit_tensors) std::memcpy(logit_tensors, s.logits, sizeof(s.logits));
  return CERB_OK;
}

extern "C" int cerb_model_create(cerb_ctx* ctx, const cerb_model_desc* desc, const cerb_layer* layers,
                                 int n_layers, const void* weight_blob, size_t blob_bytes,
                                 cerb_model** out) {
  if (!ctx || !desc || !layers || n_layers <= 0 || blob_bytes == 0 || !out)
    return fail(CERB_ERR_ARG, "cerb_model_create: bad arguments");
  *out = nullptr;
  for (int i = 0; i < n_layers; ++i) {
    const cerb_layer& l = layers[i];
    if (l.w_off < 0 || static_cast<size_t>(l.w_off) >= blob_bytes ||
        (l.b_off >= 0 && static_cast<size_t>(l.b_off) >= blob_bytes))
      return fail(CERB_ERR_ARG, "cerb_model_create: layer %d points outside the blob", i);
  }
  CERB_CUDA(cudaSetDevice(ctx->device));
  cerb_model* m = new cerb_model();
  m->ctx = ctx;
  m->desc = *desc;
  m->layers.assign(layers, layers + n_layers);
  m->blob_bytes = blob_bytes;
  if (cudaMalloc(reinterpret_cast<void**>(&m->blob), blob_bytes) != cudaSuccess) {
    delete m;
    return fail(CERB_ERR_CUDA, "cudaMalloc(%zu) for the weight blob failed", blob_bytes);
  }
  if (weight_blob &&
      cudaMemcpy(m->blob, weight_blob, blob_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(m->blob);
    delete m;
    return fail(CERB_ERR_CUDA, "weight blob H2D copy failed");
  }
  *out = m;
  return CERB_OK;
}

extern "C" void cerb_model_destroy(cerb_model* m) {
  if (!m) return;
  for (auto& kv : m->plans) cerb_plan_destroy(kv.second.plan);
  cudaSetDevice(m->ctx->device);
  if (m->blob) cudaFree(m->blob);
  delete m;
}

extern "C" int cerb_model_plan(cerb_model* m, int n, int h, int w, int out_h, int out_w, int want_logits,
                               cerb_plan** plan, int32_t* canvas_tensor) {
  if (!m || !plan) return fail(CERB_ERR_ARG, "cerb_model_plan: bad arguments");
  const auto key = std::make_tuple(n, h, w, out_h, out_w, want_logits != 0 ? 1 : 0);
  auto it = m->plans.find(key);
  if (it == m->plans.end()) {
    Spec s;
    int rc = build_spec(m->desc, m->layers.data(), static_cast<int>(m->layers.size()), n, h, w, out_h,
                        out_w, want_logits != 0, s);
    if (rc) return rc;
    cerb_plan* pl = nullptr;
    rc = plan_create_impl(m->ctx, s.tensors.data(), static_cast<int>(s.tensors.size()), s.ops.data(),
                          static_cast<int>(s.ops.size()), nullptr, m->blob_bytes, m->blob, &pl);
    if (rc) return rc;
    it = m->plans.emplace(key, cerb_model::Entry{pl, s.canvas}).first;
  }
  *plan = it->second.plan;
  if (canvas_tensor) *canvas_tensor = it->second.canvas;
  return CERB_OK;
}

extern "C" int cerb_forward(cerb_model* m, const uint8_t* tiles, int tiles_on_device, int n, int h, int w,
                            int out_h, int out_w, float* canvas_out, int canvas_on_device) {
  if (!m || !tiles) return fail(CERB_ERR_ARG, "cerb_forward: bad arguments");
  cerb_plan* pl = nullptr;
  int32_t canvas = -1;
  int rc = cerb_model_plan(m, n, h, w, out_h, out_w, 0, &pl, &canvas);
  if (rc) return rc;
  if ((rc = cerb_plan_run(pl, tiles, tiles_on_device))) return rc;
  if (!canvas_out) return CERB_OK;
  const size_t bytes = static_cast<size_t>(n) * out_h * out_w * m->desc.canvas_c * sizeof(float);
  if (canvas_on_device)
    return cerb_memcpy(m->ctx, canvas_out, cerb_plan_tensor_ptr(pl, canvas, 0), bytes, 3);
  return cerb_plan_read_tensor(pl, canvas, 0, canvas_out, bytes);
}

extern "C" int cerb_nccl_unique_id(void* id128) {
  if (!id128) return fail(CERB_ERR_ARG, "cerb_nccl_unique_id: null");
  Nccl* nc = nullptr;
  int rc = nccl_load(&nc);
  if (rc) return rc;
  const int r = nc->GetUniqueId(id128);
  return r ? nccl_fail(nc, "ncclGetUniqueId", r) : CERB_OK;
}

extern "C" int cerb_nccl_comm_create(cerb_ctx* ctx, int nranks, int rank, const void* id128, void** comm) {
  if (!ctx || !id128 || !comm || nranks <= 0 || rank < 0 || rank >= nranks)
    return fail(CERB_ERR_ARG, "cerb_nccl_comm_create: bad arguments");
  Nccl* nc = nullptr;
  int rc = nccl_load(&nc);
  if (rc) return rc;
  CERB_CUDA(cudaSetDevice(ctx->device));
  UniqueId id;
  std::memcpy(id.b, id128, 128);
  // ncclCommInitRank(ncclComm_t*, int, ncclUniqueId /* by value, 128 bytes */, int)
  typedef int (*InitFn)(void**, int, UniqueId, int);
  const int r = reinterpret_cast<InitFn>(nc->CommInitRank)(comm, nranks, id, rank);
  return r ? nccl_fail(nc, "ncclCommInitRank", r) : CERB_OK;
}

extern "C" int cerb_nccl_comm_destroy(void* comm) {
  if (!comm) return CERB_OK;
  Nccl* nc = nullptr;
  int rc = nccl_load(&nc);
  if (rc) return rc;
  const int r = nc->CommDestroy(comm);
  return r ? nccl_fail(nc, "ncclCommDestroy", r) : CERB_OK;
}

extern "C" int cerb_bcast_weights(cerb_ctx* ctx, cerb_model** model, void* comm, int root, int rank) {
  if (!ctx || !model || !comm) return fail(CERB_ERR_ARG, "cerb_bcast_weights: bad arguments");
  const bool is_root = rank == root;
  if (is_root && !*model) return fail(CERB_ERR_ARG, "cerb_bcast_weights: the root rank needs a model");
  Nccl* nc = nullptr;
  int rc = nccl_load(&nc);
  if (rc) return rc;
  CERB_CUDA(cudaSetDevice(ctx->device));
  // header: n_layers, blob_bytes, then the description and the layer table, all as bytes
  struct Header {
    int64_t n_layers, blob_bytes;
    cerb_model_desc desc;
  } hd;
  std::memset(&hd, 0, sizeof(hd));
  if (is_root) {
    hd.n_layers = static_cast<int64_t>((*model)->layers.size());
    hd.blob_bytes = static_cast<int64_t>((*model)->blob_bytes);
    hd.desc = (*model)->desc;
  }
  auto bcast_host = [&](void* host, size_t bytes) -> int {
    void* dev = nullptr;
    CERB_CUDA(cudaMalloc(&dev, bytes));
    if (is_root) CERB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    const int r = nc->Broadcast(dev, dev, bytes, /*ncclChar*/ 0, root, comm, ctx->stream);
    if (r) { cudaFree(dev); return nccl_fail(nc, "ncclBroadcast", r); }
    if (!is_root) CERB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CERB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(dev);
    return CERB_OK;
  };
  if ((rc = bcast_host(&hd, sizeof(hd)))) return rc;
  if (hd.n_layers <= 0 || hd.blob_bytes <= 0) return fail(CERB_ERR_ARG, "cerb_bcast_weights: empty model at the root");
  std::vector<cerb_layer> layers(static_cast<size_t>(hd.n_layers));
  if (is_root) layers = (*model)->layers;
  if ((rc = bcast_host(layers.data(), sizeof(cerb_layer) * layers.size()))) return rc;
  if (!is_root) {
    rc = cerb_model_create(ctx, &hd.desc, layers.data(), static_cast<int>(layers.size()), nullptr,
                           static_cast<size_t>(hd.blob_bytes), model);
    if (rc) return rc;
  }
  const int r = nc->Broadcast((*model)->blob, (*model)->blob, (*model)->blob_bytes, 0, root, comm, ctx->stream);
  if (r) return nccl_fail(nc, "ncclBroadcast(weights)", r);
  CERB_CUDA(cudaStreamSynchronize(ctx->stream));
  return CERB_OK;
}
