/* cerberus_b200 — C ABI of the B200-native Cerberus inference hot path.
 *
 * The reference (TissueImageAnalytics/cerberus) is pure Python on top of PyTorch and
 * scikit-image; it has no FFI of its own. This header is the drop-in boundary a
 * maintainer binds with ctypes to replace, for the tiled
 * inference path only:
 *
 *   B1  infer/base.py:51-53  run_step(input_batch, output_shape)
 *         = models/run_desc.py:439-502 infer_step  ->  models/net_desc.py:144-200 forward
 *       -> cerb_plan_create / cerb_plan_run / cerb_plan_read_*
 *   B2  loader/postproc.py:383-407  PostProcInstErodedContourMap.post_process
 *         (__proc_nuclei :352-381, __proc_gland :270-309, __proc_lumen :312-350)
 *       -> cerb_postproc_nuclei / cerb_postproc_gland_lumen
 *   a16 infer/tile.py:136-163  canvas stitch           -> cerb_stitch
 *   a19 infer/tile.py:187-191  lumen *= (gland > 0)    -> cerb_mask_lumen
 *   a20 loader/postproc.py:12-98 get_inst_info_dict    -> cerb_inst_info / cerb_inst_info_read
 *   a1/a2 infer/tile.py:43-106 + loader/infer_loader.py:57-69 (reflect pad + patch
 *       slicing)                                        -> cerb_extract_patches
 *
 * Conventions: every function returns 0 on success or a negative cerb_status; nothing
 * throws across the ABI; cerb_last_error() returns a thread-local, NUL-terminated
 * description of the last failure. The caller owns all host buffers. A ctx owns its
 * device buffers and one CUDA stream; a ctx is NOT thread-safe, different ctxs are
 * independent. There is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef CERBERUS_B200_H
#define CERBERUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cerb_ctx cerb_ctx;
typedef struct cerb_plan cerb_plan;
typedef struct cerb_model cerb_model;

enum cerb_status {
  CERB_OK = 0,
  CERB_ERR_CUDA = -1,     /* a CUDA runtime/driver call failed                        */
  CERB_ERR_ARG = -2,      /* invalid argument / unsupported shape                     */
  CERB_ERR_KERNEL = -3,   /* a kernel reported an internal error (pipeline watchdog)  */
  CERB_ERR_NO_DEVICE = -4 /* no sm_100 device; this library has no CPU path           */
};

enum cerb_dtype { CERB_U8 = 0, CERB_F16 = 1, CERB_F32 = 2, CERB_I32 = 3 };

/* Precision of the tensor-core path.
 * CERB_PREC_F16   : fp16 operands, fp32 accumulation (throughput mode).
 * CERB_PREC_F16X2 : every activation and weight carried as hi+lo fp16 pair, three MMAs per
 *                   product (hi*hi + lo*hi + hi*lo), fp32 accumulation: ~fp32 accuracy, used
 *                   for the 1e-3 logit gate against the fp32 reference. */
enum cerb_precision { CERB_PREC_F16 = 0, CERB_PREC_F16X2 = 1 };

/* NHWC tensor owned by a plan. `c` is the allocated channel count (pixel stride). */
typedef struct cerb_tensor_desc {
  int32_t n, h, w, c;
  int32_t dtype; /* cerb_dtype; F16 tensors get a second (lo) plane in F16X2 mode */
} cerb_tensor_desc;

enum cerb_op_kind {
  CERB_OP_PREP = 1,    /* u8 NHWC batch -> fp16 [N,H,W+8,8] zero-padded stem input (run_desc.py:440-441; /255 folded into stem weights, net_desc.py:147) */
  CERB_OP_CONV = 2,    /* conv + folded BN + bias (+residual) (+ReLU)                   */
  CERB_OP_MAXPOOL = 3, /* 3x3 s2 p1 (resnet.py:201)                                     */
  CERB_OP_UPADD = 4,   /* out = skip + bilinear2x(prev), align_corners=False (net_layers.py:45-46, net_desc.py:185-188) */
  CERB_OP_HEAD = 5,    /* 1x1 96->C (+bias), softmax / argmax / centre crop into the patch canvas (net_layers.py:36-38, run_desc.py:451-491) */
  CERB_OP_PCLASS = 6   /* Patch-Class branch (net_desc.py:64-76,169-180) + argmax broadcast (run_desc.py:459-461,479-486) */
};

enum cerb_head_mode {
  CERB_HEAD_INST = 0, /* softmax, drop channel 0 -> C-1 float channels */
  CERB_HEAD_TYPE = 1  /* argmax(softmax) -> 1 channel holding the class index as float */
};

/* One step of a plan. Unused fields are 0 / -1. Weight offsets are byte offsets into the
 * packed blob handed to cerb_plan_create (layout documented in cerberus_b200/pack.py). */
typedef struct cerb_op {
  int32_t kind;
  int32_t in0;      /* input tensor id                                                  */
  int32_t in1;      /* CONV: residual tensor id or -1; UPADD: low-res `prev` tensor id  */
  int32_t out;      /* output tensor id                                                 */
  int32_t in_coff;  /* CONV: first input channel read from in0; UPADD: first channel of in1 */
  int32_t in_c;     /* CONV: number of input channels (multiple of 64; stem: 8)         */
  int32_t out_coff; /* CONV/HEAD/PCLASS: first output channel written                   */
  int32_t cout;     /* CONV: output channels; HEAD: classes C; PCLASS: classes; UPADD: G > 1 = grouped form:
                       the G decoders' prev tensors in1 .. in1+G-1 are added to the ONE skip tensor in0
                       (read once) into out .. out+G-1                                  */
  int32_t kh, kw, stride, pad;
  int32_t relu;
  int32_t stem;       /* CONV: 1 = 7x7 stem reading the PREP tensor                     */
  int32_t head_mode;  /* HEAD: cerb_head_mode                                           */
  int32_t logits_out; /* HEAD/PCLASS: tensor id receiving raw fp32 logits, or -1        */
  int64_t w_off;      /* CONV: fp16 hi weights [cout][K]; HEAD/PCLASS: fp32 params      */
  int64_t w_lo_off;   /* CONV: fp16 lo weights (F16X2 mode), else -1                    */
  int64_t b_off;      /* fp32 bias [cout], or -1                                        */
  int32_t box_w;      /* CONV: tile box width override (power of two <= 128), 0 = auto  */
  int32_t w_shift;    /* CONV: weights are stored multiplied by 2^w_shift (keeps the fp16 lo plane out
                         of the subnormal range); the epilogue multiplies the accumulator by 2^-w_shift */
  int32_t aux_classes; /* CONV: > 0 fuses the classification-head tail into the epilogue: this conv is
                          the 1x1 64->96 hidden layer (+BN+ReLU), `out` is the fp32 CANVAS tensor, and
                          the 1x1 96->aux_classes (+bias) + softmax / argmax / centre crop of CERB_OP_HEAD
                          run per pixel in registers; head_mode / logits_out / out_coff as for HEAD */
  int32_t up_prev1;    /* CONV (64->64 3x3 s1, CERB_PREC_F16 only): tensor id + 1 of a half-resolution
                          tensor `prev`; the conv then reads  in0 + bilinear_x2(prev)  (the UPADD op fused
                          into the conv's producer, the sum is never written to HBM). 0 = none */
  int64_t aux_w_off;   /* fp32 [aux_classes][96] */
  int64_t aux_b_off;   /* fp32 [aux_classes] */
  /* CONV 64->64 3x3 s1 with aux_classes > 0 (CERB_PREC_F16 only): the WHOLE output head is fused
   * behind the conv - hidden 1x1 64->96 (+BN+ReLU; fp16 weights [96][64] at tail_w_off, fp32 bias at
   * tail_b_off, weight pre-scale tail_w_shift) and the aux_* tail; `out` is the fp32 CANVAS tensor and
   * neither the conv's 64-channel output nor the hidden tensor is written to HBM. -1 = not used. */
  int64_t tail_w_off;
  int64_t tail_b_off;
  int32_t tail_w_shift;
  int32_t side;        /* 1: nothing later in the plan reads this op's output; when the plan is replayed as a
                          CUDA graph the op runs on a parallel branch joined at the end (Patch-Class) */
} cerb_op;

/* ---- context ------------------------------------------------------------------- */
int cerb_ctx_create(int device, int precision, cerb_ctx** out);
void cerb_ctx_destroy(cerb_ctx* ctx);
const char* cerb_last_error(void);
int cerb_ctx_sync(cerb_ctx* ctx);
/* Tuning knobs applied to plans created afterwards. "conv64_mode": -1 = generic kernel for every
 * convolution, 0/1/2 = halo layout of the 64->64 3x3 kernel (csrc/conv64.cu). "conv3_mode": 0 =
 * generic kernel for the wide 3x3 stride-1 layers, 1 (default) = halo kernel csrc/conv3x3.cu for
 * Cout <= 512, 2 = for every Cout. "conv3_pair": 0 = never, 1 (default) = layers with Cout % 256
 * == 0 on the CTA-pair kernel (csrc/conv3x3c2.cu), 2 = also Cout 128 / 64. "conv3_chain": 1
 * (default) = consecutive wide 3x3 layers of one geometry run as ONE launch (csrc/conv_chain.cuh;
 * needs "dyn_sched" = 1, the default), 0 = one launch per layer. "cerb_plan_profile" spreads a
 * chain's time evenly over its layers. "use_graphs", "ws_mode", "kernel_prof": see capi.cu. */
int cerb_ctx_set_option(cerb_ctx* ctx, const char* name, int value);
/* Attribution evidence (option "kernel_prof" = 1 before creating the plan): the 64->64 3x3 kernel
 * stores, per CTA, 16 counters of clock cycles each role spent waiting (layout in csrc/conv64.cu;
 * the last launch wins). Copies the first n counters (n <= 4096) and optionally zeroes them. */
int cerb_ctx_read_prof(cerb_ctx* ctx, int64_t* out, int n, int reset);
/* Number of kernels this library launched on ctx since creation (bench's gpu_launches). */
int64_t cerb_ctx_launch_count(cerb_ctx* ctx);
/* Raw CUDA stream handle (cudaStream_t) of the ctx, for event timing by the caller. */
void* cerb_ctx_stream(cerb_ctx* ctx);

/* ---- plan: one compiled forward for a fixed [N,H,W] batch shape ------------------ */
int cerb_plan_create(cerb_ctx* ctx, const cerb_tensor_desc* tensors, int n_tensors,
                     const cerb_op* ops, int n_ops, const void* weight_blob, size_t blob_bytes,
                     cerb_plan** out);
void cerb_plan_destroy(cerb_plan* plan);
/* Runs every op. `input_u8` is the [N,H,W,3] uint8 batch for the PREP op's in0 tensor:
 * host memory when input_on_device == 0 (copied H2D on the ctx stream inside the call),
 * device memory otherwise. Asynchronous with respect to the host; pair with cerb_ctx_sync
 * or a cerb_plan_read_* call (which synchronises). */
int cerb_plan_run(cerb_plan* plan, const uint8_t* input_u8, int input_on_device);
int cerb_plan_num_ops(cerb_plan* plan);
/* Evidence for bench.py's roofline: runs the plan `reps` times on its current input with a
 * CUDA event recorded on the ctx stream after every op and returns the mean device time per
 * op in milliseconds (`ms_per_op[cerb_plan_num_ops]`, `kinds` optional). Synchronous. */
int cerb_plan_profile(cerb_plan* plan, int reps, float* ms_per_op, int32_t* kinds);
/* Device pointer of a plan tensor (plane 0 = hi / only plane, 1 = lo). */
void* cerb_plan_tensor_ptr(cerb_plan* plan, int tensor_id, int plane);
/* Synchronous D2H copy of a whole tensor plane into `host_dst` (`bytes` must match). */
int cerb_plan_read_tensor(cerb_plan* plan, int tensor_id, int plane, void* host_dst, size_t bytes);
/* Synchronous H2D copy into a tensor plane (tests feed intermediate activations). */
int cerb_plan_write_tensor(cerb_plan* plan, int tensor_id, int plane, const void* host_src,
                           size_t bytes);

/* ---- model: the whole forward driven from this header alone (SURVEY.md 8b "C ABI") ----------
 * The plan-level API above takes an op graph from its caller. cerb_model_* builds that graph
 * INSIDE the library from the architecture of models/net_desc.py:23-200 (ResNet34 encoder,
 * models/backbone/resnet.py:202-211; U-Net decoders, models/utils/net_layers.py:23-46; heads;
 * Patch-Class branch), so a host in any language runs a model directory with
 *     cerb_ctx_create -> cerb_model_create -> cerb_forward -> (cerb_postproc_* ...)
 * Only the folding of BatchNorm into the weights and their packing into one blob stay outside
 * (cerberus_b200/pack.py; layout = cerb_layer table below). */

#define CERB_MAX_DECODERS 8

enum cerb_layer_role {
  CERB_L_STEM = 0,        /* backbone.conv1 + bn1 (7x7, /255 folded in)                       */
  CERB_L_BLOCK_CONV1 = 1, /* backbone.layer<a>.<b>.conv1 + bn1   a = 1..4, b = block index    */
  CERB_L_BLOCK_CONV2 = 2, /* backbone.layer<a>.<b>.conv2 + bn2                                */
  CERB_L_BLOCK_DOWN = 3,  /* backbone.layer<a>.0.downsample (1x1 stride 2 + bn), a = 2..4     */
  CERB_L_CONV_MAP = 4,    /* conv_map 1x1 512->256 (net_desc.py:52)                           */
  CERB_L_DEC_FIRST = 5,   /* first conv of EVERY decoder's u4 block, output channels concatenated */
  CERB_L_DEC_CONV = 6,    /* decoder_head.<a>.<b>.block.<c>: a = decoder index, b = level 0..3, c = 0..1 */
  CERB_L_HEAD_HIDDEN = 7, /* output_head.<a>...x.0 (1x1 64->96 + BN + ReLU)                   */
  CERB_L_HEAD_OUT = 8,    /* output_head.<a>...x.1 (1x1 96->classes): fp32 [classes][96] at w_off, fp32 bias at b_off */
  CERB_L_PCLASS = 9       /* Patch-Class branch parameters (fp32, layout of cerberus_b200/plan.py) at w_off */
};

/* One packed layer of the weight blob. Offsets are byte offsets into the blob. */
typedef struct cerb_layer {
  int32_t role;        /* cerb_layer_role */
  int32_t a, b, c;     /* role-specific indices (see above), 0 when unused */
  int32_t cout, cin, kh, kw;
  int32_t w_shift;     /* conv weights are stored multiplied by 2^w_shift */
  int32_t classes;     /* HEAD_OUT / PCLASS */
  int64_t w_off;       /* fp16 hi weights [cout][kh*kw][cin] (HEAD_OUT / PCLASS: fp32 parameters) */
  int64_t w_lo_off;    /* fp16 lo weights, or -1 */
  int64_t b_off;       /* fp32 bias [cout], or -1 */
} cerb_layer;

/* What the checkpoint's settings.yml says about the heads (model_kwargs.decoder_kwargs filtered
 * by considered_tasks, in nn.ModuleDict order = decoder_kwargs order) and where each head's
 * channels sit in the per-patch canvas (the table of infer/tile.py:116-134). */
typedef struct cerb_model_desc {
  int32_t n_decoders;                       /* segmentation decoders (not Patch-Class) */
  int32_t head_mode[CERB_MAX_DECODERS];     /* cerb_head_mode */
  int32_t classes[CERB_MAX_DECODERS];
  int32_t canvas_coff[CERB_MAX_DECODERS];   /* first canvas channel of the head */
  int32_t has_pclass, pclass_classes, pclass_canvas_coff;
  int32_t canvas_c;                         /* channels of the canvas */
} cerb_model_desc;

/* The op graph (tensors + ops) the model runs for one batch shape, without touching a device:
 * introspection, and the CPU-side check that it equals the graph cerberus_b200/plan.py builds.
 * On entry *n_tensors / *n_ops are the capacities of the arrays, on exit the counts;
 * logit_tensors (nullable, int32 [CERB_MAX_DECODERS + 1]) receives the fp32 logit tensor of every
 * head (Patch-Class last) or -1. h and w must be multiples of 16. */
int cerb_model_spec(const cerb_model_desc* desc, const cerb_layer* layers, int n_layers, int n, int h,
                    int w, int out_h, int out_w, int want_logits, cerb_tensor_desc* tensors,
                    int* n_tensors, cerb_op* ops, int* n_ops, int32_t* canvas_tensor,
                    int32_t* logit_tensors);
/* Which UPADD ops of an op list cerb_plan_create folds into the 64->64 3x3 convolution that follows
 * them (option "fuse_upadd", fp16 mode, default 64->64 kernel): folded[i] = 1 for every such op.
 * Pure host code (no device): introspection and the CPU-side test of the folding rule - an UPADD
 * is folded iff the next op is that convolution reading its output and nothing else reads that
 * output before the tensor is written again. precision: cerb_precision. */
int cerb_plan_preview_folding(int precision, const cerb_tensor_desc* tensors, int n_tensors,
                              const cerb_op* ops, int n_ops, int32_t* folded);
/* Uploads the blob once (shared by every plan of the model). weight_blob may be NULL when the
 * weights will arrive by cerb_bcast_weights. */
int cerb_model_create(cerb_ctx* ctx, const cerb_model_desc* desc, const cerb_layer* layers,
                      int n_layers, const void* weight_blob, size_t blob_bytes, cerb_model** out);
void cerb_model_destroy(cerb_model* model);
/* infer/base.py:51-53 run_step for one batch: tiles u8 [n,h,w,3] (host, or device memory when
 * tiles_on_device != 0) -> the per-patch canvas f32 [n,out_h,out_w,canvas_c] (softmax / channel
 * drop / argmax / centre crop of models/run_desc.py:451-491 applied, channel layout of
 * infer/tile.py:116-134). Plans are built on first use of a shape and cached. canvas_out: host
 * buffer (the call synchronises), device buffer (canvas_on_device != 0: queued on the ctx stream),
 * or NULL (leave it in the plan: cerb_model_plan + cerb_plan_tensor_ptr). */
int cerb_forward(cerb_model* model, const uint8_t* tiles, int tiles_on_device, int n, int h, int w,
                 int out_h, int out_w, float* canvas_out, int canvas_on_device);
/* The cached plan of a shape (created if needed) and its canvas tensor id. */
int cerb_model_plan(cerb_model* model, int n, int h, int w, int out_h, int out_w, int want_logits,
                    cerb_plan** plan, int32_t* canvas_tensor);

/* ---- multi-GPU: one process per GPU; the only data-path collective of the inference path is
 * the start-up broadcast of the packed weights (replaces nn.DataParallel's
 * per-forward replicate of infer/base.py:46). NCCL is resolved at run time (dlopen libnccl.so.2);
 * communicators are opaque here (ncclComm_t). id128 = the 128 bytes of an ncclUniqueId: rank 0
 * creates it, the host program hands it to the other ranks (file, socket, MPI ...). */
int cerb_nccl_unique_id(void* id128);
int cerb_nccl_comm_create(cerb_ctx* ctx, int nranks, int rank, const void* id128, void** comm);
int cerb_nccl_comm_destroy(void* comm);
/* Root: *model is a created model, its description, layer table and device blob are broadcast.
 * Other ranks: *model == NULL on entry, a ready model on this rank's ctx on exit. */
int cerb_bcast_weights(cerb_ctx* ctx, cerb_model** model, void* nccl_comm, int root, int rank);

/* ---- tile plumbing (a1/a2/a16) --------------------------------------------------------
 * `flags` bit 0: the input buffer is device memory; bit 1: the output buffer is device
 * memory. A call with a host buffer is synchronous. cerb_extract_patches with both buffers
 * on the device and cerb_scatter_patches only QUEUE their work on the ctx stream (the host
 * tl_yx table is copied before they return): order later use with cerb_ctx_sync or by staying
 * on the ctx stream. */

/* infer/tile.py:64-69 (np.pad "reflect", multi-bounce when the pad exceeds the image) fused
 * with loader/infer_loader.py:57-69 (patch slicing): out[i] = padded[tl[i] : tl[i] + (ph,pw)]
 * without materialising the padded image. img: u8 [H,W,3]; tl_yx: HOST int32 [n][2], top-left
 * of each patch in padded coordinates (info_list[:,0,0] of _prepare_patching); out: u8
 * [n,ph,pw,3]. */
int cerb_extract_patches(cerb_ctx* ctx, const uint8_t* img, int H, int W, int pad_t, int pad_l,
                         const int32_t* tl_yx, int n, int ph, int pw, uint8_t* out, int flags);
/* flags bit 2 (value 4): out-of-image pixels are ZERO instead of reflected - the WSI loader's
 * read_bounds(..., pad_constant_values=0) (infer/wsi.py:936-942, tiatoolbox WSIStreamDataset). */

/* ---- WSI plumbing (infer/wsi.py) -------------------------------------------
 * The reference merges patch predictions into per-head float32 memmaps on disk
 * (tiatoolbox merge_prediction, infer/wsi.py:463,615); here ONE device-resident canvas
 * [H,W,C] (same channel layout as the per-patch canvas) receives them. With the reference's
 * stride == patch_output_shape each pixel is written once, so the running average is a clipped
 * write. patches_dev: f32 [n,oh,ow,C] (device); tl_yx: HOST int32 [n][2] output top-left. */
int cerb_scatter_patches(cerb_ctx* ctx, const float* patches_dev, int n, int oh, int ow, int C,
                         const int32_t* tl_yx, float* canvas_dev, int H, int W);
/* Copies the [h,w] window at (y0,x0) of a device image with px_bytes bytes per pixel into a
 * contiguous buffer (flags bit 1: dst is device memory, else host and the call synchronises). */
int cerb_crop2d(cerb_ctx* ctx, const void* src_dev, int H, int W, int px_bytes, int y0, int x0,
                int h, int w, void* dst, int flags);
/* infer/wsi.py:763-788 for one tissue region: out = cv2.resize(crop[..., chans] * mask, (0,0),
 * fx=0.5, fy=0.5) (float32; the arithmetic OpenCV 4.x + IPP performs for k = 1, 3, 4 channels and
 * OpenCV's own area path for k = 2: see csrc/tiles.cu). mask_host: u8 [h,w] 0/1 or NULL;
 * out_dev: f32 [oh,ow,k] with oh = cvRound(h/2), ow = cvRound(w/2). */
int cerb_region_half(cerb_ctx* ctx, const float* canvas_dev, int H, int W, int C, int y0, int x0,
                     int h, int w, const uint8_t* mask_host, const int32_t* chans, int k,
                     float* out_dev, int oh, int ow);
/* cv2.resize(canvas[..., ch], (0,0), fx=scale, fy=scale, INTER_NEAREST) into host memory
 * (infer/wsi.py:694-702, the Patch-Class map at 0.25). */
int cerb_nearest_channel(cerb_ctx* ctx, const float* canvas_dev, int H, int W, int C, int ch,
                         double scale, float* out_host, int oh, int ow);

/* canvas[..., ch] of a device-resident f32 [H,W,C] canvas as a contiguous [H,W] plane: the type /
 * Patch-Class maps loader/postproc.py:401-405 and infer/tile.py:183-184 slice out of the stitched
 * canvas. flags bit 1: `out` is device memory (queued on the ctx stream), else host (synchronous). */
int cerb_channel_plane(cerb_ctx* ctx, const float* canvas_dev, int H, int W, int C, int ch,
                       float* out, int flags);

/* infer/tile.py:136-163: canvas[tl : tl + (oh,ow)] += patch (in list order), count likewise,
 * canvas / (count + 1e-8), crop [src_y : src_y + out_h, src_x : src_x + out_w].
 * patches: f32 [n,oh,ow,C]; tl_yx: HOST int32 [n][2] (output top-left in canvas coordinates);
 * out: f32 [out_h,out_w,C]. */
int cerb_stitch(cerb_ctx* ctx, const float* patches, int n, int oh, int ow, int C,
                const int32_t* tl_yx, int canvas_h, int canvas_w, int src_y, int src_x, int out_h,
                int out_w, float* out, int flags);

/* ---- instance post-processing (a17-a19), loader/postproc.py:383-407 -------------------
 * canvas: f32 [n,H,W,C]; channels ch0 (inner) and ch0+1 (contour) of the tissue are read
 * (idx_dict[tissue + "-INST"][0]). labels_out: int32 [n,H,W]. Same `flags` as above. */

/* __proc_nuclei (:352-381). any_fg_out (nullable, int32 [n]) is 0 for images whose
 * pre-erosion mask is empty: the reference returns a float64 zero map for those. */
int cerb_postproc_nuclei(cerb_ctx* ctx, const float* canvas, int n, int H, int W, int C, int ch0,
                         int32_t* labels_out, int32_t* any_fg_out, int flags);
/* Counters: "ws_images" / "ws_tie_fallbacks" / "ws_capacity_fallbacks" = tiles of <= 65536 px seen
 * by the component-parallel nuclei watershed / redone by the exact whole-tile emulation because two
 * markers of one component tie / because the tile exceeds the shared-memory pool (synchronises);
 * "ws_large_images" = images > 65536 px labelled by the component-parallel watershed,
 * "ws_large_tied_components" = mask components of those images in which two marker entries tied
 * (settled by flooding every order of the tied entries), "ws_large_fallbacks" = images in which
 * two such orders disagreed (or there were too many) and which were redone by the exact
 * whole-image emulation. -1 for an unknown name. */
int64_t cerb_ctx_stat(cerb_ctx* ctx, const char* name);

enum cerb_tissue { CERB_TISSUE_GLAND = 0, CERB_TISSUE_LUMEN = 1 };
/* __proc_gland (:270-309) / __proc_lumen (:312-350); ds_factor as in the reference
 * (structuring element size int((ksize_-1)*ds), min object size int(1000|150 * ds^2)). The
 * reference returns float64 maps; convert at the binding. */
int cerb_postproc_gland_lumen(cerb_ctx* ctx, const float* canvas, int n, int H, int W, int C,
                              int ch0, int tissue, double ds_factor, int32_t* labels_out,
                              int flags);

/* PostProcInstErodedMap.post_process (loader/postproc.py:147-265, the "IP-ERODED-3/11" target
 * codes of infer/tile.py:35-37): foreground = channel ch0 > 0.5,
 * remove_small_objects (1500 gland / 150 lumen / 8 nuclei), label, per instance dilate with
 * ELLIPSE 11 / 3 / 3 inside the box padded by 2k (pad skipped at the border), fill holes, paint.
 * tissue: 0 gland, 1 lumen, 2 nuclei. The reference returns a float64 map. */
int cerb_postproc_eroded_map(cerb_ctx* ctx, const float* canvas, int n, int H, int W, int C,
                             int ch0, int tissue, int32_t* labels_out, int flags);

/* infer/tile.py:187-191: lumen *= (gland > 0), both device int32 label maps. */
int cerb_mask_lumen(cerb_ctx* ctx, int32_t* lumen_dev, const int32_t* gland_dev, size_t elems);

/* ---- instance tables (a20 / SURVEY 8f-2): loader/postproc.py:12-98 get_inst_info_dict and
 * tiatoolbox HoVerNet.get_instance_info (infer/wsi.py:150) -------------------------------------
 * labels: int32 [H,W]; type_map: f32 [H,W] of class ids (multiples of 0.25 in [0,16): integers, or
 * their 2x2 means after the WSI path's cv2.resize(fx=0.5), infer/wsi.py:773-790), or NULL (flags
 * bit 0: both are device memory). `up` >= 1 is the nearest-neighbour cv2.resize(fx=up, fy=up) the tile
 * mode applies to both maps before the call (infer/tile.py:196-201): all results are in the
 * upsampled image, which is never materialised. Synchronous; the table stays in the ctx until
 * the next call. Returns the number of instances (every id > 0 that occurs), the total number of
 * contour points, and whether a 0 pixel exists (np.unique(inst_map)[1:] drops the smallest id
 * when there is none). Fails if type_map holds any other value. */
int cerb_inst_info(cerb_ctx* ctx, const int32_t* labels, int H, int W, const float* type_map,
                   int up, int flags, int32_t* n_inst, int64_t* n_points, int32_t* any_background);
/* Host copies of the table of the last cerb_inst_info call, rows in ascending id order (any
 * pointer may be NULL):
 *   ids [n]; box [n][4] = rmin, cmin, rmax, cmax (max exclusive, misc/utils.py:82-91);
 *   moments [n][3] = cv2.moments m00, m10, m01 of the box crop (exact integers);
 *   type [n][2] = 4 x the majority type value by the rule of postproc.py:60-68 and its pixel
 *   count (-1, 0 without a type map); contour_off [n+1], contour_xy [n_points][2] = (x, y) image coordinates
 *   of cv2.findContours(crop, RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0] + box origin. */
int cerb_inst_info_read(cerb_ctx* ctx, int32_t* ids, int32_t* box, int64_t* moments, int32_t* type,
                        int64_t* contour_off, int32_t* contour_xy);

/* Host-only: the pickle stream (protocol-2 opcodes) of the dict items
 *   uid -> {"box": int64[4], "centroid": float64[2], "contour": int64[k,2], "prob": float|None,
 *           "type": int|None}
 * of n instances held as arrays - the `.dat` instance table of infer/wsi.py:853 (joblib.dump of a
 * dict of half a million small dicts per slide) written without creating a Python object per
 * instance. uid_hex: n x 32 ASCII characters; contour_off: int64 [n+1] offsets into contour_xy
 * (int64 [.,2]); prob / type: NULL writes None. memo: 11 pickle memo slots the caller's stream has
 * defined for (_reconstruct, ndarray, (0,), b"b", dtype int64, dtype float64, "box", "centroid",
 * "contour", "prob", "type") - see cerberus_b200/infer/dat_writer.py. Returns the stream length in
 * bytes (written if out != NULL and out_cap suffices) or a negative value on bad input. */
int64_t cerb_pickle_instances(const char* uid_hex, const int64_t* box, const double* centroid,
                              const int64_t* contour_off, const int64_t* contour_xy,
                              const double* prob, const int64_t* type, int64_t n,
                              const uint8_t* memo, uint8_t* out, int64_t out_cap);

/* cv2.getStructuringElement(MORPH_ELLIPSE, (k,k)) as row runs [j1[i], j2[i]) (1 <= k <= 32). */
int cerb_ellipse_rows(int k, int32_t* j1, int32_t* j2);

/* ---- copy stream: overlap host<->device traffic of step k+1 / k-1 with the compute of step k.
 * Each ctx owns an upload stream and a download stream used only by these calls.
 *   cerb_copy_async : cudaMemcpyAsync, kind 1 = H2D on the upload stream, 2 = D2H on the download
 *                     stream; host memory should come from cerb_host_alloc.
 *   cerb_stream_order(ctx, 0): later work on the COMPUTE stream waits for everything queued so far
 *                     on the upload stream; (ctx, 1): later work on the DOWNLOAD stream waits for
 *                     the compute stream.
 *   cerb_copy_mark / cerb_copy_wait : record an event (slot 0..7) on the download stream / block
 *                     the host until it has fired.
 *   cerb_copy_sync  : blocks the host until both copy streams are idle. */
int cerb_copy_async(cerb_ctx* ctx, void* dst, const void* src, size_t bytes, int kind);
int cerb_stream_order(cerb_ctx* ctx, int download_waits_for_compute);
int cerb_copy_mark(cerb_ctx* ctx, int slot);
int cerb_copy_wait(cerb_ctx* ctx, int slot);
int cerb_copy_sync(cerb_ctx* ctx);

/* Cross-context ordering on one device: work queued later on `waiter`'s compute stream waits for
 * everything queued so far on `signal`'s compute stream (lets a second ctx run the
 * post-processing of batch k while the first runs the forward of batch k+1). */
int cerb_ctx_wait(cerb_ctx* waiter, cerb_ctx* signal);
/* Finer grained: cerb_ctx_mark records event `slot` (0..7) on ctx's compute stream;
 * cerb_ctx_wait_mark makes work queued later on `waiter`'s compute stream wait for that event of
 * `signal` (no-op if it was never recorded). Lets a pipeline with several result slots wait for
 * the last reader of ONE slot instead of the whole stream. */
int cerb_ctx_mark(cerb_ctx* ctx, int slot);
int cerb_ctx_wait_mark(cerb_ctx* waiter, cerb_ctx* signal, int slot);

/* Pinned (page-locked) host memory for fast asynchronous H2D / D2H copies. */
void* cerb_host_alloc(size_t bytes);
void cerb_host_free(void* p);

/* Device scratch helpers for callers that keep data resident between calls. */
void* cerb_dev_alloc(cerb_ctx* ctx, size_t bytes);
int cerb_dev_free(cerb_ctx* ctx, void* p);
int cerb_memcpy(cerb_ctx* ctx, void* dst, const void* src, size_t bytes, int kind /*1 H2D, 2 D2H, 3 D2D*/);

#ifdef __cplusplus
}
#endif
#endif /* CERBERUS_B200_H */
