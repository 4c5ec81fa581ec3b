"""Host-side engine over the C ABI: device context, forward plans, and the `run_step`
drop-in (infer/base.py:51-53 -> models/run_desc.py:439-502).

PyTorch is used for `torch.load` of the checkpoint only; all device work goes through
libcerberus_b200.so (hand-written sm_100a kernels). There is no CPU fallback.
"""
import ctypes
import os
from collections import OrderedDict

import numpy as np

from . import _lib
from .plan import HEAD_NAME_MAP, PackedModel, PlanSpec

_DT_NP = {_lib.CERB_U8: np.uint8, _lib.CERB_F16: np.float16, _lib.CERB_F32: np.float32,
          _lib.CERB_I32: np.int32}


class Context:
    """One CUDA device + stream (cerb_ctx). precision: 'f16' (throughput) or 'f16x2'
    (hi+lo split operands, ~fp32 accuracy; the 1e-3 logit gate runs in this mode)."""

    def __init__(self, device=0, precision="f16"):
        self.lib = _lib.load()
        prec = {"f16": _lib.CERB_PREC_F16, "f16x2": _lib.CERB_PREC_F16X2}[precision]
        self.precision = precision
        self.device = device
        h = ctypes.c_void_p()
        _lib.check(self.lib.cerb_ctx_create(device, prec, ctypes.byref(h)), "cerb_ctx_create")
        self.handle = h
        # 64->64 3x3 kernel: 3 = even/odd N = 128 formulation (csrc/conv64x.cu, fastest); 1 = single
        # halo slab addressed through shifted descriptors (csrc/conv64.cu; also what the opt-in
        # fused variants CERB_FUSE_UPADD / CERB_FUSE_TAIL need); 0 / 2 = other halo layouts of
        # conv64.cu (tools/conv64_modes.py); -1 = generic kernel
        self.conv64_mode = int(os.environ.get("CERB_CONV64_MODE", "3"))
        self.set_option("conv64_mode", self.conv64_mode)
        if os.environ.get("CERB_CONV64S") is not None:  # split-precision 64->64 kernel (A/B switch)
            self.set_option("conv64s", int(os.environ["CERB_CONV64S"]))
        if os.environ.get("CERB_CONV3_PAIR") is not None:  # CTA-pair kernel for the wide 3x3 layers
            self.set_option("conv3_pair", int(os.environ["CERB_CONV3_PAIR"]))
        if os.environ.get("CERB_FUSE_UPADD") is not None:  # default on with the conv64x kernel (A/B switch)
            self.set_option("fuse_upadd", int(os.environ["CERB_FUSE_UPADD"]))
        if os.environ.get("CERB_CONV3_CHAIN") is not None:  # layer chains of the pair kernel (A/B switch)
            self.set_option("conv3_chain", int(os.environ["CERB_CONV3_CHAIN"]))
        if os.environ.get("CERB_DYN_SCHED") is not None:
            self.set_option("dyn_sched", int(os.environ["CERB_DYN_SCHED"]))
        if os.environ.get("CERB_USE_PDL") is not None:  # A/B switch for the PDL launches
            self.set_option("use_pdl", int(os.environ["CERB_USE_PDL"]))

    def set_option(self, name, value):
        _lib.check(self.lib.cerb_ctx_set_option(self.handle, name.encode(), int(value)),
                   "cerb_ctx_set_option")
        if name == "conv64_mode":
            self.conv64_mode = int(value)

    def sync(self):
        _lib.check(self.lib.cerb_ctx_sync(self.handle), "cerb_ctx_sync")

    def read_prof(self, n_ctas=148, reset=True):
        """Per-CTA wait-cycle counters of the last 64->64 launch (option kernel_prof): int64 [n_ctas,16]."""
        buf = (ctypes.c_int64 * (n_ctas * 16))()
        _lib.check(self.lib.cerb_ctx_read_prof(self.handle, buf, n_ctas * 16, int(reset)), "cerb_ctx_read_prof")
        return np.frombuffer(buf, dtype=np.int64).reshape(n_ctas, 16).copy()

    @property
    def launch_count(self):
        return int(self.lib.cerb_ctx_launch_count(self.handle))

    @property
    def stream(self):
        return self.lib.cerb_ctx_stream(self.handle)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.cerb_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ForwardPlan:
    """A compiled forward for one [N,H,W] batch shape (cerb_plan)."""

    def __init__(self, ctx, model, n, h, w, out_h, out_w, want_logits=False, spec=None):
        self.ctx = ctx
        self.model = model
        if spec is None:
            # The upsample+add fusion (fp16 64->64 kernel, conv64_mode 1) is exact but measured
            # SLOWER than upadd_kernel + conv64 on B200 (0.38 vs 0.34 ms per 256^2 layer: its eight
            # producer warps are instruction-bound), so it is opt-in.
            # With the default kernel (conv64_mode 3, csrc/conv64x.cu) the LIBRARY folds those UPADD ops
            # into the convolution (option fuse_upadd, default on: fix-up warps add the upsampled
            # tensor to the TMA-loaded skip halo) - nothing to do in the op list.
            fuse_up = (ctx.precision == "f16" and ctx.conv64_mode == 1
                       and os.environ.get("CERB_FUSE_UPADD", "0") == "1")
            # Last decoder conv + output head in one kernel (csrc/conv64.cu, fp16 mode only): exact
            # to fp16 rounding and saves 2 x 268 MB of HBM traffic per decoder at batch 32, but
            # measured SLOWER on B200 (0.315 ms vs 0.150 + 0.143 ms): the two tail MMAs of a tile queue
            # behind the next tile's 36 main MMAs in the in-order tensor pipe and two epilogue groups
            # cannot hide those round trips. Opt-in until the tail is software-pipelined.
            fuse_tail = (ctx.precision == "f16" and ctx.conv64_mode == 1
                         and os.environ.get("CERB_FUSE_TAIL", "0") == "1")
            # CERB_LEVEL_SYNC=1: decoders level by level with one grouped UPADD per level (A/B switch;
            # measured slower, see PlanSpec._decoders_level_sync)
            spec = PlanSpec(model, n, h, w, out_h, out_w, want_logits, fuse_upadd=fuse_up,
                            fuse_tail=fuse_tail,
                            level_sync=os.environ.get("CERB_LEVEL_SYNC", "0") == "1")
        self.spec = spec
        td, ops = self.spec.c_arrays()
        blob = model.blob
        h_ = ctypes.c_void_p()
        _lib.check(ctx.lib.cerb_plan_create(ctx.handle, td, len(td), ops, len(ops),
                                            blob.ctypes.data_as(ctypes.c_void_p), blob.nbytes,
                                            ctypes.byref(h_)), "cerb_plan_create")
        self.handle = h_
        self.n_launches = len(ops)

    def run(self, batch_u8=None, device_ptr=None):
        """Asynchronous. batch_u8: host uint8 [N,H,W,3] (copied H2D inside the call);
        device_ptr: raw device address of an already resident batch; neither: re-run on the
        plan's current input tensor."""
        lib = self.ctx.lib
        if device_ptr is not None:
            rc = lib.cerb_plan_run(self.handle, ctypes.c_void_p(device_ptr), 1)
        elif batch_u8 is not None:
            s = self.spec
            a = np.ascontiguousarray(batch_u8, dtype=np.uint8)
            if a.shape != (s.n, s.h, s.w, 3):
                raise ValueError("batch shape %r does not match the plan (%d,%d,%d,3)"
                                 % (a.shape, s.n, s.h, s.w))
            self._keep = a
            rc = lib.cerb_plan_run(self.handle, a.ctypes.data_as(ctypes.c_void_p), 0)
        else:
            rc = lib.cerb_plan_run(self.handle, None, 0)
        _lib.check(rc, "cerb_plan_run")

    def tensor_ptr(self, tid, plane=0):
        if isinstance(tid, str):
            tid = self.spec.named[tid]
        return self.ctx.lib.cerb_plan_tensor_ptr(self.handle, tid, plane)

    def read(self, tid, plane=0, combine=True):
        """Synchronous D2H of a plan tensor as a numpy NHWC array. fp16 tensors in f16x2 mode
        are returned as float32 hi+lo when `combine`."""
        if isinstance(tid, str):
            tid = self.spec.named[tid]
        _, n, h, w, c, dt = self.spec.tensors[tid]
        out = np.empty((n, h, w, c), dtype=_DT_NP[dt])
        _lib.check(self.ctx.lib.cerb_plan_read_tensor(self.handle, tid, plane,
                                                      out.ctypes.data_as(ctypes.c_void_p),
                                                      out.nbytes), "cerb_plan_read_tensor")
        if dt == _lib.CERB_F16 and combine and plane == 0 and self.ctx.precision == "f16x2":
            lo = self.read(tid, plane=1, combine=False)
            return out.astype(np.float32) + lo.astype(np.float32)
        return out

    def write(self, tid, arr, plane=0):
        if isinstance(tid, str):
            tid = self.spec.named[tid]
        _, n, h, w, c, dt = self.spec.tensors[tid]
        a = np.ascontiguousarray(arr, dtype=_DT_NP[dt]).reshape(n, h, w, c)
        _lib.check(self.ctx.lib.cerb_plan_write_tensor(self.handle, tid, plane,
                                                       a.ctypes.data_as(ctypes.c_void_p), a.nbytes),
                   "cerb_plan_write_tensor")

    def read_canvas(self):
        return self.read(self.spec.canvas)

    def read_logits(self):
        """dict head -> float32 NHWC logits (only when the plan was built with want_logits)."""
        return OrderedDict((k, self.read(t)) for k, t in self.spec.logit_tensors.items())

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.cerb_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_KIND_NAMES = {_lib.OP_PREP: "prep", _lib.OP_CONV: "conv", _lib.OP_MAXPOOL: "maxpool",
               _lib.OP_UPADD: "upadd", _lib.OP_HEAD: "head", _lib.OP_PCLASS: "pclass"}


def profile_ops(plan, device_ptr=None, reps=3):
    """Per-op device time (CUDA events on the ctx stream). Returns [(kind_name, ms), ...]."""
    lib = plan.ctx.lib
    if device_ptr is not None:
        plan.run(device_ptr=device_ptr)
        plan.ctx.sync()
    n = lib.cerb_plan_num_ops(plan.handle)
    ms = (ctypes.c_float * n)()
    kinds = (ctypes.c_int32 * n)()
    _lib.check(lib.cerb_plan_profile(plan.handle, reps, ms, kinds), "cerb_plan_profile")
    return [(_KIND_NAMES.get(kinds[i], str(kinds[i])), float(ms[i])) for i in range(n)]


def canvas_to_step_outputs(canvas, model):
    """[N,oh,ow,C] float32 patch canvas -> the list-of-dicts contract of
    models/run_desc.py:480-502 (insertion order = considered_tasks order; *-INST (h,w,2)
    float32, *-TYPE (h,w) int64, Patch-Class (h,w) float32)."""
    n = canvas.shape[0]
    per_head = OrderedDict()
    for task in model.considered_tasks:
        name = HEAD_NAME_MAP[task]
        lo, hi = model.idx_dict[name]
        v = canvas[..., lo:hi]
        if name.endswith("-TYPE"):
            v = v[..., 0].astype(np.int64)
        elif name == "Patch-Class":
            v = v[..., 0]
        per_head[name] = v
    return [OrderedDict((k, v[i]) for k, v in per_head.items()) for i in range(n)]


class CModel:
    """The model-level C ABI (cerb_model_create / cerb_forward): the op graph is built inside the
    library, Python only hands over the packed blob and its layer table. Same kernels, same
    results as ForwardPlan over PlanSpec; this is the entry a non-Python host uses."""

    def __init__(self, ctx, packed, handle=None):
        from .plan import c_model_tables
        self.ctx, self.packed = ctx, packed
        if handle is not None:  # received by cerb_bcast_weights
            self.handle = handle
            return
        desc, layers = c_model_tables(packed)
        self._keep = (desc, layers)
        h = ctypes.c_void_p()
        blob = packed.blob
        _lib.check(ctx.lib.cerb_model_create(ctx.handle, ctypes.byref(desc), layers, len(layers),
                                             blob.ctypes.data_as(ctypes.c_void_p), blob.nbytes,
                                             ctypes.byref(h)), "cerb_model_create")
        self.handle = h

    def forward(self, batch_u8, out_h, out_w, canvas_c):
        """uint8 [n,h,w,3] host batch -> float32 [n,out_h,out_w,canvas_c] canvas (synchronous)."""
        a = np.ascontiguousarray(batch_u8, dtype=np.uint8)
        n, h, w, _ = a.shape
        out = np.empty((n, out_h, out_w, canvas_c), dtype=np.float32)
        _lib.check(self.ctx.lib.cerb_forward(self.handle, a.ctypes.data_as(ctypes.c_void_p), 0, n, h, w,
                                             out_h, out_w, out.ctypes.data_as(ctypes.c_void_p), 0),
                   "cerb_forward")
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.cerb_model_destroy(self.handle)
            self.handle = None


class Engine:
    """Model directory -> run_step. Plans are cached per (N,H,W,out) shape."""

    def __init__(self, state_dict, model_args, device=0, precision="f16", packed=None):
        """`packed`: an already folded + packed model (e.g. received by broadcast_packed_model)."""
        self.model = packed if packed is not None else PackedModel(state_dict, model_args)
        self.ctx = Context(device, precision)
        self._plans = {}

    def plan_for(self, n, h, w, out_h, out_w, want_logits=False):
        key = (n, h, w, out_h, out_w, want_logits)
        if key not in self._plans:
            self._plans[key] = ForwardPlan(self.ctx, self.model, n, h, w, out_h, out_w, want_logits)
        return self._plans[key]

    def run_step(self, input_batch, output_shape):
        """Drop-in for infer/base.py:51-53. input_batch: uint8 [N,h,w,3] torch tensor or
        ndarray; returns list[N] of dict[str -> ndarray]."""
        if hasattr(input_batch, "numpy"):
            input_batch = input_batch.numpy()
        if not isinstance(output_shape, (list, tuple)):
            output_shape = [output_shape, output_shape]
        n, h, w, _ = input_batch.shape
        plan = self.plan_for(n, h, w, int(output_shape[0]), int(output_shape[1]))
        plan.run(input_batch)
        return canvas_to_step_outputs(plan.read_canvas(), self.model)

    def close(self):
        for p in self._plans.values():
            p.close()
        self._plans = {}
        self.ctx.close()
