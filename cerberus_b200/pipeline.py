"""Device-resident tail of the hot path: forward canvas -> instance label maps without
leaving HBM (BASELINE config 3: "full model + on-GPU watershed/CC post-processing").

For a batch of independent tiles whose network output covers the whole tile (in == out, the
bench workload) the stitched canvas of infer/tile.py:136-163 IS the per-patch canvas, so the
post-processing of infer/tile.py:173-191 runs directly on the plan's canvas tensor:
Nuclei / Gland / Lumen `post_process` (loader/postproc.py:383-407) + `lumen *= gland > 0`.
Only the int32 label maps and the type / patch-class planes are copied to the host.
"""
import ctypes

import numpy as np

from . import _lib


class DevicePostProc:
    TISSUES = ("Nuclei", "Gland", "Lumen")

    def __init__(self, ctx, model, n, h, w, ds_factor=1.0):
        self.ctx, self.model = ctx, model
        self.n, self.h, self.w = n, h, w
        self.ds = float(ds_factor)
        self.tissues = [t for t in self.TISSUES if (t + "-INST") in model.idx_dict
                        and t in model.considered_tasks]
        lib = ctx.lib
        self.nbytes = n * h * w * 4
        self.dev = {}
        for t in self.tissues:
            p = lib.cerb_dev_alloc(ctx.handle, self.nbytes)
            if not p:
                _lib.check(-1, "cerb_dev_alloc")
            self.dev[t] = p
        self.any_fg = np.zeros(n, dtype=np.int32)
        # pinned host buffers: the label maps are the step's result (D2H every step)
        self._pinned = {}
        self.host = {}
        for t in self.tissues:
            hp = lib.cerb_host_alloc(self.nbytes)
            if not hp:
                _lib.check(-1, "cerb_host_alloc")
            self._pinned[t] = hp
            buf = (ctypes.c_int32 * (n * h * w)).from_address(hp)
            self.host[t] = np.frombuffer(buf, dtype=np.int32).reshape(n, h, w)
        self.d2h_bytes = len(self.tissues) * self.nbytes

    def run(self, plan):
        """Asynchronous: leaves int32 label maps in device buffers."""
        lib, ctx = self.ctx.lib, self.ctx
        canvas = plan.tensor_ptr(plan.spec.canvas)
        C = self.model.canvas_c
        for t in self.tissues:
            ch0 = self.model.idx_dict[t + "-INST"][0]
            if t == "Nuclei":
                rc = lib.cerb_postproc_nuclei(ctx.handle, ctypes.c_void_p(canvas), self.n, self.h,
                                              self.w, C, ch0, ctypes.c_void_p(self.dev[t]), None, 3)
            else:
                rc = lib.cerb_postproc_gland_lumen(ctx.handle, ctypes.c_void_p(canvas), self.n,
                                                   self.h, self.w, C, ch0,
                                                   0 if t == "Gland" else 1, self.ds,
                                                   ctypes.c_void_p(self.dev[t]), 3)
            _lib.check(rc, "post-processing (%s)" % t)
        if "Gland" in self.dev and "Lumen" in self.dev:  # infer/tile.py:187-191
            _lib.check(lib.cerb_mask_lumen(ctx.handle, ctypes.c_void_p(self.dev["Lumen"]),
                                           ctypes.c_void_p(self.dev["Gland"]),
                                           self.n * self.h * self.w), "cerb_mask_lumen")

    def run_to_host(self, plan):
        """run() + D2H of every label map; returns {tissue: int32 [n,h,w]} (synchronous)."""
        self.run(plan)
        lib, ctx = self.ctx.lib, self.ctx
        for t in self.tissues:
            _lib.check(lib.cerb_memcpy(ctx.handle, self.host[t].ctypes.data_as(ctypes.c_void_p),
                                       ctypes.c_void_p(self.dev[t]), self.nbytes, 2), "cerb_memcpy")
        return self.host

    def close(self):
        for p in self.dev.values():
            self.ctx.lib.cerb_dev_free(self.ctx.handle, ctypes.c_void_p(p))
        self.dev = {}
        self.host = {}
        for hp in self._pinned.values():
            self.ctx.lib.cerb_host_free(ctypes.c_void_p(hp))
        self._pinned = {}
