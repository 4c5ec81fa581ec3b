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

    def run(self, plan, canvas_ptr=None):
        """Asynchronous: leaves int32 label maps in device buffers. `canvas_ptr`: device address
        of a [n,h,w,C] float32 canvas to read instead of the plan's own canvas tensor."""
        lib, ctx = self.ctx.lib, self.ctx
        canvas = canvas_ptr if canvas_ptr is not None else plan.tensor_ptr(plan.spec.canvas)
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


class TilePipeline:
    """Public streaming API for batches of independent tiles whose network output covers the
    tile (the bench workload): host uint8 batch in -> host int32 label maps out. Four streams
    are kept busy at once:
        upload   : H2D of batch k+1                         (engine ctx, cerb_copy_async kind 1)
        compute  : forward of batch k, canvas -> slot copy  (engine ctx)
        post     : post-processing of batches k-1, k-2      (a second ctx on the same device)
        download : D2H of their label maps                  (second ctx, cerb_copy_async kind 2)
    The post-processing kernels are latency-bound (one block per image / instance), so running
    them next to the next batches' convolutions hides most of their time. There are `depth`
    result slots (default 4) and results are handed back LAG = depth - 2 submits late: a batch
    whose nuclei watershed needs the exact whole-tile emulation (two markers of one touching-
    nuclei component with bit-equal values: ~13 ms for one image next to a running forward) then
    delays nothing as long as the average post-processing time stays below the forward time.

        pipe = TilePipeline(engine, n, h, w)
        for batch in batches:            # uint8 [n,h,w,3]
            done = pipe.submit(batch)    # -> labels of batch k - LAG (None for the first LAG)
        rest = pipe.flush()              # -> list of the remaining results, oldest first

    Returned dicts {tissue: int32 [n,h,w]} are views of pinned buffers that stay valid until the
    second-next submit()."""

    def __init__(self, engine, n, h, w, ds_factor=1.0, depth=4):
        import os
        from collections import deque
        from .engine import Context
        self.ctx, self.model = engine.ctx, engine.model
        self.lib = self.ctx.lib
        self.pctx = Context(self.ctx.device, self.ctx.precision)  # post-processing + download
        # The post-processing's watershed blocks need a whole SM each (>200 KB shared memory) and
        # take them from the persistent convolution kernels of the next forward. Those draw their
        # tiles from a global counter (dynamic scheduling), so a CTA that starts late just finds
        # less work; with the earlier static split every convolution launched meanwhile ran a
        # second wave, and a few SMs had to be reserved (CERB_POST_SMS, still available:
        # 6.39 ms/step with 4 reserved SMs vs 6.32 with none).
        self.post_sms = int(os.environ.get("CERB_POST_SMS", "0"))
        if self.post_sms > 0:
            self.ctx.set_option("conv_sms", 148 - self.post_sms)
        # Programmatic dependent launch lets the NEXT convolution's CTAs park on freed SMs while
        # they wait for their predecessor, which starves the post-processing blocks (measured with
        # a slow exact-watershed image in every 4th batch: 8.3 ms/step with PDL, 7.5 without).
        if os.environ.get("CERB_USE_PDL") is None:
            self.ctx.set_option("use_pdl", 0)
        self.depth = max(2, min(8, int(depth)))
        self.lag = self.depth - 2 if self.depth > 2 else 1
        self.plan = engine.plan_for(n, h, w, h, w)
        self.n, self.h, self.w = n, h, w
        self.in_bytes = n * h * w * 3
        self.canvas_bytes = n * h * w * self.model.canvas_c * 4
        # result slots (canvas copy + label maps): batch k-1.. are post-processed / downloaded
        # while batch k is computed
        self.post = [DevicePostProc(self.pctx, self.model, n, h, w, ds_factor)
                     for _ in range(self.depth)]
        self.canvas_copy = []
        for _ in range(self.depth):
            cp = self.lib.cerb_dev_alloc(self.ctx.handle, self.canvas_bytes)
            if not cp:
                _lib.check(-1, "TilePipeline canvas allocation")
            self.canvas_copy.append(cp)
        # three input slots: the upload of batch k never races the read of batch k-1 / k-2
        self.stage_host, self.stage_dev, self._views = [], [], []
        for _ in range(3):
            hp = self.lib.cerb_host_alloc(self.in_bytes)
            dp = self.lib.cerb_dev_alloc(self.ctx.handle, self.in_bytes)
            if not hp or not dp:
                _lib.check(-1, "TilePipeline staging allocation")
            self.stage_host.append(hp)
            self.stage_dev.append(dp)
            buf = (ctypes.c_uint8 * self.in_bytes).from_address(hp)
            self._views.append(np.frombuffer(buf, dtype=np.uint8).reshape(n, h, w, 3))
        self.k = 0
        self.pending = deque()  # result slots whose D2H is in flight, oldest first
        self.h2d_bytes = self.in_bytes
        self.d2h_bytes = self.post[0].d2h_bytes

    def submit(self, batch_u8=None, device_ptr=None, download=True):
        """Queues one batch. batch_u8: host uint8 [n,h,w,3]; or device_ptr: address of a batch
        already resident in HBM (no upload). download=False leaves the label maps on the device
        (self.post[slot].dev) and returns None."""
        lib, ctx, pctx = self.lib, self.ctx, self.pctx
        s_in, s_out = self.k % 3, self.k % self.depth
        if device_ptr is None:
            np.copyto(self._views[s_in], batch_u8)  # pageable -> pinned (host memcpy)
            _lib.check(lib.cerb_copy_async(ctx.handle, ctypes.c_void_p(self.stage_dev[s_in]),
                                           ctypes.c_void_p(self.stage_host[s_in]), self.in_bytes, 1),
                       "cerb_copy_async")
            _lib.check(lib.cerb_stream_order(ctx.handle, 0), "cerb_stream_order")  # compute waits H2D
            device_ptr = self.stage_dev[s_in]
        self.plan.run(device_ptr=device_ptr)
        # canvas -> slot copy on the compute stream, after the post-processing that last read THIS
        # slot (batch k - depth), not after everything queued on the post stream
        _lib.check(lib.cerb_ctx_wait_mark(ctx.handle, pctx.handle, s_out), "cerb_ctx_wait_mark")
        _lib.check(lib.cerb_memcpy(ctx.handle, ctypes.c_void_p(self.canvas_copy[s_out]),
                                   ctypes.c_void_p(self.plan.tensor_ptr(self.plan.spec.canvas)),
                                   self.canvas_bytes, 3), "cerb_memcpy")
        _lib.check(lib.cerb_ctx_wait(pctx.handle, ctx.handle), "cerb_ctx_wait")  # post waits compute
        post = self.post[s_out]
        post.run(self.plan, canvas_ptr=self.canvas_copy[s_out])
        _lib.check(lib.cerb_ctx_mark(pctx.handle, s_out), "cerb_ctx_mark")
        if not download:
            self.k += 1
            return None
        _lib.check(lib.cerb_stream_order(pctx.handle, 1), "cerb_stream_order")  # D2H waits for post
        for t in post.tissues:
            _lib.check(lib.cerb_copy_async(pctx.handle, post.host[t].ctypes.data_as(ctypes.c_void_p),
                                           ctypes.c_void_p(post.dev[t]), post.nbytes, 2),
                       "cerb_copy_async")
        _lib.check(lib.cerb_copy_mark(pctx.handle, s_out), "cerb_copy_mark")
        self.pending.append(s_out)
        self.k += 1
        done = None
        if len(self.pending) > self.lag:  # oldest batch: its D2H was queued `lag` submits ago
            slot = self.pending.popleft()
            _lib.check(lib.cerb_copy_wait(pctx.handle, slot), "cerb_copy_wait")
            done = self.post[slot].host
        return done

    def join_streams(self):
        """Makes the compute stream wait for the post stream (for timing with one end event)."""
        _lib.check(self.lib.cerb_ctx_wait(self.ctx.handle, self.pctx.handle), "cerb_ctx_wait")

    def flush(self):
        """Waits for everything queued; returns the results not handed out yet, oldest first."""
        out = []
        while self.pending:
            slot = self.pending.popleft()
            _lib.check(self.lib.cerb_copy_wait(self.pctx.handle, slot), "cerb_copy_wait")
            out.append(self.post[slot].host)
        self.ctx.sync()
        self.pctx.sync()
        return out

    @property
    def launch_count(self):
        return self.ctx.launch_count + self.pctx.launch_count

    def close(self):
        self.flush()
        for p in self.post:
            p.close()
        for cp in self.canvas_copy:
            self.lib.cerb_dev_free(self.ctx.handle, ctypes.c_void_p(cp))
        for hp in self.stage_host:
            self.lib.cerb_host_free(ctypes.c_void_p(hp))
        for dp in self.stage_dev:
            self.lib.cerb_dev_free(self.ctx.handle, ctypes.c_void_p(dp))
        self.stage_host, self.stage_dev, self.canvas_copy = [], [], []
        self.pctx.close()
