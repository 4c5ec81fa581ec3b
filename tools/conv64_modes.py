"""Hardware experiment: which halo layouts of csrc/conv64.cu produce correct results, and how
fast each is on the 256x256 64->64 layer (batch 32). Each mode runs in its own process."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(mode):
    import ctypes
    import torch
    import torch.nn.functional as F
    from cerberus_b200 import _lib
    from cerberus_b200.engine import Context, ForwardPlan, profile_ops
    from cerberus_b200.pack import BlobBuilder, pack_conv
    from tests.util import MiniModel, MiniSpec, nchw_to_nhwc, nhwc_to_nchw
    rng = np.random.RandomState(0)
    ok_all = True
    for (n, h, w, res) in ((1, 16, 8, False), (2, 48, 40, True), (1, 128, 128, False), (2, 50, 36, True)):
        x = rng.standard_normal((n, h, w, 64)).astype(np.float16).astype(np.float32)
        wt = (rng.standard_normal((64, 64, 3, 3)) / 24.0).astype(np.float16).astype(np.float32)
        b = rng.uniform(-0.5, 0.5, 64).astype(np.float32)
        r = rng.standard_normal((n, h, w, 64)).astype(np.float16).astype(np.float32) if res else None
        blob = BlobBuilder()
        layer = pack_conv(blob, wt.astype(np.float64), b.astype(np.float64))
        spec = MiniSpec()
        ti = spec._tensor("in", n, h, w, 64)
        to = spec._tensor("out", n, h, w, 64)
        tr = spec._tensor("res", n, h, w, 64) if res else -1
        spec._conv(layer, ti, to, relu=1, residual=tr)
        ctx = Context(0, "f16")
        ctx.set_option("conv64_mode", mode)
        plan = ForwardPlan(ctx, MiniModel(blob), 0, 0, 0, 0, 0, spec=spec)
        plan.write(ti, x.astype(np.float16))
        if res:
            plan.write(tr, r.astype(np.float16))
        plan.run()
        got = plan.read(to).astype(np.float32)
        ref = F.conv2d(torch.from_numpy(nhwc_to_nchw(x)).double(), torch.from_numpy(wt).double(),
                       torch.from_numpy(b).double(), padding=1)
        if res:
            ref = ref + torch.from_numpy(nhwc_to_nchw(r)).double()
        ref = nchw_to_nhwc(F.relu(ref).numpy())
        err = np.abs(got - ref)
        bad = float((err > 2e-3 * np.abs(ref) + 2e-3).mean())
        ok_all &= bad == 0.0
        print("mode %d case %s: max err %.4g bad frac %.4f" % (mode, (n, h, w, res), err.max(), bad))
        plan.close()
        ctx.close()
    # timing on the real layer shape
    n, h, w = 32, 256, 256
    blob = BlobBuilder()
    wt = (rng.standard_normal((64, 64, 3, 3)) / 24.0)
    layer = pack_conv(blob, wt, np.zeros(64))
    spec = MiniSpec()
    ti = spec._tensor("in", n, h, w, 64)
    to = spec._tensor("out", n, h, w, 64)
    spec._conv(layer, ti, to, relu=1)
    ctx = Context(0, "f16")
    ctx.set_option("conv64_mode", mode)
    plan = ForwardPlan(ctx, MiniModel(blob), 0, 0, 0, 0, 0, spec=spec)
    plan.run(); ctx.sync()
    prof = profile_ops(plan, reps=10)
    ms = prof[0][1]
    fl = 2.0 * n * h * w * 64 * 576
    print("mode %d: %s, 256x256 64->64 batch 32: %.4f ms = %.1f TFLOP/s" % (
        mode, "PASS" if ok_all else "FAIL", ms, fl / ms / 1e9))
    if mode >= 0:
        for dbg, what in ((1, "no global stores"), (4, "no epilogue math/stores"), (2, "one tap only"),
                          (6, "one tap, no epilogue"), (3, "one tap, no stores")):
            ctx.set_option("conv64_debug", dbg)
            pl2 = ForwardPlan(ctx, MiniModel(blob), 0, 0, 0, 0, 0, spec=spec)
            pl2.run(); ctx.sync()
            ms2 = profile_ops(pl2, reps=10)[0][1]
            print("   mode %d debug %d (%s): %.4f ms" % (mode, dbg, what, ms2))
            pl2.close()
        ctx.set_option("conv64_debug", 0)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(int(sys.argv[1]))
    else:
        for mode in (-1, 0, 1, 2):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), str(mode)], timeout=120,
                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                print("\n".join(r.stdout.strip().splitlines()[-8:]))
                if r.returncode != 0:
                    print("mode %d: exit code %d" % (mode, r.returncode))
            except subprocess.TimeoutExpired:
                print("mode %d: TIMEOUT" % mode)
