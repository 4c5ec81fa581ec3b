"""Runs ONE convolution layer (ad-hoc plan over the C ABI) many times: device time by CUDA events,
and for the 64->64 3x3 kernel the in-kernel role attribution (wait cycles per role). Also the
target of `ncu --set full -k regex:conv` captures (profiles/).
  python tools/one_conv.py CIN COUT K H W BATCH [res] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from cerberus_b200.engine import Context, ForwardPlan, profile_ops  # noqa: E402
from cerberus_b200.pack import BlobBuilder, pack_conv  # noqa: E402
from tests.util import MiniModel, MiniSpec  # noqa: E402

SLOTS = ["producer wait empty", "mma wait tempty", "mma wait full", "mma issue", "epi0 wait tfull",
         "epi0 wait staging", "epi0 math+st.shared", "epi0 fence+bar+store", "mma warp total",
         "epi0 t0 wait store-read", "epi1 wait tfull", "epi1 wait staging", "epi1 math+st.shared",
         "epi1 fence+bar+store", "-", "epi1 t0 wait store-read"]


def main():
    cin, cout, k, h, w, n = [int(x) for x in sys.argv[1:7]]
    res = len(sys.argv) > 7 and sys.argv[7] == "res"
    head = len(sys.argv) > 7 and sys.argv[7].startswith("head")  # headC: fused classification head with C classes
    reps = int(sys.argv[8]) if len(sys.argv) > 8 else 20
    rng = np.random.RandomState(0)
    blob = BlobBuilder()
    wt = rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)
    layer = pack_conv(blob, wt, np.zeros(cout))
    spec = MiniSpec()
    ti = spec._tensor("in", n, h, w, cin)
    to = spec._tensor("out", n, h, w, cout)
    tr = spec._tensor("res", n, h, w, cout) if res else -1
    if head:
        ncls = int(sys.argv[7][4:] or 3)
        from cerberus_b200 import _lib as L_
        w2 = (rng.standard_normal((ncls, 96)) / 10.0).astype(np.float32)
        b2 = rng.uniform(-0.5, 0.5, ncls).astype(np.float32)
        tc = spec._tensor("canvas", n, h, w, 9, L_.CERB_F32)
        spec._conv(layer, ti, tc, relu=1, out_coff=0, aux_classes=ncls, aux_w_off=blob.add(w2),
                   aux_b_off=blob.add(b2), head_mode=L_.HEAD_INST)
    else:
        spec._conv(layer, ti, to, relu=1, residual=tr)
    prec = os.environ.get("CERB_ONE_PREC", "f16")
    ctx = Context(0, prec)
    ctx.set_option("kernel_prof", 1)
    ctx.set_option("use_graphs", 0)
    for kv in os.environ.get("CERB_OPTS", "").split(","):
        if "=" in kv:
            ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    plan = ForwardPlan(ctx, MiniModel(blob), 0, 0, 0, 0, 0, spec=spec)
    plan.write(ti, rng.standard_normal((n, h, w, cin)).astype(np.float16))
    if prec == "f16x2":
        plan.write(ti, (rng.standard_normal((n, h, w, cin)) * 1e-4).astype(np.float16), plane=1)
    if res:
        plan.write(tr, rng.standard_normal((n, h, w, cout)).astype(np.float16))
    for _ in range(3):
        plan.run()
    ctx.sync()
    prof = profile_ops(plan, reps=reps)
    ms = prof[0][1]
    fl = 2.0 * n * h * w * cout * cin * k * k
    by = n * h * w * (cin + cout * (2 if res else 1)) * 2
    print("%dx%d %d->%d k%d batch %d%s: %.4f ms  %.1f TFLOP/s  %.0f GB/s" % (
        h, w, cin, cout, k, n, " +res" if res else "", ms, fl / ms / 1e9, by / ms / 1e6))
    if True:
        c = ctx.read_prof()
        generic = not (cin == 64 and cout == 64 and k == 3)
        if generic:
            SLOTS[5:8] = ["epi0 ld+math(+sts)", "epi0 fence+bar+mma2+wait", "epi0 tail"]
            SLOTS[11:14] = ["epi1 ld+math(+sts)", "epi1 fence+bar+mma2+wait", "epi1 tail"]
            SLOTS[9] = "mma wait A-halo (conv3x3)"
            SLOTS[15] = "-"
        tot = c[:, 8].astype(np.float64)
        print("  per-CTA cycles (mean over %d CTAs; kernel total %.0f):" % (len(c), tot.mean()))
        for i, name in enumerate(SLOTS):
            if name != "-":
                print("    %-26s %10.0f  (%5.1f%%)" % (name, c[:, i].mean(), 100.0 * c[:, i].mean() / tot.mean()))
    plan.close()
    ctx.close()


if __name__ == "__main__":
    main()
