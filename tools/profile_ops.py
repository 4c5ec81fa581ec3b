"""Per-op device time of one forward step (CUDA events between ops), with algorithmic
FLOP/s for convolutions and bytes/s for the bandwidth-bound ops. Evidence for profiles/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from cerberus_b200 import _lib, synth  # noqa: E402
from cerberus_b200.engine import Context, ForwardPlan, profile_ops  # noqa: E402
from cerberus_b200.plan import PackedModel  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    prec = sys.argv[2] if len(sys.argv) > 2 else "f16"
    size = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    margs = synth.model_args()
    model = PackedModel(synth.make_state_dict(seed=0), margs)
    ctx = Context(0, prec)
    plan = ForwardPlan(ctx, model, batch, size, size, size, size)
    tiles = synth.synthetic_tiles(batch, size, size, seed=1)
    plan.run(tiles)
    ctx.sync()
    for _ in range(2):
        plan.run()
    ctx.sync()
    prof = profile_ops(plan, reps=5)
    spec = plan.spec
    tot = 0.0
    print("%3s %-8s %-28s %9s %9s %9s" % ("#", "kind", "shape", "ms", "TFLOP/s", "GB/s"))
    agg = {}
    for i, ((kind, ms), op) in enumerate(zip(prof, spec.ops)):
        tot += ms
        desc, tf, gbs = "", "", ""
        if kind == "conv":
            _, n, h, w, c, _ = spec.tensors[op["out"]]
            cin = 3 if op["stem"] else op["in_c"]
            fl = 2.0 * n * h * w * op["cout"] * op["kh"] * op["kw"] * cin
            desc = "%dx%d %d->%d k%d s%d%s" % (h, w, cin, op["cout"], op["kh"], op["stride"],
                                                " +res" if op["in1"] >= 0 else "")
            tf = "%.1f" % (fl / ms / 1e9)
            es = 2 * (2 if prec == "f16x2" else 1)
            _, _, ih, iw, _, _ = spec.tensors[op["in0"]]
            by = n * (ih * iw * (8 if op["stem"] else op["in_c"]) + h * w * op["cout"] * (2 if op["in1"] >= 0 else 1)) * es
            gbs = "%.0f" % (by / ms / 1e6)
            key = desc
        else:
            _, n, h, w, c, _ = spec.tensors[op["out"]]
            desc = "%dx%dx%d" % (h, w, c)
            key = kind + " " + desc
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ms
        print("%3d %-8s %-28s %9.4f %9s %9s" % (i, kind, desc, ms, tf, gbs))
    print("total %.3f ms for %d tiles -> %.0f tiles/s" % (tot, batch, batch / tot * 1e3))
    print("\naggregated:")
    for k, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-40s x%-3d %8.3f ms  %5.1f%%" % (k, cnt, ms, 100 * ms / tot))


if __name__ == "__main__":
    main()
