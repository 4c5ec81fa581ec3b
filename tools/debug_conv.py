"""GPU bring-up diagnostics for the conv kernel: runs small cases in subprocesses (a trapped
kernel kills its CUDA context) and prints how the result differs from the reference."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(case):
    import torch
    import torch.nn.functional as F
    from cerberus_b200 import _lib
    from cerberus_b200.engine import Context, ForwardPlan
    from cerberus_b200.pack import BlobBuilder, pack_conv
    from tests.util import MiniModel, MiniSpec, nchw_to_nhwc, nhwc_to_nchw
    name, n, h, w, cin, cout, k, stride, wkind, prec = case
    rng = np.random.RandomState(0)
    x = rng.standard_normal((n, h, w, cin)).astype(np.float16).astype(np.float32)
    if wkind == "identity":
        wt = np.zeros((cout, cin, k, k), np.float32)
        for o in range(min(cout, cin)):
            wt[o, o, k // 2, k // 2] = 1.0
    else:
        wt = (rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float16).astype(np.float32)
    blob = BlobBuilder()
    layer = pack_conv(blob, wt.astype(np.float64), np.zeros(cout))
    spec = MiniSpec()
    oh = (h + 2 * (k // 2) - k) // stride + 1
    ow = (w + 2 * (k // 2) - k) // stride + 1
    ti = spec._tensor("in", n, h, w, cin)
    to = spec._tensor("out", n, oh, ow, cout)
    spec._conv(layer, ti, to, relu=0, stride=stride)
    ctx = Context(0, prec)
    plan = ForwardPlan(ctx, MiniModel(blob), 0, 0, 0, 0, 0, spec=spec)
    plan.write(ti, x.astype(np.float16), 0)
    plan.run()
    got = plan.read(to).astype(np.float32)
    ref = F.conv2d(torch.from_numpy(nhwc_to_nchw(x)), torch.from_numpy(wt), stride=stride, padding=k // 2)
    ref = nchw_to_nhwc(ref.numpy())
    err = np.abs(got - ref)
    bad = err > (2e-3 * np.abs(ref) + 2e-3)
    print("CASE %s: max err %.4g, bad frac %.4f, got absmean %.4g ref absmean %.4g" % (
        name, err.max(), bad.mean(), np.abs(got).mean(), np.abs(ref).mean()))
    if bad.any():
        np.set_printoptions(precision=3, suppress=True, linewidth=200)
        print(" bad per channel-group of 8:", bad.reshape(-1, cout // 8, 8).mean(axis=(0, 2)))
        print(" bad per row y:", bad.mean(axis=(0, 2, 3))[:32])
        print(" bad per col x:", bad.mean(axis=(0, 1, 3))[:32])
        print(" got[0,0,:4,:8]\n", got[0, 0, :4, :8], "\n ref[0,0,:4,:8]\n", ref[0, 0, :4, :8])
        if wkind == "identity":
            # where did channel c of pixel p land?
            g = got[0].reshape(-1, cout)
            r = ref[0].reshape(-1, cout)
            for pix in (0, 1, 8, 9):
                for ch in (0, 1, 8, 16, 33):
                    v = r[pix, ch]
                    loc = np.argwhere(np.isclose(g, v, atol=1e-3))
                    print("  ref[pix %d, ch %d]=%.3f found at" % (pix, ch, v), loc[:4].tolist())


CASES = [
    ("id_1x1_onetile", 1, 8, 16, 64, 64, 1, 1, "identity", "f16"),
    ("rand_1x1_onetile", 1, 8, 16, 64, 64, 1, 1, "rand", "f16"),
    ("rand_1x1_k128", 1, 8, 16, 128, 64, 1, 1, "rand", "f16"),
    ("rand_3x3", 1, 16, 16, 64, 64, 3, 1, "rand", "f16"),
    ("rand_3x3_multi", 2, 32, 32, 64, 128, 3, 1, "rand", "f16"),
    ("rand_3x3_s2", 1, 32, 32, 64, 64, 3, 2, "rand", "f16"),
    ("rand_3x3_x2", 1, 16, 16, 64, 64, 3, 1, "rand", "f16x2"),
]

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(CASES[int(sys.argv[1])])
    else:
        for i, c in enumerate(CASES):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), str(i)], timeout=120,
                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                out = r.stdout.strip().splitlines()
                print("\n".join(out[-40:]))
                if r.returncode != 0:
                    print("CASE %s: exit code %d" % (c[0], r.returncode))
            except subprocess.TimeoutExpired:
                print("CASE %s: TIMEOUT" % c[0])
