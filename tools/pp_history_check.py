#!/usr/bin/env python
"""Does the large-image nuclei post-processing depend on what the context did before?
Runs cerb_postproc_nuclei on one big field after different call histories and compares every
result with the CPU oracle. Test tooling."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cerberus_b200 import _lib, synth  # noqa: E402
from cerberus_b200.engine import Context  # noqa: E402
from cerberus_b200.postproc import post_process_batch  # noqa: E402
from oracle import postproc_oracle as po  # noqa: E402


def field(h, w, seed):
    f = synth.postproc_field(h, w, "Nuclei", seed=seed)
    return np.ascontiguousarray(f, dtype=np.float32)


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    big = field(size, size, 1)
    t0 = time.time()
    ref = po.proc_nuclei(big)
    t_or = time.time() - t0
    small = field(1100, 900, 2)
    mid = field(2304, 2304, 3)
    out = {"size": size, "oracle_s": t_or, "instances": int(ref.max())}

    def run(ctx, a):
        lab, _ = post_process_batch(ctx, a[None], 0, "Nuclei")
        return lab[0]

    for name, hist in (("fresh", []), ("after_small", [small]), ("after_mid_small", [mid, small]),
                       ("after_big_twice", [big, big])):
        ctx = Context(0, "f16")
        for h in hist:
            run(ctx, h)
        got = run(ctx, big)
        out[name] = {"pixels_differing_from_oracle": int((got.astype(np.int64) != ref.astype(np.int64)).sum()),
                     "fallbacks": int(ctx.lib.cerb_ctx_stat(ctx.handle, b"ws_large_fallbacks")),
                     "tied": int(ctx.lib.cerb_ctx_stat(ctx.handle, b"ws_large_tied_components"))}
        ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
