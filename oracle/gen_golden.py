"""ORACLE tooling — generates tests/golden/* by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shim.py) on seeded inputs. Run in the build
container only:   python oracle/gen_golden.py [forward] [patching] [postproc] [stitch]

Every golden file records the generator seed/arguments so the tests can rebuild identical
inputs without the reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import ref_shim  # noqa: E402


def _sub(a, s, o):
    return np.ascontiguousarray(a[..., o::s, o::s])


FORWARD_CASES = [
    # name, tasks (None = all six), n, in, out, tile seed
    ("six_256", None, 2, 256, 256, 7),
    ("six_448", None, 1, 448, 144, 8),
    ("nuclei_256", ["Nuclei"], 1, 256, 256, 9),
]


def gen_forward():
    import torch
    from cerberus_b200 import synth
    ref_shim.install()
    from models.net_desc import create_model
    from models.run_desc import infer_step

    for name, tasks, n, size, out, tseed in FORWARD_CASES:
        args = synth.model_args(tasks)
        sd = synth.make_state_dict(args["considered_tasks"], seed=0)
        net = create_model(**args)
        net.load_state_dict(sd, strict=True)
        net.eval()
        tiles = synth.synthetic_tiles(n, size, size, seed=tseed)
        with torch.no_grad():
            logits = net(torch.from_numpy(tiles).float().permute(0, 3, 1, 2).contiguous())
        step = infer_step(torch.from_numpy(tiles), net, out, args["considered_tasks"])
        rec = {"n": n, "size": size, "out": out, "tile_seed": tseed, "ckpt_seed": 0,
               "tasks": np.array(args["considered_tasks"]),
               "sd_check": np.array([float(sd["backbone.layer4.2.bn2.running_var"].double().sum()),
                                     float(sd["backbone.layer1.0.bn1.running_mean"].double().sum())])}
        for k, v in logits.items():
            v = v.numpy()
            rec["logits_sub/" + k] = _sub(v, 8, 3) if v.shape[-1] > 1 else v
            rec["logits_absmean/" + k] = np.abs(v).mean(axis=(0, 2, 3))
            rec["logits_mean/" + k] = v.astype(np.float64).mean(axis=(0, 2, 3))
        for k in step[0]:
            full = np.stack([s[k] for s in step])
            if full.dtype == np.int64:
                rec["step/" + k] = full.astype(np.uint8)  # class maps, full resolution
            elif full.ndim == 4:
                rec["step_sub/" + k] = np.ascontiguousarray(full[:, 2::4, 2::4, :])
            else:
                rec["step/" + k] = full.astype(np.uint8)  # Patch-Class plane (class index)
            rec["step_dtype/" + k] = np.array(str(step[0][k].dtype))
            rec["step_shape/" + k] = np.array(step[0][k].shape)
        path = os.path.join(GOLD, "forward_%s.npz" % name)
        np.savez_compressed(path, **rec)
        print("wrote", path, os.path.getsize(path))


def main():
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["forward", "patching", "postproc", "stitch"]
    for w in which:
        fn = globals().get("gen_" + w)
        if fn is None:
            print("skip (not implemented):", w)
            continue
        fn()


if __name__ == "__main__":
    main()
