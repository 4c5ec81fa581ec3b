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


POSTPROC_IDX = {"Lumen-INST": [0, 2], "Gland-INST": [2, 4], "Nuclei-INST": [4, 6],
                "Nuclei-TYPE": [6, 7], "Gland-TYPE": [7, 8], "Patch-Class": [8, 9]}
POSTPROC_CH = {"Lumen": 0, "Gland": 2, "Nuclei": 4}


def postproc_cases():
    """(name, tissue, ds, field) — fields are quantised to multiples of 2^-12 so that the
    stored uint16 copy reproduces the float32 input exactly."""
    from cerberus_b200 import synth
    cases = []

    def q(f):
        return (np.round(f * 4096.0) / 4096.0).astype(np.float32)

    for seed in range(3):
        for tissue in ("Nuclei", "Gland", "Lumen"):
            cases.append(("smooth_s%d" % seed, tissue, 1.0, q(synth.postproc_field(256, 256, tissue, seed))))
    for tissue in ("Nuclei", "Gland", "Lumen"):
        cases.append(("smooth_ds05", tissue, 0.5, q(synth.postproc_field(256, 256, tissue, 7))))
        cases.append(("rect_s3", tissue, 1.0, q(synth.postproc_field(192, 320, tissue, 3))))
    for seed in range(6):
        cases.append(("adv_s%d" % seed, "Nuclei", 1.0, synth.adversarial_field(64, 80, seed)))
        cases.append(("adv_s%d" % seed, "Lumen", 1.0, synth.adversarial_field(96, 128, seed + 50, big=True)))
        cases.append(("adv_s%d" % seed, "Gland", 1.0, synth.adversarial_field(160, 200, seed + 100, big=True)))
        cases.append(("adv_ds05_s%d" % seed, "Gland", 0.5, synth.adversarial_field(96, 128, seed + 150, big=True)))
        cases.append(("adv_ds05_s%d" % seed, "Lumen", 0.5, synth.adversarial_field(64, 80, seed + 200, big=True)))
    cases.append(("empty", "Nuclei", 1.0, np.zeros((32, 48, 2), np.float32)))
    cases.append(("empty", "Gland", 1.0, np.zeros((32, 48, 2), np.float32)))
    cases.append(("empty", "Lumen", 1.0, np.zeros((32, 48, 2), np.float32)))
    full = np.zeros((40, 56, 2), np.float32)
    full[..., 0] = 1.0
    cases.append(("full", "Nuclei", 1.0, full))
    cases.append(("full", "Gland", 1.0, np.tile(full, (2, 2, 1))))
    cases.append(("full", "Lumen", 1.0, full))
    thin = np.zeros((24, 40, 2), np.float32)  # mask erodes to nothing: watershed on an empty mask
    thin[10:12, 5:35, 0] = 1.0
    cases.append(("thin", "Nuclei", 1.0, thin))
    return cases


def gen_postproc():
    from oracle import postproc_oracle as po
    ref_shim.install(po)
    from loader.postproc import PostProcInstErodedContourMap as PP
    rec = {}
    names = []
    for name, tissue, ds, field in postproc_cases():
        key = "%s/%s" % (name, tissue)
        raw = np.zeros(field.shape[:2] + (9,), np.float32)
        c0 = POSTPROC_CH[tissue]
        raw[..., c0:c0 + 2] = field
        inst, _ = PP.post_process(raw, POSTPROC_IDX, tissue, ds)
        qf = np.round(field * 4096.0)
        assert np.array_equal((qf / 4096.0).astype(np.float32), field), key
        rec[key + "/field_q12"] = qf.astype(np.uint16)
        rec[key + "/inst"] = inst.astype(np.uint16)
        assert inst.max() < 65536
        rec[key + "/dtype"] = np.array(str(inst.dtype))
        rec[key + "/ds"] = np.array(ds)
        names.append(key)
        print(key, field.shape, "instances", int(inst.max()), inst.dtype)
    rec["names"] = np.array(names)
    path = os.path.join(GOLD, "postproc.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path))


PATCHING_CASES = [(256, 256, 448, 144), (256, 256, 256, 256), (300, 517, 448, 144), (40, 60, 448, 144),
                  (500, 333, 256, 256)]


def gen_patching():
    """infer/tile.py:43-106 on seeded images: the patch grid and hashes of the padded image
    (multi-bounce reflect when the pad exceeds the image: 40x60 with a 152/… pad)."""
    import hashlib
    ref_shim.install()
    from infer.tile import _prepare_patching
    rec = {"cases": np.array(PATCHING_CASES)}
    for (h, w, i, o) in PATCHING_CASES:
        img = np.random.RandomState(h * 1000 + w).randint(0, 256, (h, w, 3)).astype(np.uint8)
        padded, info, src_pos = _prepare_patching(img, i, o, 0)
        key = "%dx%d_%d_%d" % (h, w, i, o)
        rec[key + "/info"] = info.astype(np.int32)
        rec[key + "/src_pos"] = np.array(src_pos)
        rec[key + "/padded_shape"] = np.array(padded.shape)
        rec[key + "/padded_sha1"] = np.array(hashlib.sha1(np.ascontiguousarray(padded).tobytes()).hexdigest())
        # three patches as the loader slices them (infer_loader.py:57-69)
        sel = [0, info.shape[0] // 3, info.shape[0] - 1]
        rec[key + "/sel"] = np.array(sel)
        rec[key + "/patch_sha1"] = np.array([
            hashlib.sha1(np.ascontiguousarray(
                padded[info[k, 0, 0, 0]:info[k, 0, 1, 0], info[k, 0, 0, 1]:info[k, 0, 1, 1]]).tobytes()).hexdigest()
            for k in sel])
        print(key, info.shape, padded.shape)
    path = os.path.join(GOLD, "patching.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path))


def tile_case_inputs(seed=0, h=300, w=380, in_size=448, out_size=144):
    """Synthetic per-patch network outputs cut from smooth fields, in the list-of-dicts format
    of models/run_desc.py:494-502, for infer/tile.py:_post_process_patches."""
    from cerberus_b200 import synth
    from cerberus_b200.infer.tile import patch_grid
    info, src_pos, (padt, padb, padl, padr) = patch_grid(h, w, in_size, out_size, 0)
    H, W = h + padt + padb, w + padl + padr
    q = lambda f: (np.round(f * 4096.0) / 4096.0).astype(np.float32)
    rng = np.random.RandomState(seed)
    heads = {
        "Nuclei-INST": q(synth.postproc_field(H, W, "Nuclei", seed)),
        "Nuclei-TYPE": rng.randint(0, 7, (H // 16 + 1, W // 16 + 1)).repeat(16, 0).repeat(16, 1)[:H, :W].astype(np.int64),
        "Gland-INST": q(synth.postproc_field(H, W, "Gland", seed + 1)),
        "Gland-TYPE": rng.randint(0, 3, (H // 32 + 1, W // 32 + 1)).repeat(32, 0).repeat(32, 1)[:H, :W].astype(np.int64),
        "Lumen-INST": q(synth.postproc_field(H, W, "Lumen", seed + 2)),
        "Patch-Class": rng.randint(0, 9, (H // 64 + 1, W // 64 + 1)).repeat(64, 0).repeat(64, 1)[:H, :W].astype(np.float32),
    }
    plist = []
    for k in range(info.shape[0]):
        (oy0, ox0), (oy1, ox1) = info[k, 1]
        pdata = {name: np.ascontiguousarray(v[oy0:oy1, ox0:ox1]) for name, v in heads.items()}
        plist.append((pdata, (info[k, 1, 0], info[k, 1, 1]), 0))
    image_info = {"src_pos": src_pos, "src_shape": (h, w), "name": "case%d" % seed,
                  "src_image": np.zeros((h, w, 3), np.uint8)}
    return plist, image_info


def gen_stitch():
    """infer/tile.py:109-212 (_post_process_patches) end to end on synthetic patch outputs."""
    from cerberus_b200 import synth
    from oracle import postproc_oracle as po
    ref_shim.install(po)
    from infer.tile import _post_process_patches
    margs = synth.model_args()
    codes = dict(synth.DEFAULT_REQ_TARGET_CODE)
    rec = {}
    for seed in (0, 1):
        plist, image_info = tile_case_inputs(seed)
        name, _, inst, info, types, pclass = _post_process_patches(
            plist, image_info, codes, ["gland", "lumen", "nuclei", "patch-class"], margs)
        for t, m in inst.items():
            rec["s%d/inst/%s" % (seed, t)] = m.astype(np.uint16)
            rec["s%d/inst_dtype/%s" % (seed, t)] = np.array(str(m.dtype))
            rec["s%d/ids/%s" % (seed, t)] = np.array(sorted(int(k) for k in info[t].keys()))
            rec["s%d/types/%s" % (seed, t)] = np.array([info[t][k].get("type", -1) for k in sorted(info[t].keys())])
            rec["s%d/centroids/%s" % (seed, t)] = np.array([info[t][k]["centroid"] for k in sorted(info[t].keys())]).reshape(-1, 2)
            if types[t] is not None:
                rec["s%d/type_map/%s" % (seed, t)] = types[t].astype(np.uint8)
        rec["s%d/pclass" % seed] = pclass.astype(np.uint8)
        print(seed, {t: int(m.max()) for t, m in inst.items()})
    path = os.path.join(GOLD, "tile_postproc.npz")
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
