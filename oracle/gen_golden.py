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
    # resnet18 encoder (same BasicBlock family, models/backbone/resnet.py:292-302): name carries the backbone
    ("r18_256", ["Nuclei", "Gland#TYPE", "Patch-Class"], 1, 256, 256, 10),
]


def gen_forward():
    import torch
    from cerberus_b200 import synth
    ref_shim.install()
    from models.net_desc import create_model
    from models.run_desc import infer_step

    only = [a[len("forward:"):] for a in sys.argv[1:] if a.startswith("forward:")]
    for name, tasks, n, size, out, tseed in FORWARD_CASES:
        if only and name not in only:
            continue
        backbone = "resnet18" if name.startswith("r18") else "resnet34"
        args = synth.model_args(tasks, backbone=backbone)
        sd = synth.make_state_dict(args["considered_tasks"], seed=0, backbone=backbone)
        net = create_model(**args)
        net.load_state_dict(sd, strict=True)
        net.eval()
        tiles = synth.synthetic_tiles(n, size, size, seed=tseed)
        with torch.no_grad():
            logits = net(torch.from_numpy(tiles).float().permute(0, 3, 1, 2).contiguous())
        step = infer_step(torch.from_numpy(tiles), net, out, args["considered_tasks"])
        rec = {"n": n, "size": size, "out": out, "tile_seed": tseed, "ckpt_seed": 0, "backbone": np.array(backbone),
               "tasks": np.array(args["considered_tasks"]),
               "sd_check": np.array([float(sd["backbone.layer4.%d.bn2.running_var" % (synth.BACKBONE_BLOCKS[backbone][3] - 1)].double().sum()),
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


def gen_postproc_eroded():
    """PostProcInstErodedMap.post_process (loader/postproc.py:147-265), UNMODIFIED,
    on the inner channel of every ds = 1 field of postproc.npz, for all three tissues."""
    from oracle import postproc_oracle as po
    ref_shim.install(po)
    from loader.postproc import PostProcInstErodedMap as PP
    rec = {}
    names = []
    for name, tissue0, ds, field in postproc_cases():
        if ds != 1.0:
            continue
        for tissue in ("Gland", "Lumen", "Nuclei"):
            if tissue != tissue0 and not name.startswith("adv"):
                continue
            raw = np.ascontiguousarray(field[..., :1])
            inst, type_map = PP.post_process(raw, {tissue + "-INST": [0, 1]}, tissue)
            assert type_map is None and inst.max() < 65536
            key = "%s/%s/%s" % (name, tissue0, tissue)
            rec[key + "/inst"] = inst.astype(np.uint16)
            rec[key + "/dtype"] = np.array(str(inst.dtype))
            names.append(key)
            print(key, field.shape, "instances", int(inst.max()), inst.dtype)
    rec["names"] = np.array(names)
    path = os.path.join(GOLD, "postproc_eroded.npz")
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


def instinfo_cases():
    """Seeded label / type maps for the instance-table golden: (name, inst_map, type_map | None,
    ds_factor, up). `up` > 1: the reference sees cv2.resize(fx=up, fy=up, INTER_NEAREST) copies,
    as in infer/tile.py:196-201. Rebuilt identically by tests/test_gpu_instinfo.py."""
    from scipy import ndimage
    from oracle import postproc_oracle as po
    cases = []
    rng = np.random.RandomState(11)

    def field(h, w, sigma, seed):
        f = ndimage.gaussian_filter(np.random.RandomState(seed).randn(h, w), sigma)
        return (f - f.min()) / (f.max() - f.min())

    # watershed-like nuclei labels with a blocky type map
    f = field(160, 200, 3, 1)
    nuc = ndimage.label(f > 0.55)[0].astype(np.int32)
    typ = (field(160, 200, 12, 2) * 6.99).astype(np.int32).astype(np.float32)
    cases.append(("nuclei", nuc, typ, 1.0, 1))
    cases.append(("nuclei_up2", nuc, typ, 1.0, 2))
    cases.append(("nuclei_notype", nuc, None, 1.0, 1))
    # gland-like float64 labels: big blobs, sparse ids, holes, pieces split by a higher id
    g = ndimage.label(field(180, 150, 9, 3) > 0.5)[0].astype(np.float64) * 3
    g[field(180, 150, 4, 4) > 0.72] = 0          # holes
    g[60:75, :] = np.where(g[60:75, :] > 0, 50, 0)  # a band repainted with another id splits blobs
    gt = (field(180, 150, 20, 5) * 2.99).astype(np.int32).astype(np.float32)
    cases.append(("gland_up2", g, gt, 1.0, 2))
    # WSI flavour: half-resolution maps, type values are 2x2 means (multiples of 0.25), ds 0.5
    gq = (rng.randint(0, 9, g.shape) / 4.0).astype(np.float32)
    gq = ndimage.uniform_filter(gq, 1)
    cases.append(("gland_ds05_quarter", g, gq, 0.5, 1))
    # salt-and-pepper ids: nested holes / islands, one-pixel and two-pixel instances, lines
    noise = rng.randint(0, 4, (48, 61)).astype(np.int32)
    noise[rng.rand(48, 61) < 0.3] = 0
    cases.append(("noise", noise, rng.randint(0, 5, noise.shape).astype(np.float32), 1.0, 1))
    cases.append(("noise_up2", noise, None, 1.0, 2))
    # many medium components with holes per id (the LAST top-level component's border is wanted)
    patchy = (field(96, 120, 1.6, 6) > 0.5).astype(np.int32) * (1 + (np.arange(120)[None, :] // 47))
    cases.append(("patchy", patchy, None, 1.0, 1))
    cases.append(("patchy_up2", patchy, (patchy * 2 % 5).astype(np.float32), 1.0, 2))
    thin = np.zeros((20, 30), np.int32)
    thin[2, 3] = 1            # single pixel -> 1 point -> skipped
    thin[5, 2:12] = 2         # horizontal line -> 2 points -> skipped
    thin[8:15, 20] = 3        # vertical line
    thin[10:13, 5:9] = 4      # rectangle -> 4 points
    thin[16, 3] = thin[17, 4] = thin[18, 5] = 5  # diagonal (8-connected)
    thin[15:19, 12:16] = 6
    thin[16:18, 13:15] = 0    # ring
    thin[16, 13] = 7          # island in the ring's hole
    cases.append(("thin", thin, None, 1.0, 1))
    # no background pixel at all: np.unique(...)[1:] drops the smallest id
    full = (1 + (np.arange(24)[:, None] // 8) * 3 + np.arange(36)[None, :] // 12).astype(np.int32)
    cases.append(("no_background", full, (full % 3).astype(np.float32), 1.0, 1))
    # two instances, 0 is the most frequent type -> runner-up wins; equal counts -> smaller id
    tie = np.zeros((12, 12), np.int32)
    tie[1:5, 1:9] = 1
    tie[6:11, 2:10] = 2
    tt = np.zeros((12, 12), np.float32)
    tt[1:5, 1:3] = 3
    tt[1:5, 3:5] = 2          # instance 1: 16 x type 0, 8 x type 3, 8 x type 2 -> 2
    tt[6:11, 2:6] = 5
    tt[6:11, 6:10] = 4        # instance 2: 20 x 5, 20 x 4 -> 4
    cases.append(("type_rules", tie, tt, 1.0, 1))
    return cases


def gen_instinfo():
    """loader/postproc.py:12-98 (get_inst_info_dict), UNMODIFIED, on the maps above."""
    import cv2
    from oracle import postproc_oracle as po
    ref_shim.install(po)
    from loader.postproc import get_inst_info_dict
    rec = {}
    for name, inst, typ, ds, up in instinfo_cases():
        a, t = inst, typ
        if up != 1:
            a = cv2.resize(inst, (0, 0), fx=up, fy=up, interpolation=cv2.INTER_NEAREST)
            t = None if typ is None else cv2.resize(typ, (0, 0), fx=up, fy=up,
                                                    interpolation=cv2.INTER_NEAREST)
        info = get_inst_info_dict(a, t, ds)
        keys = list(info.keys())
        rec[name + "/inst_sum"] = np.array([float(np.asarray(inst, np.float64).sum()), inst.shape[0], inst.shape[1]])
        rec[name + "/ids"] = np.array([float(k) for k in keys])
        rec[name + "/box"] = np.array([info[k]["box"] for k in keys]).reshape(-1, 2, 2)
        rec[name + "/centroid"] = np.array([info[k]["centroid"] for k in keys], dtype=np.float64).reshape(-1, 2)
        cnt = [np.asarray(info[k]["contour"]).reshape(-1, 2) for k in keys]
        rec[name + "/contour_off"] = np.cumsum([0] + [len(c) for c in cnt])
        rec[name + "/contour"] = (np.concatenate(cnt) if cnt else np.zeros((0, 2))).astype(np.int32)
        if typ is not None:
            rec[name + "/type"] = np.array([info[k]["type"] for k in keys])
            rec[name + "/type_prob"] = np.array([info[k]["type_prob"] for k in keys], dtype=np.float64)
        print(name, inst.shape, "instances", len(keys), "points", int(rec[name + "/contour_off"][-1]))
    path = os.path.join(GOLD, "instinfo.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path))


def main():
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["forward", "patching", "postproc", "postproc_eroded", "stitch", "instinfo"]
    for w in which:
        fn = globals().get("gen_" + w.split(":")[0])  # "forward:r18_256" regenerates one forward case
        if fn is None:
            print("skip (not implemented):", w)
            continue
        fn()


if __name__ == "__main__":
    main()
