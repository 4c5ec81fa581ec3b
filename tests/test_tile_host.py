"""CPU: host-side tile logic (patch grid, reflect padding, CLI parsing, plan building, C-ABI
symbol table) against golden vectors produced by the reference's infer/tile.py."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest

from cerberus_b200 import _lib, synth
from cerberus_b200.cli import parse_usage
from cerberus_b200.infer.tile import _prepare_patching, patch_grid
from cerberus_b200.plan import PackedModel, PlanSpec, canvas_layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_prepare_patching_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "patching.npz"))
    for (h, w, i, o) in g["cases"]:
        key = "%dx%d_%d_%d" % (h, w, i, o)
        img = np.random.RandomState(h * 1000 + w).randint(0, 256, (h, w, 3)).astype(np.uint8)
        padded, info, src_pos = _prepare_patching(img, int(i), int(o), 0)
        assert np.array_equal(info, g[key + "/info"]), key
        assert list(src_pos) == list(g[key + "/src_pos"])
        assert tuple(padded.shape) == tuple(g[key + "/padded_shape"])
        assert hashlib.sha1(np.ascontiguousarray(padded).tobytes()).hexdigest() == str(g[key + "/padded_sha1"])
        info2, _, _ = patch_grid(int(h), int(w), int(i), int(o), 0)
        assert np.array_equal(info2, info)
        # duplicate grid quirk (tile.py:90-103): second half repeats the first
        half = info.shape[0] // 2
        assert np.array_equal(info[:half], info[half:])


def test_canvas_layout_matches_tile_py():
    idx, n = canvas_layout(synth.DEFAULT_DECODER_KWARGS)
    assert n == 9
    assert idx == {"Lumen-INST": [0, 2], "Gland-INST": [2, 4], "Nuclei-INST": [4, 6],
                   "Nuclei-TYPE": [6, 7], "Gland-TYPE": [7, 8], "Patch-Class": [8, 9]}


def test_plan_flops_match_survey(six_head_sd):
    """121.128 GFLOP per 256^2 tile (6 heads), 370.953 @448^2, 54.891 enc+Nuclei."""
    m = PackedModel(six_head_sd, synth.model_args())
    assert abs(PlanSpec(m, 1, 256, 256, 256, 256).conv_flops() / 1e9 - 121.128) < 1e-3
    assert abs(PlanSpec(m, 1, 448, 448, 144, 144).conv_flops() / 1e9 - 370.953) < 1e-3
    sd1 = synth.make_state_dict(["Nuclei"], seed=0)
    m1 = PackedModel(sd1, synth.model_args(["Nuclei"]))
    assert abs(PlanSpec(m1, 1, 256, 256, 256, 256).conv_flops() / 1e9 - 54.891) < 1e-3


def test_plan_never_writes_a_conv_input_in_place(six_head_sd):
    m = PackedModel(six_head_sd, synth.model_args())
    s = PlanSpec(m, 2, 256, 256, 256, 256)
    for op in s.ops:
        if op["kind"] == _lib.OP_CONV:
            assert op["out"] != op["in0"] and op["out"] != op["in1"]
    # skip features x0..x3 are never overwritten after they are produced
    last_write = {}
    for i, op in enumerate(s.ops):
        last_write[op["out"]] = i
    for name in ("x0", "x1", "x2", "x3", "x4"):
        tid = s.named[name]
        readers = [i for i, op in enumerate(s.ops) if tid in (op["in0"], op["in1"])]
        assert max(readers) > last_write[tid] or name == "x4"


def test_rejects_other_backbones(six_head_sd):
    args = dict(synth.model_args(), encoder_backbone_name="resnet50")
    with pytest.raises(ValueError):
        PackedModel(six_head_sd, args)


def test_module_prefix_is_stripped(six_head_sd):
    """infer/base.py:31-45: DataParallel checkpoints."""
    sd = {"module." + k: v for k, v in six_head_sd.items()}
    a = PackedModel(sd, synth.model_args())
    b = PackedModel(six_head_sd, synth.model_args())
    assert np.array_equal(a.blob, b.blob)


def test_cli_parser_has_reference_flags_and_defaults():
    doc = open(os.path.join(ROOT, "run_infer_tile.py")).read().split('"""')[1]
    a = parse_usage(doc, ["--model=/m", "--input_dir", "/in"])
    assert a["--gpu"] == "0" and a["--batch_size"] == "10" and a["--patch_input_shape"] == "448"
    assert a["--patch_output_shape"] == "144" and a["--output_dir"] == "output/"
    assert a["--nr_inference_workers"] == "0" and a["--model"] == "/m" and a["--input_dir"] == "/in"


def test_library_loads_and_exports_every_header_symbol(built_lib):
    """No GPU needed: the .so loads (static cudart, no libcuda link) and exports every function
    include/cerberus_b200.h declares; compute entry points fail loudly without a device."""
    hdr = open(os.path.join(ROOT, "include", "cerberus_b200.h")).read()
    declared = set(re.findall(r"\b(cerb_\w+)\s*\(", hdr))
    declared -= {"cerb_status", "cerb_dtype"}
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(built_lib, name), name
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    import torch
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        rc = built_lib.cerb_ctx_create(0, 0, ctypes.byref(h))
        assert rc == -4 and b"no CPU path" in built_lib.cerb_last_error()


def test_process_file_list_overlaps_loading_and_saving_with_threads(tmp_path, monkeypatch):
    """--nr_inference_workers / --nr_post_proc_workers size loader / writer thread pools; every
    file is processed once, in order, on the calling thread; a failed write is not swallowed."""
    import threading

    import cv2
    from cerberus_b200.infer import tile as tile_mod

    in_dir, out_dir = tmp_path / "in", tmp_path / "out"
    in_dir.mkdir()
    for i in range(7):
        cv2.imwrite(str(in_dir / ("img%d.png" % i)), np.full((8, 9, 3), 10 * i, np.uint8))
    m = object.__new__(tile_mod.InferManager)
    main = threading.get_ident()
    seen, saved = [], []

    groups, fin_threads = [], set()

    def fake_forward(named_images, slot=0):
        assert threading.get_ident() == main      # the forward never leaves the calling thread
        groups.append(len(named_images))
        return list(named_images)

    def fake_forward_wrapped(named_images, slot=0):
        return {"metas": fake_forward(named_images, slot)}

    def fake_finish(meta, group, to_host=True):
        fin_threads.add(threading.get_ident())    # the finishing stage: worker threads
        name, img = meta
        assert img.shape == (8, 9, 3)
        seen.append((name, int(img[0, 0, 0])))
        return (name, img, {}, {}, {}, None)

    def fake_save(results, root):
        if results[0] == "img5" and fail["on"]:
            raise RuntimeError("disk full")
        saved.append((results[0], threading.get_ident() != main))

    fail = {"on": False}
    monkeypatch.setattr(m, "_forward_group", fake_forward_wrapped, raising=False)
    monkeypatch.setattr(m, "_finish_one", fake_finish, raising=False)
    monkeypatch.setattr(m, "_sync_forward", lambda: None, raising=False)
    monkeypatch.setattr(tile_mod.InferManager, "_save", staticmethod(fake_save))
    for workers in (0, 2):
        seen.clear()
        saved.clear()
        groups.clear()
        m.process_file_list({"input_dir": str(in_dir), "output_dir": str(out_dir), "postproc_list": ["gland"],
                             "nr_inference_workers": workers, "nr_post_proc_workers": workers,
                             "patch_input_shape": 16, "patch_output_shape": 4, "patch_output_overlap": 0})
        # 8x9 px at 16/4: 2 x 3 grid, duplicated = 12 entries per file; files are cached until more
        # than 256 entries are pending (infer/tile.py:322-323) -> all 7 files form one group
        assert groups == [7]
        assert 1 <= len(fin_threads) <= max(1, workers) and main not in fin_threads
        fin_threads.clear()
        assert sorted(seen) == [("img%d" % i, 10 * i) for i in range(7)]
        assert sorted(s[0] for s in saved) == ["img%d" % i for i in range(7)]
        assert all(s[1] == (workers > 0) for s in saved)
    fail["on"] = True
    with pytest.raises(RuntimeError):
        m.process_file_list({"input_dir": str(in_dir), "output_dir": str(out_dir), "postproc_list": ["gland"],
                             "nr_inference_workers": 2, "nr_post_proc_workers": 2,
                             "patch_input_shape": 16, "patch_output_shape": 4, "patch_output_overlap": 0})


def test_process_file_list_groups_and_shards_files(tmp_path, monkeypatch):
    """Cache groups close when more than 256 patch-grid entries are pending (infer/tile.py:322-323);
    with world_size 2 the sorted file list is split rank-strided at file granularity."""
    import cv2
    from cerberus_b200.infer import tile as tile_mod
    in_dir, out_dir = tmp_path / "in", tmp_path / "out"
    in_dir.mkdir()
    for i in range(9):  # 40x40 px at 16/4: 10 x 10 grid, duplicated = 200 entries per file
        cv2.imwrite(str(in_dir / ("f%d.png" % i)), np.full((40, 40, 3), i, np.uint8))
    monkeypatch.setattr(tile_mod.InferManager, "_save", staticmethod(lambda results, root: None))
    for world in (1, 2):
        names = []
        for rank in range(world):
            m = object.__new__(tile_mod.InferManager)
            m.rank, m.world_size = rank, world
            groups = []
            monkeypatch.setattr(m, "_forward_group", lambda ni, slot=0, g=groups: (
                g.append([n for n, _ in ni]) or {"metas": list(ni)}), raising=False)
            monkeypatch.setattr(m, "_finish_one", lambda meta, grp, to_host=True: (
                meta[0], meta[1], {}, {}, {}, None), raising=False)
            monkeypatch.setattr(m, "_sync_forward", lambda: None, raising=False)
            m.process_file_list({"input_dir": str(in_dir), "output_dir": str(out_dir), "postproc_list": ["gland"],
                                 "nr_inference_workers": 0, "nr_post_proc_workers": 0,
                                 "patch_input_shape": 16, "patch_output_shape": 4, "patch_output_overlap": 0})
            assert all(len(g) == 2 for g in groups[:-1]) and 1 <= len(groups[-1]) <= 2
            names.append([n for g in groups for n in g])
        if world == 1:
            assert names[0] == ["f%d" % i for i in range(9)]
        else:
            assert names[0] == ["f0", "f2", "f4", "f6", "f8"] and names[1] == ["f1", "f3", "f5", "f7"]
