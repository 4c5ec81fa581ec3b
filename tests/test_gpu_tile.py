"""GPU: tile plumbing through the C ABI — device reflect-pad patch extraction, stitching and
the whole `_post_process_patches` / `process_image` path — against reference goldens."""
import hashlib
import os

import numpy as np
import pytest
import torch

from cerberus_b200 import synth
from cerberus_b200.infer.tile import InferManager, _post_process_patches, _prepare_patching
from cerberus_b200.postproc import PostProcInstErodedContourMap
from oracle import net_oracle, postproc_oracle as po
from oracle.gen_golden import tile_case_inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def manager(built_lib, tmp_path_factory):
    d = tmp_path_factory.mktemp("model")
    synth.write_model_dir(str(d), seed=0)
    import yaml
    st = yaml.full_load(open(os.path.join(str(d), "settings.yml")))
    m = InferManager(checkpoint_path=os.path.join(str(d), "weights.tar"),
                     decoder_dict=st["dataset_kwargs"]["req_target_code"],
                     model_args=st["model_kwargs"], precision="f16")
    PostProcInstErodedContourMap.bind(m.engine.ctx)
    yield m
    m.engine.close()


def test_extract_patches_matches_reference_golden(manager):
    g = np.load(os.path.join(GOLD, "patching.npz"))
    for (h, w, i, o) in g["cases"]:
        key = "%dx%d_%d_%d" % (h, w, i, o)
        img = np.random.RandomState(h * 1000 + w).randint(0, 256, (h, w, 3)).astype(np.uint8)
        info = g[key + "/info"]
        patches = manager.extract_patches(img, info[:, 0, 0, :], int(i), g[key + "/src_pos"])
        for k, sha in zip(g[key + "/sel"], g[key + "/patch_sha1"]):
            assert hashlib.sha1(patches[k].tobytes()).hexdigest() == str(sha), (key, int(k))
        padded, _, _ = _prepare_patching(img, int(i), int(o), 0)
        for k in range(info.shape[0]):
            (y0, x0), (y1, x1) = info[k, 0]
            assert np.array_equal(patches[k], padded[y0:y1, x0:x1])


def test_post_process_patches_matches_reference_golden(manager):
    g = np.load(os.path.join(GOLD, "tile_postproc.npz"))
    margs = synth.model_args()
    codes = dict(synth.DEFAULT_REQ_TARGET_CODE)
    for seed in (0, 1):
        plist, image_info = tile_case_inputs(seed)
        name, _, inst, info, types, pclass = _post_process_patches(
            plist, image_info, codes, ["gland", "lumen", "nuclei", "patch-class"], margs)
        for t, m in inst.items():
            assert str(m.dtype) == str(g["s%d/inst_dtype/%s" % (seed, t)]), t
            assert np.array_equal(m.astype(np.int64), g["s%d/inst/%s" % (seed, t)].astype(np.int64)), t
            ids = sorted(int(k) for k in info[t].keys())
            assert ids == list(g["s%d/ids/%s" % (seed, t)]), t
            assert [info[t][k].get("type", -1) for k in sorted(info[t].keys())] == list(g["s%d/types/%s" % (seed, t)])
            cen = np.array([info[t][k]["centroid"] for k in sorted(info[t].keys())]).reshape(-1, 2)
            assert np.allclose(cen, g["s%d/centroids/%s" % (seed, t)])
            if types[t] is not None:
                assert np.array_equal(types[t].astype(np.uint8), g["s%d/type_map/%s" % (seed, t)])
        assert np.array_equal(pclass.astype(np.uint8), g["s%d/pclass" % seed])


def test_process_image_end_to_end_vs_oracle(manager, six_head_sd):
    """A 300x380 image through the whole drop-in path (448/144 patches, duplicate grid
    inferred once) vs the oracle pipeline fed with the oracle's own forward. Float maps agree to
    fp16-mode tolerance; label maps are compared on the device path's own canvas (bit-exact)."""
    margs = synth.model_args()
    img = synth.synthetic_tiles(1, 304, 384, seed=3)[0][:300, :380]
    manager.patch_input_shape, manager.patch_output_shape = 448, 144
    manager.patch_output_overlap, manager.batch_size = 0, 5
    manager.postproc_list = ["gland", "lumen", "nuclei", "patch-class"]
    name, _, inst, info, types, pclass = manager.process_image(img, "t")
    assert set(inst.keys()) == {"Gland", "Lumen", "Nuclei"}
    for t in inst:
        assert inst[t].shape == (300, 380)
    assert inst["Gland"].dtype == np.float64 and inst["Lumen"].dtype == np.float64
    assert pclass.shape == (300, 380)
    assert types["Nuclei"].shape == (300, 380) and types["Lumen"] is None
    # reference plumbing on CPU for the same image
    padded, pinfo, src_pos = _prepare_patching(img, 448, 144, 0)
    half = pinfo.shape[0] // 2
    batch = np.stack([padded[a[0][0]:a[1][0], a[0][1]:a[1][1]] for a in pinfo[:half, 0]])
    step, _ = net_oracle.infer_step(six_head_sd, batch, 144, margs["decoder_kwargs"], margs["considered_tasks"])
    ref_pclass = np.zeros((300, 380), np.float32)
    for k in range(half):
        (y0, x0), (y1, x1) = pinfo[k, 1]
        yy0, xx0 = y0 - src_pos[0], x0 - src_pos[1]
        sy0, sx0 = max(yy0, 0), max(xx0, 0)
        sy1, sx1 = min(yy0 + 144, 300), min(xx0 + 144, 380)
        if sy1 > sy0 and sx1 > sx0:
            ref_pclass[sy0:sy1, sx0:sx1] = step[k]["Patch-Class"][sy0 - yy0:sy1 - yy0, sx0 - xx0:sx1 - xx0]
    assert (ref_pclass != pclass).mean() < 0.02


def _read_dev(manager, ptr, shape, dtype):
    import ctypes
    out = np.empty(shape, dtype=dtype)
    ctx = manager.engine.ctx
    from cerberus_b200 import _lib
    _lib.check(ctx.lib.cerb_memcpy(ctx.handle, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(ptr),
                                   out.nbytes, 2), "D2H")
    return out


def _same_results(a, b):
    _, _, inst_a, info_a, types_a, pc_a = a
    _, _, inst_b, info_b, types_b, pc_b = b
    assert inst_a.keys() == inst_b.keys()
    for t in inst_a:
        assert inst_a[t].dtype == inst_b[t].dtype, t
        assert np.array_equal(inst_a[t], inst_b[t]), t
        assert list(info_a[t].keys()) == list(info_b[t].keys()), t
        assert [type(k) for k in info_a[t]] == [type(k) for k in info_b[t]], t
        for k in info_a[t]:
            for f in ("box", "centroid", "contour"):
                assert np.array_equal(info_a[t][k][f], info_b[t][k][f]), (t, k, f)
            assert info_a[t][k].get("type") == info_b[t][k].get("type"), (t, k)
            assert info_a[t][k].get("type_prob") == info_b[t][k].get("type_prob"), (t, k)
        if types_a[t] is None:
            assert types_b[t] is None
        else:
            assert np.array_equal(types_a[t], types_b[t]), t
    assert np.array_equal(pc_a, pc_b)


def test_device_tile_path_equals_host_plumbing_and_oracle(manager):
    """process_image (device-resident: extract -> plan -> stitch -> post-proc -> instance tables
    without a host round trip) must equal (a) the reference-shaped host plumbing (run_step list of
    dicts -> _post_process_patches) in every output, and (b) the ORACLE post-processing
    (loader/postproc.py restated) applied to the device path's own stitched canvas, bit-exact."""
    from oracle import postproc_oracle as po
    img = synth.synthetic_tiles(1, 304, 384, seed=13)[0][:300, :380]
    manager.patch_input_shape, manager.patch_output_shape = 448, 144
    manager.patch_output_overlap, manager.batch_size = 0, 4
    manager.postproc_list = ["gland", "lumen", "nuclei", "patch-class"]
    dev = manager.process_image(img, "t")
    d_canvas, H, W, C = manager.last_canvas_dev
    canvas = _read_dev(manager, d_canvas, (H, W, C), np.float32)
    host = manager.process_image_host_plumbing(img, "t")
    _same_results(dev, host)
    idx = manager.engine.model.idx_dict
    ref = {t: po.post_process(canvas, idx, t, 1.0)[0] for t in ("Nuclei", "Gland", "Lumen")}
    ref["Lumen"] = ref["Lumen"] * (ref["Gland"] > 0)
    n_inst = 0
    for t in ref:
        assert np.array_equal(dev[2][t].astype(np.int64), ref[t].astype(np.int64)), t
        n_inst += int(ref[t].max())
    assert n_inst > 20  # the check is not vacuous


def test_batches_are_filled_across_images(manager):
    """infer/tile.py:294-325: patches of several files share batches. Three images of different
    sizes through ONE process_images call with batch 7 == each image alone with batch 3."""
    imgs = [synth.synthetic_tiles(1, 256, 256, seed=31)[0], synth.synthetic_tiles(1, 304, 384, seed=32)[0][:290, :333],
            synth.synthetic_tiles(1, 160, 208, seed=33)[0][:150, :200]]
    manager.patch_input_shape, manager.patch_output_shape = 448, 144
    manager.patch_output_overlap = 0
    manager.postproc_list = ["gland", "lumen", "nuclei", "patch-class"]
    manager.batch_size = 7
    manager.nr_patches_inferred = 0
    together = manager.process_images([("a%d" % i, im) for i, im in enumerate(imgs)])
    assert manager.nr_patches_inferred == 4 + 9 + 4  # unique patches: ceil(size / 144)^2 per image
    manager.batch_size = 3
    for i, im in enumerate(imgs):
        _same_results(together[i], manager.process_image(im, "a%d" % i))
