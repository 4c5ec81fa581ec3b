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
