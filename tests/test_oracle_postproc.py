"""CPU: pins the post-processing oracle (oracle/postproc_oracle.c).
  * against tests/golden/postproc.npz = outputs of the UNMODIFIED reference
    loader/postproc.py code (oracle/gen_golden.py);
  * against scipy.ndimage / OpenCV directly for the primitives those libraries own;
  * the C watershed against the pure-Python transcription of the skimage 0.19 heap algorithm.
scikit-image itself is not installed: the heap order is restated, not diffed."""
import os

import numpy as np
import pytest

from oracle import postproc_oracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "postproc.npz")
IDX = {"Lumen-INST": [0, 2], "Gland-INST": [2, 4], "Nuclei-INST": [4, 6]}


def golden_cases():
    g = np.load(GOLD)
    for key in g["names"]:
        key = str(key)
        tissue = key.split("/")[1]
        field = (g[key + "/field_q12"].astype(np.float32) / 4096.0).astype(np.float32)
        yield key, tissue, float(g[key + "/ds"]), field, g[key + "/inst"].astype(np.int64), str(g[key + "/dtype"])


def test_oracle_matches_reference_postproc_golden():
    n = 0
    for key, tissue, ds, field, inst, dtype in golden_cases():
        raw = np.zeros(field.shape[:2] + (6,), np.float32)
        lo = IDX[tissue + "-INST"][0]
        raw[..., lo:lo + 2] = field
        got, type_map = po.post_process(raw, IDX, tissue, ds)
        assert str(got.dtype) == dtype, key
        assert np.array_equal(got.astype(np.int64), inst), key
        assert type_map is None
        n += 1
    assert n >= 50


def test_primitives_match_scipy_and_opencv():
    cv2 = pytest.importorskip("cv2")
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.RandomState(0)
    for i in range(8):
        h, w = rng.randint(5, 120, 2)
        fg = rng.rand(h, w) > rng.uniform(0.3, 0.7)
        l1, n1 = ndi.label(fg)
        l2, n2 = po.label4(fg)
        assert n1 == n2 and np.array_equal(l1, l2)
        assert np.array_equal(ndi.binary_fill_holes(fg), po.fill_holes(fg))
        for k in (1, 2, 3, 4, 5, 7, 10, 11, 15):
            e = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
            assert np.array_equal(e, po.ellipse(k)), k
            assert np.array_equal(cv2.dilate(fg.astype(np.uint8), e), po.dilate(fg, e)), k
        assert np.array_equal(cv2.erode(fg.astype(np.uint8), po.ellipse(3)), po.erode_cross(fg))


def test_remove_small_objects_semantics():
    a = np.zeros((8, 8), bool)
    a[0, 0:3] = True          # size 3
    a[4:6, 4:6] = True        # size 4
    out = po.remove_small_objects(a, min_size=4)
    assert out.dtype == bool and not out[0, 0] and out[4, 4]   # strict "<"
    lab = np.zeros((8, 8), np.int32)
    lab[0, 0:3] = 5
    lab[4:6, 4:6] = 2
    out = po.remove_small_objects(lab, min_size=4)
    assert out[0, 0] == 0 and out[4, 4] == 2                   # labels are kept, not renumbered


def test_watershed_c_matches_python_transcription():
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 10 ** 6))
    def run(seed):
        rng = np.random.RandomState(seed)
        h, w = rng.randint(2, 24, 2)
        levels = rng.choice([2, 3, 5, 1000])
        img = -np.round(rng.rand(h, w) * levels) / levels      # many exact ties
        mk = (rng.rand(h, w) > 0.8) * rng.randint(1, 6, (h, w))
        mask = rng.rand(h, w) > 0.15
        assert np.array_equal(po.watershed_spec(img, mk, mask), po.watershed(img, mk, mask=mask))

    run()


GOLD_ERODED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "postproc_eroded.npz")


def eroded_golden_cases():
    """(key, tissue, single-channel field [H,W,1], reference label map, dtype): outputs of the
    UNMODIFIED PostProcInstErodedMap.post_process (oracle/gen_golden.py postproc_eroded) on the
    inner channel of the ds = 1 fields of postproc.npz."""
    base = np.load(GOLD)
    g = np.load(GOLD_ERODED)
    for key in g["names"]:
        key = str(key)
        name, tissue0, tissue = key.split("/")
        field = (base["%s/%s/field_q12" % (name, tissue0)].astype(np.float32) / 4096.0).astype(np.float32)
        yield key, tissue, np.ascontiguousarray(field[..., :1]), g[key + "/inst"].astype(np.int64), str(g[key + "/dtype"])


def test_eroded_map_oracle_matches_reference_golden():
    n = with_instances = 0
    for key, tissue, field, inst, dtype in eroded_golden_cases():
        got, type_map = po.post_process_eroded(field, {tissue + "-INST": [0, 1]}, tissue)
        assert str(got.dtype) == dtype and type_map is None, key
        assert np.array_equal(got.astype(np.int64), inst), key
        n += 1
        with_instances += inst.max() > 0
    assert n >= 70 and with_instances >= 40
