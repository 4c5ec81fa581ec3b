"""CPU tests of the instance-table path, no GPU needed:
  * cerberus_b200/csrc/contour_core.h — the border-following code the kernels run — compiled for
    the host (tests/native/contour_host.cpp) and diffed against the OpenCV of this image on
    seeded masks (noise with nested holes / islands, blobs, label mosaics, x2 upsampling), both
    the serial scan and the 32-pixel chunked scan the warps execute;
  * oracle/instinfo_oracle.py (the checker of the GPU tests) against tests/golden/instinfo.npz,
    which the UNMODIFIED reference get_inst_info_dict produced (oracle/gen_golden.py instinfo).
"""
import ctypes
import os
import subprocess

import cv2
import numpy as np
import pytest
from scipy import ndimage

from oracle.gen_golden import instinfo_cases
from oracle.instinfo_oracle import get_inst_info_dict

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def contour_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("native") / "contour_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so,
                    os.path.join(HERE, "native", "contour_host.cpp")], check=True)
    return ctypes.CDLL(so)


def _ours(lib, lab, inst_id, up, chunked):
    lab = np.ascontiguousarray(lab, dtype=np.int32)
    out = np.zeros((lab.size * up * up * 4 + 16, 2), np.int32)
    box = np.zeros(4, np.int32)
    n = lib.contour0_host(lab.ctypes.data_as(ctypes.c_void_p), lab.shape[0], lab.shape[1],
                          int(inst_id), up, out.ctypes.data_as(ctypes.c_void_p), out.shape[0],
                          box.ctypes.data_as(ctypes.c_void_p), chunked)
    assert n >= 0, n
    return out[:n], box


def _opencv(lab, inst_id, up):
    """loader/postproc.py:19-31 on one instance."""
    if up != 1:
        lab = cv2.resize(lab.astype(np.int32), (0, 0), fx=up, fy=up, interpolation=cv2.INTER_NEAREST)
    m = lab == inst_id
    r0, r1 = np.where(np.any(m, 1))[0][[0, -1]]
    c0, c1 = np.where(np.any(m, 0))[0][[0, -1]]
    crop = m[r0:r1 + 1, c0:c1 + 1].astype(np.uint8)
    c = cv2.findContours(crop, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    return c[0][0].reshape(-1, 2).astype(np.int32), np.array([r0, c0, r1 + 1, c1 + 1])


def _mask(rng, it):
    H, W = rng.randint(3, 40), rng.randint(3, 90)
    kind = it % 4
    if kind == 0:
        return (rng.rand(H, W) < rng.uniform(0.2, 0.9)).astype(np.int32)
    if kind == 1:
        f = ndimage.gaussian_filter(rng.randn(H, W), rng.uniform(0.5, 3))
        return (f > np.percentile(f, rng.uniform(20, 80))).astype(np.int32)
    if kind == 2:
        f = ndimage.gaussian_filter(rng.randn(H, W), 1.5)
        return (f > 0).astype(np.int32) * rng.randint(1, 4, (H, W))
    lab = np.zeros((H, W), np.int32)
    for _ in range(rng.randint(1, 4)):
        y0, x0 = rng.randint(0, H - 1), rng.randint(0, W - 1)
        y1, x1 = rng.randint(y0 + 1, H + 1), rng.randint(x0 + 1, W + 1)
        lab[y0:y1, x0:x1] = 1 - lab[y0:y1, x0:x1] if rng.rand() < 0.5 else 1
    return lab


def test_border_following_matches_opencv(contour_lib):
    rng = np.random.RandomState(0)
    cases = 0
    for it in range(1200):
        lab = _mask(rng, it)
        for inst_id in np.unique(lab[lab > 0]):
            for up in (1, 2):
                a, box_a = _ours(contour_lib, lab, inst_id, up, it & 1)
                b, box_b = _opencv(lab, inst_id, up)
                assert np.array_equal(box_a, box_b), (it, inst_id, up)
                assert a.shape == b.shape and np.array_equal(a, b), (it, inst_id, up)
                cases += 1
    assert cases > 3000


def test_chunked_scan_on_wide_instances(contour_lib):
    """Boxes wider than several 32-pixel chunks, holes and islands straddling chunk borders."""
    rng = np.random.RandomState(5)
    for it in range(40):
        f = ndimage.gaussian_filter(rng.randn(70, 300), rng.uniform(1.0, 4.0))
        lab = (f > np.percentile(f, rng.uniform(30, 70))).astype(np.int32)
        a, _ = _ours(contour_lib, lab, 1, 1, 1)
        s, _ = _ours(contour_lib, lab, 1, 1, 0)
        b, _ = _opencv(lab, 1, 1)
        assert np.array_equal(a, b) and np.array_equal(s, b), it


def test_instinfo_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "instinfo.npz"))
    for name, inst, typ, ds, up in instinfo_cases():
        assert np.array_equal(g[name + "/inst_sum"],
                              [float(np.asarray(inst, np.float64).sum()), inst.shape[0], inst.shape[1]]), name
        a, t = inst, typ
        if up != 1:
            a = cv2.resize(inst, (0, 0), fx=up, fy=up, interpolation=cv2.INTER_NEAREST)
            t = None if typ is None else cv2.resize(typ, (0, 0), fx=up, fy=up,
                                                    interpolation=cv2.INTER_NEAREST)
        info = get_inst_info_dict(a, t, ds)
        check_against_golden(g, name, info, typ is not None)


def check_against_golden(g, name, info, has_type):
    keys = list(info.keys())
    assert [float(k) for k in keys] == list(g[name + "/ids"]), name
    box = np.array([info[k]["box"] for k in keys]).reshape(-1, 2, 2)
    assert np.array_equal(box, g[name + "/box"]), name
    cen = np.array([info[k]["centroid"] for k in keys], dtype=np.float64).reshape(-1, 2)
    assert np.array_equal(cen, g[name + "/centroid"]), name  # bit-exact float64
    off = g[name + "/contour_off"]
    for i, k in enumerate(keys):
        assert np.array_equal(np.asarray(info[k]["contour"]).reshape(-1, 2),
                              g[name + "/contour"][off[i]:off[i + 1]]), (name, k)
    if has_type:
        assert [info[k]["type"] for k in keys] == list(g[name + "/type"]), name
        assert np.array_equal(np.array([info[k]["type_prob"] for k in keys]), g[name + "/type_prob"]), name
