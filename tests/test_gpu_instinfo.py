"""GPU: the device instance tables (cerb_inst_info) against
  * tests/golden/instinfo.npz — output of the UNMODIFIED reference get_inst_info_dict
    (loader/postproc.py:12-98): ids, boxes, float64 centroids, contours, type, type_prob, all exact;
  * the OpenCV-based oracle on larger seeded maps (nuclei labels of the device watershed on a
    700x900 field, gland labels, x2 upsampling, device-resident inputs);
  * the tiatoolbox-flavoured table of the WSI path (get_instance_info)."""
import ctypes
import os

import cv2
import numpy as np
import pytest
from scipy import ndimage

from cerberus_b200 import _lib, instinfo, synth
from cerberus_b200.engine import Context
from cerberus_b200.postproc import post_process_batch
from oracle import instinfo_oracle as oi
from oracle.gen_golden import instinfo_cases
from tests.test_contour_host import GOLD, check_against_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built_lib):
    c = Context(0, "f16")
    yield c
    c.close()


def _same(a, b, type_keys=("type", "type_prob")):
    assert list(a.keys()) == list(b.keys())
    assert [type(k) for k in a.keys()] == [type(k) for k in b.keys()]
    for k in a:
        assert np.array_equal(a[k]["box"], b[k]["box"]), k
        assert np.array_equal(a[k]["centroid"], b[k]["centroid"]), k
        assert a[k]["centroid"].dtype == b[k]["centroid"].dtype
        assert np.array_equal(a[k]["contour"], b[k]["contour"]), k
        for t in type_keys:
            assert a[k].get(t) == b[k].get(t), (k, t)


def test_inst_info_matches_reference_golden(ctx):
    g = np.load(os.path.join(GOLD, "instinfo.npz"))
    for name, inst, typ, ds, up in instinfo_cases():
        info = instinfo.get_inst_info_dict(inst, typ, ds, ctx=ctx, up=up)
        check_against_golden(g, name, info, typ is not None)


def test_inst_info_large_maps_vs_oracle(ctx):
    # nuclei: labels of the device watershed on a 700x900 field (thousands of instances)
    f = synth.postproc_field(700, 900, "Nuclei", seed=3)
    canvas = np.zeros((1, 700, 900, 2), np.float32)
    canvas[0] = f
    lab, _ = post_process_batch(ctx, canvas, 0, "Nuclei", 1.0)
    lab = lab[0]
    typ = (ndimage.gaussian_filter(np.random.RandomState(1).rand(700, 900), 15) * 40 % 7).astype(
        np.int32).astype(np.float32)
    assert lab.max() > 300
    _same(instinfo.get_inst_info_dict(lab, typ, ctx=ctx), oi.get_inst_info_dict(lab, typ))
    _same(instinfo.get_inst_info_dict(lab, None, ctx=ctx), oi.get_inst_info_dict(lab, None))
    # tile mode: x2 nearest copies on the reference side, `up=2` here
    sub, tsub = lab[:300, :340], typ[:300, :340]
    up = lambda a: cv2.resize(a, (0, 0), fx=2, fy=2, interpolation=cv2.INTER_NEAREST)  # noqa: E731
    _same(instinfo.get_inst_info_dict(sub, tsub, ctx=ctx, up=2),
          oi.get_inst_info_dict(up(sub), up(tsub)))
    # gland: float64 labels, wide boxes (many 32-pixel chunks per row), ds_factor 0.5
    gf = synth.postproc_field(600, 800, "Gland", seed=5)
    canvas = np.zeros((1, 600, 800, 2), np.float32)
    canvas[0] = gf
    gl, _ = post_process_batch(ctx, canvas, 0, "Gland", 1.0)
    gl = gl[0].astype(np.float64)
    assert gl.max() >= 3
    gt = (np.arange(800)[None, :] // 100 % 3 + np.zeros((600, 1))).astype(np.float32)
    _same(instinfo.get_inst_info_dict(gl, gt, 0.5, ctx=ctx), oi.get_inst_info_dict(gl, gt, 0.5))
    # WSI flavour (tiatoolbox get_instance_info): flat boxes, `prob`
    a = instinfo.get_instance_info(lab, typ, ctx=ctx)
    b = oi.get_instance_info(lab, typ)
    _same(a, b, type_keys=("type", "prob"))


def test_inst_info_device_inputs_and_errors(ctx):
    rng = np.random.RandomState(2)
    lab = ndimage.label(ndimage.gaussian_filter(rng.randn(200, 333), 2) > 0.3)[0].astype(np.int32)
    typ = rng.randint(0, 4, lab.shape).astype(np.float32)
    lib = ctx.lib
    d_lab = lib.cerb_dev_alloc(ctx.handle, lab.nbytes)
    d_typ = lib.cerb_dev_alloc(ctx.handle, typ.nbytes)
    _lib.check(lib.cerb_memcpy(ctx.handle, d_lab, lab.ctypes.data_as(ctypes.c_void_p), lab.nbytes, 1), "memcpy")
    _lib.check(lib.cerb_memcpy(ctx.handle, d_typ, typ.ctypes.data_as(ctypes.c_void_p), typ.nbytes, 1), "memcpy")
    t_dev = instinfo.inst_table(ctx, d_lab, d_typ, on_device=True, shape=lab.shape)
    t_host = instinfo.inst_table(ctx, lab, typ)
    for f in ("ids", "box", "moments", "type", "contour_off", "contour_xy"):
        assert np.array_equal(getattr(t_dev, f), getattr(t_host, f)), f
    assert len(t_host.ids) == lab.max()
    lib.cerb_dev_free(ctx.handle, d_lab)
    lib.cerb_dev_free(ctx.handle, d_typ)
    # a table after a bigger one (workspace reuse), an empty map, loud failures
    small = np.zeros((5, 7), np.int32)
    assert instinfo.get_inst_info_dict(small, None, ctx=ctx) == {}
    small[1:4, 2:6] = 9
    info = instinfo.get_inst_info_dict(small, None, ctx=ctx)
    assert list(info.keys()) == [9] and info[9]["contour"].tolist() == [[2, 1], [2, 3], [5, 3], [5, 1]]
    with pytest.raises(RuntimeError):
        instinfo.get_inst_info_dict(small, np.full(small.shape, 0.3, np.float32), ctx=ctx)
    with pytest.raises(RuntimeError):
        instinfo.get_inst_info_dict(-small, None, ctx=ctx)
    with pytest.raises(ValueError):
        instinfo.get_inst_info_dict(small + 0.5, None, ctx=ctx)
