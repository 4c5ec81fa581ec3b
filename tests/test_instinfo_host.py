"""CPU: the Python half of the instance tables (cerberus_b200/instinfo.py — row selection, key and
array dtypes, centroid arithmetic, type decoding, ds_factor rounding, tile offsets) against the
reference golden and the OpenCV loop, with the device call replaced by a CPU twin of its output
format (tests/native/contour_host.cpp::inst_table_host, built on the same contour_core.h). The
kernels themselves are covered by tests/test_gpu_instinfo.py."""
import ctypes
import os

import cv2
import numpy as np
import pytest

from cerberus_b200 import instinfo
from oracle import instinfo_oracle as oi
from oracle.gen_golden import instinfo_cases
from tests.test_contour_host import GOLD, check_against_golden, contour_lib  # noqa: F401


@pytest.fixture()
def host_tables(contour_lib, monkeypatch):  # noqa: F811
    def fake_inst_table(ctx, inst_map, type_map=None, up=1, on_device=False, shape=None):
        assert not on_device
        lab = np.ascontiguousarray(inst_map, dtype=np.int32)
        H, W = lab.shape
        typ = None if type_map is None else np.ascontiguousarray(type_map, dtype=np.float32)
        n_max = int(lab.max()) + 1
        t = instinfo.InstTable()
        t.ids = np.zeros(n_max, np.int32)
        t.box = np.zeros((n_max, 4), np.int32)
        t.moments = np.zeros((n_max, 3), np.int64)
        t.type = np.zeros((n_max, 2), np.int32)
        t.contour_off = np.zeros(n_max + 1, np.int64)
        xy = np.zeros((lab.size * up * up * 2 + 16, 2), np.int32)
        bg = ctypes.c_int32(0)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        n = contour_lib.inst_table_host(p(lab), H, W, p(typ) if typ is not None else None, up,
                                        p(t.ids), p(t.box), p(t.moments), p(t.type),
                                        p(t.contour_off), p(xy), ctypes.c_int64(xy.shape[0]),
                                        ctypes.byref(bg))
        assert n >= 0
        t.ids, t.box, t.moments, t.type = t.ids[:n], t.box[:n], t.moments[:n], t.type[:n]
        t.contour_off = t.contour_off[:n + 1]
        t.contour_xy = xy[:int(t.contour_off[-1])]
        t.any_background = bool(bg.value)
        return t

    monkeypatch.setattr(instinfo, "inst_table", fake_inst_table)
    return fake_inst_table


def test_dict_assembly_matches_reference_golden(host_tables):
    g = np.load(os.path.join(GOLD, "instinfo.npz"))
    for name, inst, typ, ds, up in instinfo_cases():
        info = instinfo.get_inst_info_dict(inst, typ, ds, up=up)
        check_against_golden(g, name, info, typ is not None)
        ref = oi.get_inst_info_dict(
            inst if up == 1 else cv2.resize(inst, (0, 0), fx=up, fy=up, interpolation=cv2.INTER_NEAREST),
            typ if (typ is None or up == 1) else cv2.resize(typ, (0, 0), fx=up, fy=up,
                                                            interpolation=cv2.INTER_NEAREST), ds)
        assert [type(k) for k in info] == [type(k) for k in ref], name      # np.unique key dtype
        for k in ref:
            for f in ("box", "centroid", "contour"):
                assert info[k][f].dtype == ref[k][f].dtype, (name, f)


def test_tiatoolbox_flavour_and_tile_offsets(host_tables):
    for name, inst, typ, ds, up in instinfo_cases():
        if up != 1 or inst.dtype.kind == "f":
            continue
        a = instinfo.get_instance_info(inst, typ)
        b = oi.get_instance_info(inst, typ)
        assert list(a.keys()) == list(b.keys()), name
        for k in b:
            for f in ("box", "centroid", "contour"):
                assert np.array_equal(a[k][f], b[k][f]) and a[k][f].dtype == b[k][f].dtype, (name, k, f)
            assert a[k]["type"] == b[k]["type"] and a[k]["prob"] == b[k]["prob"], (name, k)
        # infer/wsi.py:225-227: box + [tl, tl], centroid + tl, contour + tl
        table = instinfo.inst_table(None, inst, typ)
        rows = instinfo._rows(table)
        tl = np.array([4032, 8064], dtype=np.int64)
        shifted = instinfo.tiatoolbox_dicts(table, rows, offset_xy=tl, has_type=typ is not None)
        for d, k in zip(shifted, b):
            assert np.array_equal(d["box"], b[k]["box"] + np.concatenate([tl] * 2))
            assert np.array_equal(d["centroid"], b[k]["centroid"] + tl)
            assert np.array_equal(d["contour"], b[k]["contour"] + tl)
