"""Mirror of loader/postproc.py:12-98 (get_inst_info_dict) and of tiatoolbox's
HoVerNet.get_instance_info (infer/wsi.py:150) over the C ABI.

The reference loops over instances in Python and calls cv2.moments / cv2.findContours on each
bounding-box crop; here `cerb_inst_info` builds the whole table on the device (box, moments,
majority type, contour [0][0] — csrc/instinfo.cu, csrc/contour_core.h) and this module only
reshapes it into the reference's dict. Same keys, dtypes, skip rules and ds_factor rounding.
No CPU fallback: without a bound device context the call raises.
"""
import ctypes

import numpy as np

from . import _lib

_ctx = None


def bind(ctx):
    """Device context used when a call does not pass one (set by InferManager)."""
    global _ctx
    _ctx = ctx


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class InstTable:
    """Host copy of the device table, rows in ascending instance id."""
    __slots__ = ("ids", "box", "moments", "type", "contour_off", "contour_xy", "any_background")


def inst_table(ctx, inst_map, type_map=None, up=1, on_device=False, shape=None):
    """Runs cerb_inst_info. `inst_map` / `type_map` are host arrays ([H,W]; any integer-valued
    dtype) or, with on_device=True, device pointers (int) of int32 / float32 [H,W] = `shape`."""
    ctx = ctx or _ctx
    if ctx is None:
        raise RuntimeError("instinfo.bind(ctx) has not been called: the instance tables are built "
                           "on the CUDA device only")
    lib = ctx.lib
    if on_device:
        H, W = shape
        lab_p = ctypes.c_void_p(inst_map)
        typ_p = ctypes.c_void_p(type_map) if type_map else None
        flags = 1
    else:
        inst_map = np.asarray(inst_map)
        lab = np.ascontiguousarray(inst_map, dtype=np.int32)
        if lab.shape != inst_map.shape or (inst_map.dtype.kind == "f" and (lab != inst_map).any()):
            raise ValueError("instance map holds non-integer ids")
        H, W = lab.shape
        lab_p = _ptr(lab)
        typ = None
        if type_map is not None:
            typ = np.ascontiguousarray(type_map, dtype=np.float32)
            if typ.shape != lab.shape:
                raise ValueError("type map shape %r != instance map shape %r" % (typ.shape, lab.shape))
        typ_p = _ptr(typ) if typ is not None else None
        flags = 0
    n = ctypes.c_int32(0)
    npts = ctypes.c_int64(0)
    any_bg = ctypes.c_int32(0)
    _lib.check(lib.cerb_inst_info(ctx.handle, lab_p, H, W, typ_p, int(up), flags, ctypes.byref(n),
                                  ctypes.byref(npts), ctypes.byref(any_bg)), "cerb_inst_info")
    t = InstTable()
    t.ids = np.empty(n.value, dtype=np.int32)
    t.box = np.empty((n.value, 4), dtype=np.int32)
    t.moments = np.empty((n.value, 3), dtype=np.int64)
    t.type = np.empty((n.value, 2), dtype=np.int32)
    t.contour_off = np.zeros(n.value + 1, dtype=np.int64)
    t.contour_xy = np.empty((npts.value, 2), dtype=np.int32)
    t.any_background = bool(any_bg.value)
    _lib.check(lib.cerb_inst_info_read(ctx.handle, _ptr(t.ids), _ptr(t.box), _ptr(t.moments),
                                       _ptr(t.type), _ptr(t.contour_off), _ptr(t.contour_xy)),
               "cerb_inst_info_read")
    return t


def _rows(table):
    """Instances in the order of `np.unique(inst_map)[1:]` that survive the reference's contour
    checks (postproc.py:31-37: fewer than 3 points -> skipped)."""
    first = 0 if table.any_background else 1  # [1:] drops the smallest id when there is no 0
    counts = np.diff(table.contour_off)
    return (np.nonzero(counts[first:] >= 3)[0] + first).tolist()


def tiatoolbox_dicts(table, rows, offset_xy=(0, 0), has_type=True):
    """Rows of a device table as tiatoolbox-style instance dicts (box flat [x0, y0, x1, y1], int64
    contours), shifted by `offset_xy` (the tile's top-left in infer/wsi.py:225-227). Every field is
    computed for all rows at once - half a million nuclei per slide go through here - with the
    reference's operation order ((m10 / m00 + box origin) + offset in float64)."""
    rows = np.asarray(rows, dtype=np.int64)
    off = np.asarray(offset_xy, dtype=np.int64)
    box = table.box[rows][:, [1, 0, 3, 2]].astype(np.int64)          # x0, y0, x1, y1
    m = table.moments[rows].astype(np.float64)
    cen = np.stack([m[:, 1] / m[:, 0], m[:, 2] / m[:, 0]], axis=1)
    cen = cen + box[:, :2]
    if off.any():
        cen = cen + off
        box = box + np.concatenate([off, off])
    xy = table.contour_xy.astype(np.int64) + off
    starts = table.contour_off[rows].tolist()
    stops = table.contour_off[rows + 1].tolist()
    if has_type:
        types = (table.type[rows, 0] / 4.0).astype(np.int64).tolist()  # int(np.float) truncates
        probs = (table.type[rows, 1] / (table.moments[rows, 0] + 1.0e-6)).tolist()
    else:
        types = probs = [None] * len(rows)
    return [{"box": box[j], "centroid": cen[j], "contour": xy[starts[j]:stops[j]], "prob": probs[j],
             "type": types[j]} for j in range(len(rows))]


def get_inst_info_dict(inst_map, type_map, ds_factor=1.0, ctx=None, up=1, key_dtype=None,
                       on_device=False, shape=None):
    """loader/postproc.py:12-98. `up` folds the cv2.resize(fx=up, fy=up, INTER_NEAREST) of
    infer/tile.py:196-201 into the call (pass the maps at processing resolution). With
    on_device=True the maps are device pointers of int32 / float32 [H,W] = `shape` (pass
    key_dtype: the reference's keys carry the dtype of its label map, e.g. float64 for glands)."""
    table = inst_table(ctx, inst_map, type_map, up=up, on_device=on_device, shape=shape)
    if key_dtype is None:
        key_dtype = np.int32 if on_device else np.asarray(inst_map).dtype.type  # np.unique keeps the dtype
        if up != 1 and key_dtype is np.int64:
            key_dtype = np.int32  # the cv2.resize this call folds in hands int64 maps back as int32
    has_type = bool(type_map) if on_device else type_map is not None
    # All fields for all rows at once (a 1000 x 1000 tile holds ~1500 nuclei; a per-instance
    # Python loop over numpy scalars costs 15 us each): same arithmetic and dtypes as the
    # reference's loop - bbox int64 [[rmin, cmin], [rmax, cmax]], centroid float64 (m10 / m00 +
    # cmin, m01 / m00 + rmin), contour = the device's int32 points, type int, type_prob float.
    rows = np.asarray(_rows(table), dtype=np.int64)
    info = {}
    if len(rows):
        box = table.box[rows].astype(np.int64).reshape(-1, 2, 2)
        m = table.moments[rows].astype(np.float64)
        cen = np.stack([m[:, 1] / m[:, 0] + box[:, 0, 1], m[:, 2] / m[:, 0] + box[:, 0, 0]], axis=1)
        starts, stops = table.contour_off[rows], table.contour_off[rows + 1]
        if np.array_equal(stops[:-1], starts[1:]):
            contours = np.split(table.contour_xy[starts[0]:stops[-1]], (stops[:-1] - starts[0]).tolist())
        else:
            contours = [table.contour_xy[a:b] for a, b in zip(starts.tolist(), stops.tolist())]
        contours = [c.copy() for c in contours]  # own their data, like the reference's arrays
        keys = [key_dtype(v) for v in table.ids[rows].tolist()]
        if has_type:
            types = (table.type[rows, 0] / 4.0).astype(np.int64).tolist()  # int(np.float) truncates (postproc.py:69)
            probs = (table.type[rows, 1] / (table.moments[rows, 0] + 1.0e-6)).tolist()
            for k, b, c, ct, ty, pr in zip(keys, list(box), list(cen), contours, types, probs):
                info[k] = {"box": b, "centroid": c, "contour": ct, "type": ty, "type_prob": pr}
        else:
            for k, b, c, ct in zip(keys, list(box), list(cen), contours):
                info[k] = {"box": b, "centroid": c, "contour": ct}
    if ds_factor != 1.0:
        for inst_id in list(info.keys()):
            d = info[inst_id]
            new = {"box": np.round(d["box"] / ds_factor).astype("int"),
                   "centroid": np.round(d["centroid"] / ds_factor).astype("int"),
                   "contour": np.round(d["contour"] / ds_factor).astype("int")}
            if "type" in d:
                new["type"], new["type_prob"] = d["type"], d["type_prob"]
            info[inst_id] = new
    return info


def get_instance_info(pred_inst, pred_type=None, ctx=None, on_device=False, shape=None):
    """tiatoolbox HoVerNet.get_instance_info (infer/wsi.py:150): box is flat [x0, y0, x1, y1],
    keys `type` / `prob` (None without a type map). With on_device=True `pred_inst` / `pred_type`
    are device pointers of int32 / float32 [H,W] = `shape`."""
    table = inst_table(ctx, pred_inst, pred_type, on_device=on_device, shape=shape)
    key_dtype = np.int32 if on_device else np.asarray(pred_inst).dtype.type
    rows = _rows(table)
    dicts = tiatoolbox_dicts(table, rows, has_type=pred_type is not None)
    return {key_dtype(table.ids[i]): d for i, d in zip(rows, dicts)}
