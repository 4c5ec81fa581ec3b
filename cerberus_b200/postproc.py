"""Host mirror of loader/postproc.py:147-407 (PostProcInstErodedContourMap, and the
PostProcInstErodedMap variant of the IP-ERODED-* codes) over the C ABI.

Same call signature and return dtypes as the reference; the arithmetic runs in
csrc/postproc.cu on the device. No CPU fallback.
"""
import copy
import ctypes

import numpy as np

from . import _lib

_TISSUE = {"GLAND": 0, "LUMEN": 1}


def _as_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def post_process_batch(ctx, canvas, ch0, tissue_mode, ds_factor=1.0):
    """canvas: float32 [n,H,W,C] host array. Returns (int32 [n,H,W] labels, any_fg [n] or None)."""
    lib = ctx.lib
    canvas = np.ascontiguousarray(canvas, dtype=np.float32)
    n, H, W, C = canvas.shape
    out = np.empty((n, H, W), dtype=np.int32)
    mode = tissue_mode.upper()
    if mode == "NUCLEI":
        any_fg = np.zeros(n, dtype=np.int32)
        _lib.check(lib.cerb_postproc_nuclei(ctx.handle, _as_ptr(canvas), n, H, W, C, ch0,
                                            _as_ptr(out), _as_ptr(any_fg), 0),
                   "cerb_postproc_nuclei")
        return out, any_fg
    if mode not in _TISSUE:
        raise AssertionError("unknown tissue mode %r" % tissue_mode)
    _lib.check(lib.cerb_postproc_gland_lumen(ctx.handle, _as_ptr(canvas), n, H, W, C, ch0,
                                             _TISSUE[mode], float(ds_factor), _as_ptr(out), 0),
               "cerb_postproc_gland_lumen")
    return out, None


class PostProcInstErodedContourMap:
    """Drop-in for the reference class: `post_process(raw_map, idx_dict, tissue_mode,
    ds_factor) -> (inst_map, type_map)`. Bind a device context once with `bind(ctx)`."""

    _ctx = None

    @classmethod
    def bind(cls, ctx):
        cls._ctx = ctx

    @classmethod
    def post_process(cls, raw_map, idx_dict, tissue_mode, ds_factor=1.0):
        if cls._ctx is None:
            raise RuntimeError("PostProcInstErodedContourMap.bind(ctx) has not been called: the "
                               "post-processing runs on the CUDA device only")
        assert tissue_mode.upper() in ("LUMEN", "GLAND", "NUCLEI")
        tissue_ch = "%s-INST" % tissue_mode
        idx_dict = copy.deepcopy(idx_dict)
        assert tissue_ch in list(idx_dict.keys())
        raw_map = np.asarray(raw_map)
        lo, hi = idx_dict[tissue_ch]
        if hi - lo < 2:
            raise ValueError("%s needs two channels (inner, contour)" % tissue_ch)
        labels, any_fg = post_process_batch(cls._ctx, raw_map[None], lo, tissue_mode, ds_factor)
        if tissue_mode.upper() == "NUCLEI":
            # int32 watershed output, or float64 zeros when the mask is empty (postproc.py:378-380)
            inst_map = labels[0] if any_fg[0] else np.zeros(labels[0].shape)
        else:
            inst_map = labels[0].astype(np.float64)  # postproc.py:290,331
        type_ch = tissue_mode + "-" + "TYPE"
        if type_ch in list(idx_dict.keys()):
            type_map = raw_map[..., idx_dict[type_ch][0]:idx_dict[type_ch][1]]
            type_map = np.squeeze(type_map)
        else:
            type_map = None
        return inst_map, type_map


class PostProcInstErodedMap:
    """Drop-in for loader/postproc.py:147-265 (target codes IP-ERODED-3 / IP-ERODED-11,
    infer/tile.py:35-37): `post_process(raw_map, idx_dict, tissue_mode, scale) ->
    (float64 inst_map, type_map)`; the type map is the raw channel slice (not squeezed), as in
    the reference. Shares the device context bound to PostProcInstErodedContourMap."""

    _TISSUE = {"GLAND": 0, "LUMEN": 1, "NUCLEI": 2}

    @classmethod
    def post_process(cls, raw_map, idx_dict, tissue_mode, scale=1.0):
        ctx = PostProcInstErodedContourMap._ctx
        if ctx is None:
            raise RuntimeError("PostProcInstErodedContourMap.bind(ctx) has not been called: the "
                               "post-processing runs on the CUDA device only")
        assert tissue_mode.upper() in cls._TISSUE
        tissue_ch = "%s-INST" % tissue_mode
        assert tissue_ch in list(idx_dict.keys())
        raw_map = np.asarray(raw_map)
        lo, hi = idx_dict[tissue_ch]
        if hi - lo != 1:
            raise ValueError("%s must be a single channel for the eroded-map post-processing "
                             "(the reference's np.squeeze + 2-D label map breaks otherwise)" % tissue_ch)
        canvas = np.ascontiguousarray(raw_map[None], dtype=np.float32)
        n, H, W, C = canvas.shape
        out = np.empty((n, H, W), dtype=np.int32)
        _lib.check(ctx.lib.cerb_postproc_eroded_map(ctx.handle, _as_ptr(canvas), n, H, W, C, lo,
                                                    cls._TISSUE[tissue_mode.upper()], _as_ptr(out), 0),
                   "cerb_postproc_eroded_map")
        inst_map = out[0].astype(np.float64)  # loader/postproc.py:157,187,217
        type_ch = tissue_mode + "-" + "TYPE"
        if type_ch in list(idx_dict.keys()):
            type_map = raw_map[..., idx_dict[type_ch][0]:idx_dict[type_ch][1]]
        else:
            type_map = None
        return inst_map, type_map
