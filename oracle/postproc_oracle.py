"""ORACLE — test infrastructure only (see oracle/postproc_oracle.c for the restated algorithms
and their reference citations). numpy/ctypes front-end:

  post_process(raw_map, idx_dict, tissue_mode, ds_factor)   <- loader/postproc.py:383-407
  remove_small_objects / watershed                           <- scikit-image 0.19.2 API subset
        (injected into the reference by oracle/ref_shim.py when generating goldens)
  watershed_spec                                             <- pure-Python transcription of the
        same heap algorithm used to cross-check the C code on small cases
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpostproc_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "postproc_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.run(["make", "-C", _HERE, "-s"], check=True)
        _lib = ctypes.CDLL(_SO)
        _lib.orc_label4.restype = ctypes.c_int
        _lib.orc_proc_nuclei.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def label4(fg):
    fg = np.ascontiguousarray(fg, dtype=np.uint8)
    out = np.empty(fg.shape, dtype=np.int32)
    n = lib().orc_label4(_p(fg), fg.shape[0], fg.shape[1], _p(out))
    return out, n


def fill_holes(fg):
    fg = np.ascontiguousarray(fg, dtype=np.uint8)
    out = np.empty(fg.shape, dtype=np.uint8)
    lib().orc_fill_holes(_p(fg), fg.shape[0], fg.shape[1], _p(out))
    return out.astype(bool)


def ellipse(k):
    e = np.zeros((k, k), dtype=np.uint8)
    if k > 0:
        lib().orc_ellipse(k, _p(e))
    return e


def erode_cross(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    out = np.empty_like(a)
    lib().orc_erode_cross(_p(a), a.shape[0], a.shape[1], _p(out))
    return out


def dilate(a, elem):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    elem = np.ascontiguousarray(elem, dtype=np.uint8)
    out = np.empty_like(a)
    lib().orc_dilate(_p(a), a.shape[0], a.shape[1], _p(elem), elem.shape[0], _p(out))
    return out


def remove_small_objects(ar, min_size=64, connectivity=1, in_place=False):
    """skimage.morphology.remove_small_objects (0.19): bool input is labelled with
    4-connectivity first; integer input is used as labels."""
    ar = np.asarray(ar)
    if min_size == 0:
        return ar.copy()
    if ar.dtype == bool:
        lab, _ = label4(ar)
    else:
        lab = np.ascontiguousarray(ar, dtype=np.int32).copy()
    lib().orc_remove_small(_p(lab), lab.size, int(min_size))
    if ar.dtype == bool:
        return lab > 0
    out = ar.copy()
    out[lab == 0] = 0
    return out


def watershed(image, markers=None, connectivity=1, offset=None, mask=None, compactness=0,
              watershed_line=False):
    """skimage.segmentation.watershed subset used by loader/postproc.py:378."""
    assert connectivity == 1 and compactness == 0 and not watershed_line and markers is not None
    image = np.ascontiguousarray(image, dtype=np.float64)
    if mask is None:
        mask = np.ones(image.shape, bool)
    mask8 = np.ascontiguousarray(np.asarray(mask, dtype=bool), dtype=np.uint8)
    mk = np.ascontiguousarray(np.asarray(markers) * mask8.astype(bool), dtype=np.int32)
    out = np.empty(image.shape, dtype=np.int32)
    lib().orc_watershed(_p(image), _p(mk), _p(mask8), image.shape[0], image.shape[1], _p(out))
    return out


def watershed_spec(image, markers, mask):
    """Pure-Python transcription of skimage 0.19 watershed_raveled + heap_general.pxi."""
    image = np.asarray(image, dtype=np.float64)
    H, W = image.shape
    PW = W + 2
    img = np.pad(image, 1).ravel()
    msk = np.pad(np.asarray(mask, dtype=bool), 1).ravel()
    out = np.pad((np.asarray(markers) * np.asarray(mask, dtype=bool)).astype(np.int32), 1).ravel().copy()
    heap = []

    def smaller(a, b):
        return a[0] < b[0] if a[0] != b[0] else a[1] < b[1]

    def push(e):
        heap.append(e)
        c = len(heap) - 1
        while c > 0:
            p = (c + 1) // 2 - 1
            if smaller(heap[c], heap[p]):
                heap[c], heap[p] = heap[p], heap[c]
                c = p
            else:
                break

    def pop():
        top = heap[0]
        last = heap.pop()
        n = len(heap)
        if n == 0:
            return top
        heap[0] = last
        i = 0
        while True:
            l, r = 2 * i + 1, 2 * i + 2
            s = i
            if l < n:
                if smaller(heap[l], heap[i]):
                    s = l
                if r < n and smaller(heap[r], heap[s]):
                    s = r
            else:
                break
            if s == i:
                break
            heap[i], heap[s] = heap[s], heap[i]
            i = s
        return top

    for idx in np.flatnonzero(out):
        push((img[idx], 0, int(idx)))
    age = 1
    while heap:
        v, a, idx = pop()
        for d in (-PW, -1, 1, PW):
            j = idx + d
            if not msk[j] or out[j]:
                continue
            age += 1
            out[j] = out[idx]
            push((img[j], age, j))
    return out.reshape(H + 2, PW)[1:-1, 1:-1].copy()


def proc_nuclei(inst_fg):
    fg = np.ascontiguousarray(inst_fg, dtype=np.float32)
    H, W, c = fg.shape
    out = np.empty((H, W), dtype=np.int32)
    ok = lib().orc_proc_nuclei(_p(fg), c, H, W, _p(out))
    if not ok:
        return np.zeros((H, W))  # float64, loader/postproc.py:380
    return out


def proc_gland_lumen(inst_fg, tissue, ds_factor=1.0):
    fg = np.ascontiguousarray(inst_fg, dtype=np.float32)
    H, W, c = fg.shape
    out = np.empty((H, W), dtype=np.int32)
    lib().orc_proc_gland_lumen(_p(fg), c, H, W, 0 if tissue.upper() == "GLAND" else 1,
                               ctypes.c_double(ds_factor), _p(out))
    return out.astype(np.float64)  # loader/postproc.py:290,331


def proc_eroded_map(inst_fg, tissue):
    """PostProcInstErodedMap.__proc_gland / __proc_lumen / __proc_nuclei (loader/postproc.py:149-236)."""
    fg = np.ascontiguousarray(inst_fg, dtype=np.float32)
    if fg.ndim == 2:
        fg = fg[..., None]
    H, W, c = fg.shape
    out = np.empty((H, W), dtype=np.int32)
    lib().orc_proc_eroded_map(_p(fg), c, H, W, {"GLAND": 0, "LUMEN": 1, "NUCLEI": 2}[tissue.upper()],
                              _p(out))
    return out.astype(np.float64)  # loader/postproc.py:157,187,217


def post_process_eroded(raw_map, idx_dict, tissue_mode, scale=1.0):
    """PostProcInstErodedMap.post_process (loader/postproc.py:238-265): the type map is the raw
    channel slice, NOT squeezed."""
    lo, hi = idx_dict[tissue_mode + "-INST"]
    inst_map = proc_eroded_map(raw_map[..., lo:hi], tissue_mode)
    type_ch = tissue_mode + "-TYPE"
    type_map = raw_map[..., idx_dict[type_ch][0]:idx_dict[type_ch][1]] if type_ch in idx_dict else None
    return inst_map, type_map


def post_process(raw_map, idx_dict, tissue_mode, ds_factor=1.0):
    """loader/postproc.py:383-407."""
    tissue_ch = tissue_mode + "-INST"
    lo, hi = idx_dict[tissue_ch]
    inst_fg = raw_map[..., lo:hi]
    if tissue_mode.upper() == "NUCLEI":
        inst_map = proc_nuclei(inst_fg)
    else:
        inst_map = proc_gland_lumen(inst_fg, tissue_mode, ds_factor)
    type_ch = tissue_mode + "-TYPE"
    if type_ch in idx_dict:
        type_map = np.squeeze(raw_map[..., idx_dict[type_ch][0]:idx_dict[type_ch][1]])
    else:
        type_map = None
    return inst_map, type_map
