"""GPU: the WSI path (cerberus_b200/infer/wsi.py) through the C ABI.

Kernel-level parity against numpy / OpenCV (the libraries the reference itself calls), and the
whole run_infer_wsi.py flow on a small synthetic slide against the CPU restatement of the WSI tail
(oracle/wsi_oracle.py) fed with the SAME merged prediction canvas."""
import ctypes
import os

import cv2
import joblib
import numpy as np
import pytest
import scipy.io as sio
import torch
import yaml

from cerberus_b200 import _lib, synth
from cerberus_b200.engine import Context
from cerberus_b200.infer.wsi import InferManager
from cerberus_b200.infer.wsi_geometry import filter_coordinates, get_coordinates
from cerberus_b200.infer.wsi_reader import ArraySlide

pytestmark = pytest.mark.gpu


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _dp(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.fixture(scope="module")
def ctx(built_lib):
    c = Context(0, "f16")
    yield c
    c.close()


def test_zero_padded_extract_matches_read_bounds(ctx):
    img = np.random.RandomState(1).randint(1, 256, (90, 130, 3)).astype(np.uint8)
    slide = ArraySlide(img, 0.5)
    tl = np.array([[-20, -30], [0, 0], [60, 100], [-200, -200], [85, 5]], dtype=np.int32)  # (y, x)
    out = np.empty((len(tl), 48, 48, 3), dtype=np.uint8)
    _lib.check(ctx.lib.cerb_extract_patches(ctx.handle, _vp(img), 90, 130, 0, 0, _vp(tl), len(tl), 48,
                                            48, _vp(out), 4), "extract")
    for i, (y, x) in enumerate(tl):
        assert np.array_equal(out[i], slide.read_bounds([x, y, x + 48, y + 48])), i


def test_scatter_writes_each_patch_once_and_clips(ctx):
    rng = np.random.RandomState(2)
    H, W, C, o = 100, 150, 9, 48
    _, pout = get_coordinates((W, H), [96, 96], [o, o], [o, o])
    patches = rng.rand(len(pout), o, o, C).astype(np.float32)
    ref = np.zeros((H, W, C), dtype=np.float32)
    for p, (x0, y0, x1, y1) in zip(patches, pout):
        ref[y0:y1, x0:x1] = p[:max(0, min(y1, H) - y0), :max(0, min(x1, W) - x0)]
    dev = torch.device("cuda", 0)
    canvas = torch.zeros((H, W, C), dtype=torch.float32, device=dev)
    pd = torch.from_numpy(patches).to(dev)
    torch.cuda.synchronize()
    tl = np.ascontiguousarray(pout[:, [1, 0]], dtype=np.int32)
    _lib.check(ctx.lib.cerb_scatter_patches(ctx.handle, _dp(pd), len(pout), o, o, C, _vp(tl), _dp(canvas),
                                            H, W), "scatter")
    _lib.check(ctx.lib.cerb_ctx_sync(ctx.handle), "sync")  # device-resident plumbing calls are asynchronous
    assert np.array_equal(canvas.cpu().numpy(), ref)


@pytest.mark.parametrize("chans", [[2, 3, 7], [0, 1], [5]])
@pytest.mark.parametrize("shape", [(64, 80), (63, 81), (37, 50), (5, 7)])
def test_region_half_matches_cv2_bilinear(ctx, shape, chans):
    """cv2.resize(crop * mask, (0,0), fx=0.5, fy=0.5) - bit-exact, odd sizes included, for the
    channel counts the WSI path uses (Gland: INST+TYPE = 3, Lumen: INST = 2)."""
    h, w = shape
    rng = np.random.RandomState(h * 100 + w)
    H, W, C = h + 11, w + 9, 9
    canvas = rng.rand(H, W, C).astype(np.float32)
    mask = (rng.rand(h, w) > 0.3).astype(np.uint8)
    chans = np.array(chans, dtype=np.int32)
    k = len(chans)
    y0, x0 = 6, 4
    ref = cv2.resize(canvas[y0:y0 + h, x0:x0 + w][..., chans] * mask[..., None], (0, 0), fx=0.5, fy=0.5)
    oh, ow = ref.shape[:2]
    dev = torch.device("cuda", 0)
    cd = torch.from_numpy(canvas).to(dev)
    out = torch.empty((oh, ow, k), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    _lib.check(ctx.lib.cerb_region_half(ctx.handle, _dp(cd), H, W, C, y0, x0, h, w, _vp(mask), _vp(chans),
                                        k, _dp(out), oh, ow), "region_half")
    assert np.array_equal(out.cpu().numpy(), ref.reshape(oh, ow, k))


def test_nearest_channel_matches_cv2(ctx):
    rng = np.random.RandomState(5)
    for (H, W) in ((100, 150), (101, 147), (8, 9)):
        canvas = rng.rand(H, W, 9).astype(np.float32)
        ref = cv2.resize(np.ascontiguousarray(canvas[..., 8]), (0, 0), fx=0.25, fy=0.25,
                         interpolation=cv2.INTER_NEAREST)
        out = np.empty(ref.shape, dtype=np.float32)
        cd = torch.from_numpy(canvas).to("cuda:0")
        torch.cuda.synchronize()
        _lib.check(ctx.lib.cerb_nearest_channel(ctx.handle, _dp(cd), H, W, 9, 8, 0.25, _vp(out),
                                                ref.shape[0], ref.shape[1]), "nearest")
        assert np.array_equal(out, ref), (H, W)


def _write_case(tmp, H=700, W=900):
    """Synthetic slide: H&E-like texture tiles mosaicked to H x W, tissue mask with two blobs at
    half resolution."""
    n_y, n_x = -(-H // 256), -(-W // 256)
    tiles = synth.synthetic_tiles(n_y * n_x, 256, 256, seed=77)
    slide = tiles.reshape(n_y, n_x, 256, 256, 3).transpose(0, 2, 1, 3, 4).reshape(n_y * 256, n_x * 256, 3)
    slide = np.ascontiguousarray(slide[:H, :W])
    os.makedirs(os.path.join(tmp, "wsi"), exist_ok=True)
    os.makedirs(os.path.join(tmp, "msk"), exist_ok=True)
    np.save(os.path.join(tmp, "wsi", "slideA.npy"), slide)
    mask = np.zeros((H // 2, W // 2), dtype=np.uint8)
    cv2.ellipse(mask, (W // 8, H // 4), (W // 10, H // 5), 0, 0, 360, 255, -1)
    cv2.ellipse(mask, (W // 3, H // 4), (W // 9, H // 6), 0, 0, 360, 255, -1)
    cv2.imwrite(os.path.join(tmp, "msk", "slideA.png"), mask)
    return slide, mask


def _table_key(v):
    return tuple(int(x) for x in np.asarray(v["box"]).flatten()) + (v.get("type"),)


def test_process_wsi_list_matches_cpu_restatement(built_lib, tmp_path):
    from oracle import wsi_oracle
    tmp = str(tmp_path)
    slide, mask_png = _write_case(tmp)
    H, W = slide.shape[:2]
    model_dir = os.path.join(tmp, "model")
    synth.write_model_dir(model_dir, seed=0)
    st = yaml.full_load(open(os.path.join(model_dir, "settings.yml")))
    m = InferManager(checkpoint_path=os.path.join(model_dir, "weights.tar"),
                     decoder_dict=st["dataset_kwargs"]["req_target_code"], model_args=st["model_kwargs"], precision="f16")
    m.keep_canvas = True
    run_args = {
        "nr_inference_workers": 0, "nr_post_proc_workers": 0, "batch_size": 6,
        "input_list": [os.path.join(tmp, "wsi", "slideA.npy")],
        "mask_list": [os.path.join(tmp, "msk", "slideA.png")],
        "output_dir": os.path.join(tmp, "out"), "patch_input_shape": 448, "patch_output_shape": 144,
        "save_thumb": True, "save_mask": True, "mask_dir": os.path.join(tmp, "msk") + "/",
        "postproc_list": ["gland", "lumen", "nuclei", "patch-class"], "msk_dir": os.path.join(tmp, "msk") + "/",
        "tile_shape": 2048, "chunk_shape": 15000, "ambiguous_size": 64,
        "cache_path": os.path.join(tmp, "cache"), "logging_dir": os.path.join(tmp, "log"),
        "wsi_proc_mag": 0.5,
        "postproc_tile_shape": 300,  # test hook: -> 288-pixel tiles, so strips and crosses exist
    }
    res = m.process_wsi_list(run_args)["slideA"]
    out = joblib.load(os.path.join(tmp, "out", "dat", "slideA.dat"))
    assert set(out.keys()) >= {"Nuclei", "proc_resolution", "base_resolution", "proc_dimensions"}
    assert os.path.exists(os.path.join(tmp, "out", "thumb", "slideA.png"))
    assert os.path.exists(os.path.join(tmp, "out", "mask", "slideA.png"))

    # (1) merged canvas == every selected patch through run_step, placed with numpy
    canvas = m.last_canvas.cpu().numpy()
    wsi_mask = (cv2.cvtColor(cv2.imread(os.path.join(tmp, "msk", "slideA.png")), cv2.COLOR_BGR2GRAY) > 0).astype(np.uint8)
    pin, pout = get_coordinates((W, H), [448, 448], [144, 144], [144, 144])
    sel = filter_coordinates(wsi_mask, pout, (H, W))
    assert 0 < sel.sum() < len(sel)  # the mask really filters
    pin, pout = pin[sel], pout[sel]
    rd = ArraySlide(slide, 0.5)
    idx = m.engine.model.idx_dict
    ref_canvas = np.zeros_like(canvas)
    for s in range(0, len(pin), 6):
        batch = np.stack([rd.read_bounds(b) for b in pin[s:s + 6]])
        pad = 6 - len(batch)
        if pad:
            batch = np.concatenate([batch, np.zeros((pad, 448, 448, 3), np.uint8)])
        outs = m.run_step(batch, 144)
        for o, (x0, y0, x1, y1) in zip(outs, pout[s:s + 6]):
            hh, ww = max(0, min(y1, H) - y0), max(0, min(x1, W) - x0)
            for k, v in o.items():
                lo, hi = idx[k]
                vv = v if v.ndim == 3 else v[..., None]
                ref_canvas[y0:y0 + hh, x0:x0 + ww, lo:hi] = vv[:hh, :ww]
    assert np.array_equal(canvas, ref_canvas)

    # (2) instance tables == CPU restatement of the WSI tail on the same canvas
    ref_nuc = wsi_oracle.nuclei_tables(canvas, idx, pout, [300, 300], [144, 144], 64)
    got = sorted(_table_key(v) for v in out["Nuclei"].values())
    want = sorted(_table_key(v) for v in ref_nuc)
    assert len(want) > 50, "the synthetic slide must give the de-duplication real work"
    assert got == want
    ref_gl = wsi_oracle.gland_lumen_tables(canvas, idx, wsi_mask)
    for t in ("Gland", "Lumen"):
        got = sorted(_table_key(v) for v in out.get(t, {}).values())
        want = sorted(_table_key(v) for v in ref_gl[t])
        assert got == want, t
    assert len(ref_gl["Gland"]) > 0

    # (3) Patch-Class map
    pc = sio.loadmat(os.path.join(tmp, "out", "tissue", "slideA.mat"))["pclass"]
    ref_pc = cv2.resize(np.ascontiguousarray(canvas[..., idx["Patch-Class"][0]]), (0, 0), fx=0.25, fy=0.25,
                        interpolation=cv2.INTER_NEAREST)
    ref_pc = ref_pc * cv2.resize(wsi_mask, (ref_pc.shape[1], ref_pc.shape[0]), interpolation=cv2.INTER_NEAREST)
    assert np.array_equal(pc, ref_pc)

    # (4) skip-if-done (infer/wsi.py:969)
    t0 = os.path.getmtime(os.path.join(tmp, "out", "dat", "slideA.dat"))
    assert m.process_wsi_list(run_args) == {}
    assert os.path.getmtime(os.path.join(tmp, "out", "dat", "slideA.dat")) == t0
    m.engine.close()
