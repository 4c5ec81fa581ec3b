"""Mirror of infer/tile.py (reference): tile-mode inference manager.

Same public surface — `_prepare_patching`, `_post_process_patches`,
`InferManager.process_file_list(run_args)` with the run_args keys of run_infer_tile.py:53-63 —
and the same on-disk outputs (<tissue>_mat/<name>.mat, pclass_mat/<name>.mat, overlay/<name>.jpg).
What changes is where the work runs:
  * reflect padding + patch slicing (tile.py:64-69, infer_loader.py:57-69): one device kernel
    (cerb_extract_patches), the padded image is never materialised;
  * the duplicate patch grid the reference appends when overlap == 0 (tile.py:90-103, every
    patch inferred twice and averaged) is inferred ONCE; (a + a) / 2 == a exactly in fp32, so
    the stitched canvas is bit-identical;
  * stitching (tile.py:136-163): cerb_stitch; post-processing (tile.py:173-191): device kernels.
"""
import ctypes
import math
import os
import pathlib

import cv2
import numpy as np
import scipy.io as sio

from .. import _lib
from ..instinfo import get_inst_info_dict
from ..postproc import PostProcInstErodedContourMap, PostProcInstErodedMap
from . import base

# infer/tile.py:35-40 (the IP-ERODED-3/11 codes are not used by any shipped settings: SURVEY.md 8f-4)
_postproc_func_dict = {
    "IP-ERODED-3": PostProcInstErodedMap,
    "IP-ERODED-11": PostProcInstErodedMap,
    "IP-ERODED-CONTOUR-3": PostProcInstErodedContourMap,
    "IP-ERODED-CONTOUR-11": PostProcInstErodedContourMap,
}


def patch_grid(im_h, im_w, input_size, output_size, output_overlap_size=0):
    """Geometry part of _prepare_patching (tile.py:43-106) without touching pixels.
    Returns (info_list [n,2,2,2] int32, [padt, padl], padded_shape)."""
    win_size = input_size
    msk_size = step_size = output_size

    def get_last_steps(length, msk_size, step_size):
        nr_step = math.ceil((length - msk_size) / step_size)
        last_step = (nr_step + 1) * step_size
        return int(last_step), int(nr_step + 1)

    last_h, _ = get_last_steps(im_h, msk_size, output_size)
    last_w, _ = get_last_steps(im_w, msk_size, output_size)
    diff = win_size - step_size
    padt = padl = diff // 2
    padb = last_h + win_size - im_h
    padr = last_w + win_size - im_w
    padded_shape = np.array([im_h + padt + padb, im_w + padl + padr])

    input_tl_y = np.arange(0, last_h, step_size, dtype=np.int32)
    input_tl_x = np.arange(0, last_w, step_size, dtype=np.int32)
    input_tl_y, input_tl_x = np.meshgrid(input_tl_y, input_tl_x)  # 'xy': column-major patch order
    input_tl = np.stack([input_tl_y.flatten(), input_tl_x.flatten()], axis=-1)
    output_tl = input_tl + diff // 2
    output_br = output_tl + output_size
    input_br = input_tl + input_size
    sel = np.any(input_br > padded_shape, axis=-1)
    info_list = np.stack(
        [np.stack([input_tl[~sel], input_br[~sel]], axis=1),
         np.stack([output_tl[~sel], output_br[~sel]], axis=1)], axis=1)
    if output_overlap_size == 0:  # tile.py:90-103: the grid is appended a second time
        ovl_output_tl = output_tl + output_overlap_size
        ovl_input_tl = ovl_output_tl - diff // 2
        ovl_output_br = ovl_output_tl + output_size
        ovl_input_br = ovl_input_tl + input_size
        sel = np.any(ovl_input_br > padded_shape, axis=-1)
        ovl = np.stack(
            [np.stack([ovl_input_tl[~sel], ovl_input_br[~sel]], axis=1),
             np.stack([ovl_output_tl[~sel], ovl_output_br[~sel]], axis=1)], axis=1)
        info_list = np.concatenate([info_list, ovl], axis=0)
    return info_list, [padt, padl], (padt, padb, padl, padr)


def _prepare_patching(img, input_size, output_size, output_overlap_size):
    """Drop-in for tile.py:43-106 (host version: returns the reflect-padded image)."""
    info_list, src_pos, (padt, padb, padl, padr) = patch_grid(
        img.shape[0], img.shape[1], input_size, output_size, output_overlap_size)
    padded_img = np.pad(img, ((padt, padb), (padl, padr), (0, 0)), "reflect")
    return padded_img, info_list, src_pos


def idx_dict_of(model_args):
    from ..plan import canvas_layout
    return canvas_layout(model_args["decoder_kwargs"])


def stitch_canvas(ctx, patch_canvas, out_tl, canvas_hw, src_pos, src_shape):
    """tile.py:136-163 on the device. patch_canvas: float32 [n,oh,ow,C] (host)."""
    n, oh, ow, C = patch_canvas.shape
    out = np.empty((src_shape[0], src_shape[1], C), dtype=np.float32)
    tl = np.ascontiguousarray(out_tl, dtype=np.int32)
    pc = np.ascontiguousarray(patch_canvas, dtype=np.float32)
    _lib.check(ctx.lib.cerb_stitch(ctx.handle, pc.ctypes.data_as(ctypes.c_void_p), n, oh, ow, C,
                                   tl.ctypes.data_as(ctypes.c_void_p), int(canvas_hw[0]),
                                   int(canvas_hw[1]), int(src_pos[0]), int(src_pos[1]),
                                   int(src_shape[0]), int(src_shape[1]),
                                   out.ctypes.data_as(ctypes.c_void_p), 0), "cerb_stitch")
    return out


def _post_process_patches(patch_info_list, image_info, postproc_code=None, postproc_list=None,
                          model_args=None, ctx=None):
    """Drop-in for tile.py:109-212. patch_info_list: [(pdata dict, (out_tl, out_br), file_idx)].
    Stitching and post-processing run on the device bound to PostProcInstErodedContourMap."""
    src_pos, src_shape = image_info["src_pos"], image_info["src_shape"]
    idx_dict, nr_out_chs = idx_dict_of(model_args)
    ctx = ctx if ctx is not None else PostProcInstErodedContourMap._ctx
    ch_code_list = list(patch_info_list[0][0].keys())
    out_br_list = np.array([v[1][1] for v in patch_info_list])
    hw = np.max(out_br_list, axis=0).tolist()
    # per-patch canvases in the idx_dict channel layout (what the head kernel writes)
    n = len(patch_info_list)
    oh, ow = (np.array(patch_info_list[0][1][1]) - np.array(patch_info_list[0][1][0])).tolist()
    pc = np.zeros((n, oh, ow, nr_out_chs), dtype=np.float32)
    tl = np.zeros((n, 2), dtype=np.int32)
    for i, (pdata, (patch_tl, patch_br), _) in enumerate(patch_info_list):
        tl[i] = patch_tl
        for ch_code, ch_val in pdata.items():
            if ch_val.ndim == 2:
                ch_val = np.expand_dims(ch_val, -1)
            lo, hi = idx_dict[ch_code]
            pc[i, ..., lo:hi] = ch_val
    raw_canvas = stitch_canvas(ctx, pc, tl, hw, src_pos, src_shape[:2])

    pred_inst_map_dict, pred_type_map_dict, pred_inst_info_dict = {}, {}, {}
    pclass_map = None
    for tissue_code in postproc_list:
        tissue_code = tissue_code.capitalize()
        if tissue_code + "-INST" in postproc_code.keys():
            code = postproc_code[tissue_code + "-INST"]
            if code not in _postproc_func_dict:
                raise NotImplementedError("post-proc code %r is outside the hot path (only "
                                          "IP-ERODED-CONTOUR-* is shipped)" % code)
            proc_func = _postproc_func_dict[code]
            inst_map, type_map = proc_func.post_process(raw_canvas, idx_dict, tissue_code)
            pred_inst_map_dict[tissue_code] = inst_map
            pred_type_map_dict[tissue_code] = type_map
        elif tissue_code == "Patch-class":
            pclass_map = raw_canvas[..., idx_dict["Patch-Class"][0]]

    if "lumen" in postproc_list and "gland" in postproc_list:  # tile.py:187-191
        binary_gland = pred_inst_map_dict["Gland"].copy()
        binary_gland[binary_gland > 0] = 1
        pred_inst_map_dict["Lumen"] = binary_gland * pred_inst_map_dict["Lumen"]

    # tile.py:193-203: the reference resizes the instance / type maps x2 (cv2 INTER_NEAREST) and
    # calls get_inst_info_dict on the copies; the device version addresses the upsampled image
    # through `up=2` without materialising it.
    pred_type_tmp = None
    for tissue_code in postproc_list:
        tissue_code = tissue_code.capitalize()
        if tissue_code != "Patch-class":
            if tissue_code != "Lumen":
                if pred_type_map_dict[tissue_code] is not None:
                    pred_type_tmp = pred_type_map_dict[tissue_code]
            # reference quirk (tile.py:193-203): Lumen inherits the previous tissue's type map
            pred_inst_info_dict[tissue_code] = get_inst_info_dict(
                pred_inst_map_dict[tissue_code], pred_type_tmp, ctx=ctx, up=2)

    return (image_info["name"], image_info["src_image"], pred_inst_map_dict, pred_inst_info_dict,
            pred_type_map_dict, pclass_map)


_OVERLAY_COLOURS = {"Gland": (255, 165, 0), "Lumen": (0, 255, 0), "Nuclei": (0, 0, 255)}


def _overlay(src_image, inst_info_dict):
    """Plain contour overlay (the reference's visualize_instances_dict_orig + dataset.yml colour
    table is visual-only and out of the hot path)."""
    out = src_image.copy()
    for tissue, info in inst_info_dict.items():
        cnts = [v["contour"].reshape(-1, 1, 2).astype(np.int32) for v in info.values()]
        cv2.drawContours(out, cnts, -1, _OVERLAY_COLOURS.get(tissue, (255, 255, 0)), 2)
    return out


def recur_find_ext(root_dir, ext_list):
    """misc/utils.py:250-265."""
    file_path_list = []
    for cur_path, _, file_list in os.walk(root_dir):
        for file_name in file_list:
            if pathlib.Path(file_name).suffix in ext_list:
                file_path_list.append(os.path.join(cur_path, file_name))
    file_path_list.sort()
    return file_path_list


class InferManager(base.InferManager):
    """Run inference on tiles (mirror of tile.py:215-429)."""

    def extract_patches(self, img, tl_yx, patch_size, pad_tl):
        """Device reflect-pad + slice: uint8 [n,patch,patch,3] for the given padded-space
        top-lefts (tile.py:64-69 + infer_loader.py:57-69)."""
        ctx = self.engine.ctx
        img = np.ascontiguousarray(img, dtype=np.uint8)
        tl = np.ascontiguousarray(tl_yx, dtype=np.int32)
        out = np.empty((tl.shape[0], patch_size, patch_size, 3), dtype=np.uint8)
        _lib.check(ctx.lib.cerb_extract_patches(
            ctx.handle, img.ctypes.data_as(ctypes.c_void_p), img.shape[0], img.shape[1],
            int(pad_tl[0]), int(pad_tl[1]), tl.ctypes.data_as(ctypes.c_void_p), tl.shape[0],
            patch_size, patch_size, out.ctypes.data_as(ctypes.c_void_p), 0), "cerb_extract_patches")
        return out

    # ------------------------------------------------------------------ device-resident path
    def _dev(self, key, nbytes, ctx=None):
        """Grow-only device buffer owned by the manager, one pool per context (the finishing
        stage runs on its own context / thread and must not share buffers with the forward)."""
        ctx = ctx if ctx is not None else self.engine.ctx
        pools = self.__dict__.setdefault("_dev_bufs", {})
        bufs = pools.setdefault(id(ctx), {"ctx": ctx})
        cur = bufs.get(key)
        if cur is None or cur[1] < nbytes:
            if cur is not None:
                ctx.sync()
                ctx.lib.cerb_dev_free(ctx.handle, ctypes.c_void_p(cur[0]))
            p = ctx.lib.cerb_dev_alloc(ctx.handle, int(nbytes))
            if not p:
                _lib.check(-1, "cerb_dev_alloc(%d)" % nbytes)
            bufs[key] = cur = (p, int(nbytes))
        return cur[0]

    def release_device_buffers(self):
        for bufs in self.__dict__.get("_dev_bufs", {}).values():
            ctx = bufs["ctx"]
            if getattr(ctx, "handle", None):
                for k, v in bufs.items():
                    if k != "ctx":
                        ctx.lib.cerb_dev_free(ctx.handle, ctypes.c_void_p(v[0]))
        self._dev_bufs = {}
        for fin in self.__dict__.pop("_finish_ctxs", []):
            fin.close()
        self.__dict__.pop("_finish_pool", None)

    def _sync_forward(self):
        self.engine.ctx.sync()

    def _finisher_ctx(self):
        """Borrows a context (own stream, own post-processing workspaces) for one finishing task;
        give it back with `_release_finisher_ctx`. Contexts are created on demand and reused
        across `process_file_list` calls (at most one per concurrently running finishing thread)."""
        import queue
        pool = self.__dict__.setdefault("_finish_pool", queue.SimpleQueue())
        try:
            return pool.get_nowait()
        except queue.Empty:
            from ..engine import Context
            ctx = Context(self.engine.ctx.device, self.engine.ctx.precision)
            self.__dict__.setdefault("_finish_ctxs", []).append(ctx)
            return ctx

    def _release_finisher_ctx(self, ctx):
        self._finish_pool.put(ctx)

    def _finish_one(self, m, group, to_host=True):
        """Worker-thread task: a16-a20 for one image of a forwarded group."""
        ctx = self._finisher_ctx()
        try:
            return self._finish_image(m, group["store"] + m["first"] * group["cbytes"], group["P_out"],
                                      group["C"], self.engine.model.idx_dict, to_host, ctx)
        finally:
            self._release_finisher_ctx(ctx)

    def process_images(self, named_images, to_host=True):
        """A group of RGB uint8 images -> one result tuple per image (the tuple
        _post_process_patches returns). This is what `process_file_list` runs per cache group
        (the reference caches files until > 256 patches and batches across them,
        infer/tile.py:294-325). Everything between the uploaded image and the label maps stays in
        HBM:  image H2D -> cerb_extract_patches (reflect pad + slicing) straight into the batch
        buffer -> forward plan of `batch_size` patches (batches are filled ACROSS images; the tail
        of the last batch runs on stale patches whose outputs are ignored) -> per-patch canvases
        collected in a device store -> cerb_stitch -> post-processing -> instance tables.
        Per image the host receives the label maps, the type / Patch-Class planes and the
        instance tables, nothing else. Every distinct patch is inferred once: the duplicate grid
        the reference appends (tile.py:90-103) averages a patch with itself, which is exact."""
        import time as _time
        tm = self.__dict__.setdefault("stage_seconds", {"extract": 0.0, "forward": 0.0, "finish": 0.0})
        group = self._forward_group(named_images, slot=0)
        self.engine.ctx.sync()
        t0 = _time.perf_counter()
        out = self._finish_group(group, self.engine.ctx, to_host)
        tm["finish"] += _time.perf_counter() - t0
        return out

    def _forward_group(self, named_images, slot=0):
        """Stage 1 (forward context): upload, extract, forward; the per-patch canvases of the
        group end up in the device store `slot`. Asynchronous - the caller synchronises."""
        import time as _time
        eng = self.engine
        ctx, lib, model = eng.ctx, eng.ctx.lib, eng.model
        tm = self.__dict__.setdefault("stage_seconds", {"extract": 0.0, "forward": 0.0, "finish": 0.0})
        t_stage = _time.perf_counter()
        P_in, P_out = int(self.patch_input_shape), int(self.patch_output_shape)
        B = max(1, int(self.batch_size))
        C = model.canvas_c
        plan = eng.plan_for(B, P_in, P_in, P_out, P_out)
        metas, total = [], 0
        for name, img in named_images:
            img = np.ascontiguousarray(img, dtype=np.uint8)
            info_list, src_pos, _ = patch_grid(img.shape[0], img.shape[1], P_in, P_out,
                                               self.patch_output_overlap)
            in_tl = info_list[:, 0, 0, :]
            _, first = np.unique(in_tl, axis=0, return_index=True)
            order = np.sort(first)
            hw = np.max(info_list[:, 1, 1, :], axis=0)
            metas.append({"name": name, "img": img, "src_pos": src_pos, "first": total,
                          "in_tl": np.ascontiguousarray(in_tl[order], dtype=np.int32),
                          "out_tl": np.ascontiguousarray(info_list[order, 1, 0, :], dtype=np.int32),
                          "canvas_hw": (int(hw[0]), int(hw[1]))})
            total += len(order)
        nb = (total + B - 1) // B
        pbytes, cbytes = P_in * P_in * 3, P_out * P_out * C * 4
        d_patches = self._dev("patches", nb * B * pbytes)
        d_store = self._dev("store%d" % slot, nb * B * cbytes)
        max_img = max(m["img"].nbytes for m in metas)
        d_img = self._dev("img", max_img)
        for m in metas:  # a1/a2: reflect pad + slicing on the device, into the batch buffer
            img = m["img"]
            _lib.check(lib.cerb_memcpy(ctx.handle, ctypes.c_void_p(d_img),
                                       img.ctypes.data_as(ctypes.c_void_p), img.nbytes, 1), "image H2D")
            _lib.check(lib.cerb_extract_patches(
                ctx.handle, ctypes.c_void_p(d_img), img.shape[0], img.shape[1], int(m["src_pos"][0]),
                int(m["src_pos"][1]), m["in_tl"].ctypes.data_as(ctypes.c_void_p), len(m["in_tl"]),
                P_in, P_in, ctypes.c_void_p(d_patches + m["first"] * pbytes), 3), "cerb_extract_patches")
            ctx.sync()  # d_img is reused by the next image
        tm["extract"] += _time.perf_counter() - t_stage
        t_stage = _time.perf_counter()
        canvas_ptr = plan.tensor_ptr(plan.spec.canvas)
        for b in range(nb):  # a3-a15
            plan.run(device_ptr=d_patches + b * B * pbytes)
            _lib.check(lib.cerb_memcpy(ctx.handle, ctypes.c_void_p(d_store + b * B * cbytes),
                                       ctypes.c_void_p(canvas_ptr), B * cbytes, 3), "canvas -> store")
        self.nr_patches_inferred = getattr(self, "nr_patches_inferred", 0) + total
        if getattr(self, "time_stages", False):
            ctx.sync()
        tm["forward"] += _time.perf_counter() - t_stage
        return {"metas": metas, "store": d_store, "cbytes": cbytes, "P_out": P_out, "C": C}

    def _finish_group(self, group, ctx, to_host=True):
        """Stage 2 (any context): a16-a20 for every image of a forwarded group."""
        idx_dict = self.engine.model.idx_dict
        return [self._finish_image(m, group["store"] + m["first"] * group["cbytes"], group["P_out"],
                                   group["C"], idx_dict, to_host, ctx) for m in group["metas"]]

    def _finish_image(self, m, d_patch_canvas, P_out, C, idx_dict, to_host, ctx=None):
        """a16-a20 for one image, device-resident: stitch, post-process, instance tables."""
        import time as _time
        ctx = ctx if ctx is not None else self.engine.ctx
        lib = ctx.lib
        img = m["img"]
        H, W = img.shape[:2]
        n = len(m["out_tl"])
        hw4 = H * W * 4
        fine = self.__dict__.setdefault("finish_seconds", {}) if getattr(self, "time_stages", False) else None
        t_ph = [_time.perf_counter()]

        def phase(name):
            if fine is not None:
                ctx.sync()
                now = _time.perf_counter()
                fine[name] = fine.get(name, 0.0) + now - t_ph[0]
                t_ph[0] = now

        d_canvas = self._dev("canvas", H * W * C * 4, ctx)
        _lib.check(lib.cerb_stitch(ctx.handle, ctypes.c_void_p(d_patch_canvas), n, P_out, P_out, C,
                                   m["out_tl"].ctypes.data_as(ctypes.c_void_p), m["canvas_hw"][0],
                                   m["canvas_hw"][1], int(m["src_pos"][0]), int(m["src_pos"][1]), H, W,
                                   ctypes.c_void_p(d_canvas), 3), "cerb_stitch")
        self.last_canvas_dev = (d_canvas, H, W, C)
        phase("stitch")

        def plane(ch, key):
            """canvas[..., ch]: device plane (for the instance tables) + host copy (for the .mat)."""
            d = self._dev(key, hw4, ctx)
            _lib.check(lib.cerb_channel_plane(ctx.handle, ctypes.c_void_p(d_canvas), H, W, C, ch,
                                              ctypes.c_void_p(d), 2), "cerb_channel_plane")
            host = None
            if to_host:
                host = np.empty((H, W), dtype=np.float32)
                _lib.check(lib.cerb_memcpy(ctx.handle, host.ctypes.data_as(ctypes.c_void_p),
                                           ctypes.c_void_p(d), hw4, 2), "type plane D2H")
            return d, host

        inst_dev, inst_map_dict, type_map_dict, type_dev = {}, {}, {}, {}
        pclass_map = None
        d_flag = self._dev("any_fg", 256, ctx)
        for tissue_code in self.postproc_list:
            tissue_code = tissue_code.capitalize()
            if tissue_code + "-INST" in self.decoder_dict.keys():
                code = self.decoder_dict[tissue_code + "-INST"]
                if code not in _postproc_func_dict:
                    raise NotImplementedError("post-proc code %r is outside the hot path" % code)
                lo, hi = idx_dict[tissue_code + "-INST"]
                d_lab = self._dev("lab." + tissue_code, hw4, ctx)
                any_fg = 1
                if _postproc_func_dict[code] is PostProcInstErodedMap:
                    if hi - lo != 1:
                        raise ValueError("%s-INST must be a single channel for IP-ERODED-*" % tissue_code)
                    _lib.check(lib.cerb_postproc_eroded_map(
                        ctx.handle, ctypes.c_void_p(d_canvas), 1, H, W, C, lo,
                        {"Gland": 0, "Lumen": 1, "Nuclei": 2}[tissue_code], ctypes.c_void_p(d_lab), 3),
                        "cerb_postproc_eroded_map")
                    as_float = True
                elif tissue_code == "Nuclei":
                    _lib.check(lib.cerb_postproc_nuclei(ctx.handle, ctypes.c_void_p(d_canvas), 1, H, W, C,
                                                        lo, ctypes.c_void_p(d_lab),
                                                        ctypes.c_void_p(d_flag), 3), "cerb_postproc_nuclei")
                    flag = np.zeros(1, dtype=np.int32)
                    _lib.check(lib.cerb_memcpy(ctx.handle, flag.ctypes.data_as(ctypes.c_void_p),
                                               ctypes.c_void_p(d_flag), 4, 2), "any_fg D2H")
                    any_fg = int(flag[0])
                    as_float = False
                else:
                    _lib.check(lib.cerb_postproc_gland_lumen(
                        ctx.handle, ctypes.c_void_p(d_canvas), 1, H, W, C, lo,
                        0 if tissue_code == "Gland" else 1, 1.0, ctypes.c_void_p(d_lab), 3),
                        "cerb_postproc_gland_lumen")
                    as_float = True
                inst_dev[tissue_code] = (d_lab, as_float, any_fg)
                tkey = tissue_code + "-TYPE"
                if tkey in idx_dict:
                    type_dev[tissue_code], type_map_dict[tissue_code] = plane(idx_dict[tkey][0],
                                                                             "type." + tissue_code)
                else:
                    type_dev[tissue_code], type_map_dict[tissue_code] = None, None
            elif tissue_code == "Patch-class":
                _, pclass_map = plane(idx_dict["Patch-Class"][0], "pclass")
        phase("postproc+planes")
        if "lumen" in self.postproc_list and "gland" in self.postproc_list:  # tile.py:187-191
            _lib.check(lib.cerb_mask_lumen(ctx.handle, ctypes.c_void_p(inst_dev["Lumen"][0]),
                                           ctypes.c_void_p(inst_dev["Gland"][0]), H * W), "cerb_mask_lumen")
        # tile.py:193-203: x2 nearest resize of the maps, then get_inst_info_dict; the device
        # version addresses the upsampled image through up=2. Lumen inherits the previous
        # tissue's type map (the reference's stale `pred_type_tmp`).
        inst_info_dict = {}
        type_tmp = None
        for tissue_code in self.postproc_list:
            tissue_code = tissue_code.capitalize()
            if tissue_code == "Patch-class" or tissue_code not in inst_dev:
                continue
            d_lab, as_float, any_fg = inst_dev[tissue_code]
            if tissue_code != "Lumen" and type_dev[tissue_code] is not None:
                type_tmp = type_dev[tissue_code]
            # keys carry the dtype of the reference's label map (np.unique keeps it)
            kd = np.float64 if (as_float or not any_fg) else np.int32
            inst_info_dict[tissue_code] = get_inst_info_dict(
                d_lab, type_tmp, ctx=ctx, up=2, on_device=True, shape=(H, W), key_dtype=kd)
            phase("inst_info")
            if to_host:
                lab = np.empty((H, W), dtype=np.int32)
                _lib.check(lib.cerb_memcpy(ctx.handle, lab.ctypes.data_as(ctypes.c_void_p),
                                           ctypes.c_void_p(d_lab), hw4, 2), "labels D2H")
                # reference dtypes: nuclei int32 (float64 zeros when the mask is empty,
                # postproc.py:378-380), gland / lumen float64 (postproc.py:290,331)
                inst_map_dict[tissue_code] = lab.astype(np.float64) if (as_float or not any_fg) else lab
                phase("labels_d2h")
        return (m["name"], img, inst_map_dict, inst_info_dict, type_map_dict, pclass_map)

    def process_image(self, img, name="image"):
        """One RGB uint8 image -> the tuple _post_process_patches returns."""
        return self.process_images([(name, img)])[0]

    def process_image_host_plumbing(self, img, name="image"):
        """The same through the reference-shaped host plumbing (`run_step` list of dicts ->
        `_post_process_patches`): kept for API parity and as a cross-check of the device path."""
        PostProcInstErodedContourMap.bind(self.engine.ctx)
        info_list, src_pos, _ = patch_grid(img.shape[0], img.shape[1], self.patch_input_shape,
                                           self.patch_output_shape, self.patch_output_overlap)
        in_tl = info_list[:, 0, 0, :]
        uniq, first, inverse = np.unique(in_tl, axis=0, return_index=True, return_inverse=True)
        order = np.sort(first)
        remap = {int(f): i for i, f in enumerate(order)}
        patches = self.extract_patches(img, in_tl[order], self.patch_input_shape, src_pos)
        outputs = []
        for s in range(0, len(order), self.batch_size):
            outputs.extend(self.run_step(patches[s:s + self.batch_size], self.patch_output_shape))
        first_of = first[np.asarray(inverse).reshape(-1)]
        patch_info_list = [(outputs[remap[int(first_of[i])]], (info_list[i, 1, 0], info_list[i, 1, 1]), 0)
                           for i in range(info_list.shape[0])]
        image_info = {"src_pos": src_pos, "src_shape": img.shape[:2], "src_image": img, "name": name}
        return _post_process_patches(patch_info_list, image_info, self.decoder_dict,
                                     self.postproc_list, self.model_args, ctx=self.engine.ctx)

    def process_file_list(self, run_args):
        """Process image tiles < 5000x5000 (tile.py:218-429): same skip-if-done resume rule,
        same outputs. Files are cached into groups until more than 256 patches are pending
        (tile.py:294-325) and each group goes through `process_images`; under
        `torchrun` / `--gpu=0,1,..` (run_infer_tile.py) the sorted file list is sharded over the
        ranks at file granularity, so stitching and post-processing stay rank-local."""
        for variable, value in run_args.items():
            self.__setattr__(variable, value)
        file_path_list_all = recur_find_ext(self.input_dir, [".png", ".jpg"])
        file_path_list = []
        for file_path in file_path_list_all:
            base_name = os.path.basename(file_path).split(".")[0]
            missing = 0
            for tissue_check in self.postproc_list:
                if not os.path.exists("%s/%s_mat/%s.mat" % (self.output_dir, tissue_check, base_name)):
                    missing += 1
            if missing > 0:
                file_path_list.append(file_path)
        file_path_list.sort()
        assert len(file_path_list) > 0, "Not Detected Any Files From Path"
        rank, world = int(getattr(self, "rank", 0)), int(getattr(self, "world_size", 1))
        if world > 1:
            from ..dist import shard_units
            file_path_list = [file_path_list[i] for i in shard_units(len(file_path_list), rank, world)]
        # The reference decodes images in DataLoader worker processes (--nr_inference_workers) and
        # post-processes / saves in a ProcessPoolExecutor (--nr_post_proc_workers). Here the forward
        # stays on this thread and the finishing stage on worker threads with their own contexts (a
        # cerb_ctx is single-threaded); the same two flags size a
        # loader thread pool (PNG decode of the next images) and a writer thread pool (overlay,
        # .mat files of the previous images), both of which spend their time inside OpenCV /
        # scipy.io with the GIL released. 0 workers = everything inline, as in the reference.
        from concurrent.futures import ThreadPoolExecutor
        n_load = int(getattr(self, "nr_inference_workers", 0) or 0)
        n_save = int(getattr(self, "nr_post_proc_workers", 0) or 0)
        loaders = ThreadPoolExecutor(n_load) if n_load > 0 else None
        savers = ThreadPoolExecutor(n_save) if n_save > 0 else None
        ahead = max(2 * n_load, 1)
        pending_saves = []
        # finishing stage: --nr_post_proc_workers threads (1..4), each with its own context
        finisher = ThreadPoolExecutor(max(1, min(n_save, 4)))
        inflight, n_groups = None, 0

        def deliver(group, futs):
            import time as _time
            t0 = _time.perf_counter()
            results = [f.result() for f in futs]
            tm = self.__dict__.setdefault("stage_seconds", {"extract": 0.0, "forward": 0.0, "finish": 0.0})
            tm["finish"] += _time.perf_counter() - t0  # time the main thread WAITED for the finisher
            for (file_path, _), res in zip(group, results):
                if savers is not None:
                    pending_saves.append((file_path, savers.submit(self._save, res, self.output_dir)))
                    while len(pending_saves) > 4 * n_save:  # bound the results held in memory
                        done_path, f2 = pending_saves.pop(0)
                        f2.result()
                        print("Done Assembling %s" % done_path)
                else:
                    self._save(res, self.output_dir)
                    print("Done Assembling %s" % file_path)

        try:
            loads = {}
            i = 0
            while i < len(file_path_list):
                # cache files until more than 256 patch-grid entries are pending (tile.py:322-323)
                group, pending = [], 0
                while i < len(file_path_list):
                    file_path = file_path_list[i]
                    if loaders is not None:
                        for j in range(i, min(i + ahead, len(file_path_list))):
                            if j not in loads:
                                loads[j] = loaders.submit(self._load_rgb, file_path_list[j])
                        img = loads.pop(i).result()
                    else:
                        img = self._load_rgb(file_path)
                    i += 1
                    group.append((file_path, img))
                    info_list, _, _ = patch_grid(img.shape[0], img.shape[1], self.patch_input_shape,
                                                 self.patch_output_shape, self.patch_output_overlap)
                    pending += info_list.shape[0]
                    if pending > 256:
                        break
                # Two-stage pipeline: the forward of THIS group runs on the engine's context while
                # worker threads finish (stitch / post-processing / instance tables) the images of
                # the PREVIOUS group on their own contexts - the finishing stage is many small
                # kernels and host work, the forward keeps the GPU busy. Two device stores alternate.
                fwd = self._forward_group([(pathlib.Path(fp).stem, im) for fp, im in group], slot=n_groups & 1)
                n_groups += 1
                self._sync_forward()
                if inflight is not None:
                    deliver(*inflight)
                inflight = (group, [finisher.submit(self._finish_one, m, fwd) for m in fwd["metas"]])
            if inflight is not None:
                deliver(*inflight)
                inflight = None
            for done_path, fut in pending_saves:  # a failed write raises here: no silent crash
                fut.result()
                print("Done Assembling %s" % done_path)
        finally:
            finisher.shutdown(wait=True)
            for pool in (loaders, savers):
                if pool is not None:
                    pool.shutdown(wait=True)
        return

    @staticmethod
    def _load_rgb(file_path):
        img = cv2.imread(file_path)
        if img is None:
            raise IOError("cannot read image %s" % file_path)
        return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)

    @staticmethod
    def _save(results, save_root_dir):
        """proc_callback of tile.py:243-288."""
        base_name, src_image, inst_map_dict, inst_info_dict, type_map_dict, pclass_map = results
        os.makedirs("%s/overlay/" % save_root_dir, exist_ok=True)
        src2 = cv2.resize(src_image, (0, 0), fx=2, fy=2, interpolation=cv2.INTER_NEAREST)
        overlay = cv2.cvtColor(_overlay(src2, inst_info_dict), cv2.COLOR_BGR2RGB)
        cv2.imwrite("%s/overlay/%s.jpg" % (save_root_dir, base_name), overlay)
        for tissue_code, pred_inst in inst_map_dict.items():
            type_pred = []
            inst_id = list(inst_info_dict[tissue_code].keys())
            for pred_dict in inst_info_dict[tissue_code].values():
                type_pred.append(pred_dict["type"] if "type" in pred_dict else -1)
            type_map = type_map_dict[tissue_code]
            os.makedirs("%s/%s_mat/" % (save_root_dir, tissue_code.lower()), exist_ok=True)
            mat_dict = {"inst_map": pred_inst, "type": type_pred, "id": inst_id}
            if type_map is not None:
                mat_dict["type_map"] = type_map
            sio.savemat("%s/%s_mat/%s.mat" % (save_root_dir, tissue_code.lower(), base_name), mat_dict)
        if pclass_map is not None:
            os.makedirs("%s/pclass_mat/" % save_root_dir, exist_ok=True)
            sio.savemat("%s/pclass_mat/%s.mat" % (save_root_dir, base_name), {"pclass": pclass_map})
