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

    def process_image(self, img, name="image"):
        """One RGB uint8 image -> the tuple _post_process_patches returns."""
        PostProcInstErodedContourMap.bind(self.engine.ctx)
        info_list, src_pos, _ = patch_grid(img.shape[0], img.shape[1], self.patch_input_shape,
                                           self.patch_output_shape, self.patch_output_overlap)
        # infer every distinct patch once (the reference's appended duplicate grid averages a
        # patch with itself, which is exact)
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
        same outputs."""
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
        # The reference decodes images in DataLoader worker processes (--nr_inference_workers) and
        # post-processes / saves in a ProcessPoolExecutor (--nr_post_proc_workers). Here every GPU
        # call stays on this thread (a cerb_ctx is single-threaded); the same two flags size a
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
        try:
            loads = {}
            for i, file_path in enumerate(file_path_list):
                if loaders is not None:
                    for j in range(i, min(i + ahead, len(file_path_list))):
                        if j not in loads:
                            loads[j] = loaders.submit(self._load_rgb, file_path_list[j])
                    img = loads.pop(i).result()
                else:
                    img = self._load_rgb(file_path)
                results = self.process_image(img, pathlib.Path(file_path).stem)
                if savers is not None:
                    pending_saves.append((file_path, savers.submit(self._save, results, self.output_dir)))
                    while len(pending_saves) > 4 * n_save:  # bound the results held in memory
                        done_path, fut = pending_saves.pop(0)
                        fut.result()
                        print("Done Assembling %s" % done_path)
                else:
                    self._save(results, self.output_dir)
                    print("Done Assembling %s" % file_path)
            for done_path, fut in pending_saves:  # a failed write raises here: no silent crash
                fut.result()
                print("Done Assembling %s" % done_path)
        finally:
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
