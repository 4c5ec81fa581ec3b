"""WSI mode: mirror of infer/wsi.py (InferManager.process_wsi_list / process_single_file) on the
B200-native engine (BASELINE config 4).

Same run_args, same skip-if-done rule, same outputs (`<out>/dat/<name>.dat` joblib dict with the
Nuclei / Gland / Lumen instance tables, `<out>/tissue/<name>.mat` Patch-Class map), different
plumbing:

  reference                                             here
  ----------------------------------------------------  --------------------------------------------
  tiatoolbox WSIReader + 12 loader processes            ArraySlide (wsi_reader.py); the slide lives in
  (infer/wsi.py:888-942)                                HBM, patches are cut by cerb_extract_patches
                                                        (zero padded) straight into the network input
  6 float32 memmaps + 6 count memmaps on disk           ONE device canvas [H, W, 9] float32
  (merge_prediction, :550-556, :603-621)                (cerb_scatter_patches)
  nn.DataParallel over the visible GPUs                 one process per GPU: batches are strided over
                                                        ranks, canvases summed with one NCCL
                                                        all-reduce (every pixel has one writer)
  nuclei post-proc in 6 worker processes (:643-683)     cerb_postproc_nuclei on canvas crops
  gland / lumen: cv2 resize + post_process (:720-800)   cerb_region_half + cerb_postproc_gland_lumen

tiatoolbox / shapely are not available offline; their placement and de-duplication rules are
restated in wsi_geometry.py (parity unpinned, see there). Per-instance contours / moments stay on
the host with OpenCV exactly as in the reference (SURVEY.md 8f-2 is the device version).

Known limits of this round: array-backed slides only; no resampling between scan and
processing resolution; the nuclei watershed of a 4032^2 post-processing tile runs on the
component-parallel watershed but falls back to the whole-tile emulation (one warp, seconds)
when two markers of one touching-nuclei component tie.
"""
import logging
import os
import pathlib
import pickle
import time
import uuid
from datetime import datetime

import cv2
import numpy as np
import torch

from .. import _lib
from ..instinfo import _rows as instinfo_rows
from ..instinfo import get_inst_info_dict, inst_table, tiatoolbox_dicts
from . import base
from .wsi_geometry import (boxes_intersect, filter_coordinates, get_coordinates, get_tile_info,
                           select_tile_instances)
from .wsi_reader import ArraySlide

# target key gen code : post proc class (infer/wsi.py:49-54); only the contour variants are built
_SUPPORTED_POSTPROC = ("IP-ERODED-CONTOUR-3", "IP-ERODED-CONTOUR-11")

HEAD_NAMES = ["Nuclei-INST", "Nuclei-TYPE", "Gland-INST", "Gland-TYPE", "Lumen-INST", "Patch-Class"]


def _cv_round(v):
    """cv::saturate_cast<int>(double): round half to even (cv2.resize's dsize from fx / fy)."""
    return int(np.rint(v))


def tiatoolbox_bounding_box(img):
    """tiatoolbox.utils.misc.get_bounding_box: [start_x, start_y, end_x, end_y], end exclusive."""
    rows = np.any(img, axis=1)
    cols = np.any(img, axis=0)
    rmin, rmax = np.where(rows)[0][[0, -1]]
    cmin, cmax = np.where(cols)[0][[0, -1]]
    return np.array([cmin, rmin, cmax + 1, rmax + 1])


def _ptr(t):
    return _lib.ctypes.c_void_p(t.data_ptr())


_NP_RECONSTRUCT = np.empty(0).__reduce__()[0]  # the callable numpy's own pickles name


def _reduce_ndarray(a):
    """What ndarray.__reduce__ returns for a C-ordered array, built without numpy's per-call
    overhead (3.5 us per array; a slide's table holds 1.6 million small arrays). The stream and
    the arrays read back (writable, owning their data) are the ones a plain pickle gives."""
    if a.dtype.hasobject or type(a) is not np.ndarray:
        return a.__reduce_ex__(pickle.HIGHEST_PROTOCOL)
    return (_NP_RECONSTRUCT, (np.ndarray, (0,), b"b"), (1, a.shape, a.dtype, False, a.tobytes()))


def dump_dat(obj, path):
    """The `.dat` instance table (infer/wsi.py:853 uses joblib.dump): a protocol-5 pickle, which
    joblib.load / pickle.load read back identically."""
    with open(path, "wb") as fh:
        pk = pickle.Pickler(fh, protocol=pickle.HIGHEST_PROTOCOL)
        pk.dispatch_table = {np.ndarray: _reduce_ndarray}
        pk.dump(obj)


def _unique_ids(n):
    """n random 128-bit hex keys (the reference draws uuid.uuid4().hex per instance, infer/wsi.py:265;
    one os.urandom call instead of half a million uuid objects)."""
    raw = os.urandom(16 * n).hex()
    return [raw[32 * i:32 * i + 32] for i in range(n)]


class InferManager(base.InferManager):
    # ------------------------------------------------------------------ helpers
    def _parse_args(self, run_args):
        """infer/wsi.py:444-452."""
        for variable, value in run_args.items():
            self.__setattr__(variable, value)
        self.chunk_shape = [self.chunk_shape, self.chunk_shape]
        self.tile_shape = [self.tile_shape, self.tile_shape]
        self.patch_input_shape = [self.patch_input_shape, self.patch_input_shape]
        self.patch_output_shape = [self.patch_output_shape, self.patch_output_shape]

    def _dist(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and not getattr(self, "force_single", False):
            return dist, dist.get_rank(), dist.get_world_size()
        return None, 0, 1

    # ------------------------------------------------------------------ inference
    @staticmethod
    def _row_bands(patch_outputs, world):
        """Splits the (row-major) patch list into `world` contiguous groups of whole patch ROWS,
        balanced by patch count: [(p_lo, p_hi)] per rank. A rank's patches then cover a
        horizontal band of the slide that no other rank writes."""
        n = len(patch_outputs)
        ys = patch_outputs[:, 1]
        row_start = np.concatenate([[0], np.nonzero(np.diff(ys))[0] + 1, [n]])  # first patch of each row
        bands, lo = [], 0
        for r in range(world):
            target = n * (r + 1) / world
            k = int(np.argmin(np.abs(row_start - target)))  # nearest row boundary
            hi = max(int(row_start[k]), lo) if r < world - 1 else n
            bands.append((lo, hi))
            lo = hi
        return bands

    def _infer_slide(self, slide, patch_inputs, patch_outputs):
        """Raw prediction of every selected patch, merged into one device canvas [H, W, C]
        (infer/wsi.py:585-621). With P ranks the patch rows are split into P contiguous bands
        (balanced by patch count); a rank uploads only the slide rows its band reads, infers its
        patches, and the bands are then exchanged by one NCCL broadcast per band - no arithmetic
        and no pixel travels twice (each canvas pixel has exactly one writer)."""
        eng, ctx, lib = self.engine, self.engine.ctx, self.engine.ctx.lib
        dist, rank, world = self._dist()
        H, W = slide.img.shape[:2]
        C = eng.model.canvas_c
        pin, pout = self.patch_input_shape[0], self.patch_output_shape[0]
        B = int(self.batch_size)
        dev = torch.device("cuda", ctx.device)
        plan = eng.plan_for(B, pin, pin, pout, pout)
        bands = self._row_bands(patch_outputs, world)
        p_lo, p_hi = bands[rank]
        canvas = torch.zeros((H, W, C), dtype=torch.float32, device=dev)
        if p_hi > p_lo:
            # slide rows read by this band (zero padding applies only outside the slide itself)
            r0 = max(int(patch_inputs[p_lo:p_hi, 1].min()), 0)
            r1 = min(int(patch_inputs[p_lo:p_hi, 3].max()), H)
            slide_dev = torch.from_numpy(np.array(slide.img[r0:r1], dtype=np.uint8, order="C")).to(dev)
            patch_buf = torch.empty((B, pin, pin, 3), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize(dev)
            far = -4 * pin  # padding entries of the last batch read only zeros
            for start in range(p_lo, p_hi, B):
                k = min(B, p_hi - start)
                tl_in = np.full((B, 2), far, dtype=np.int32)
                tl_in[:k, 0] = patch_inputs[start:start + k, 1] - r0
                tl_in[:k, 1] = patch_inputs[start:start + k, 0]
                _lib.check(lib.cerb_extract_patches(ctx.handle, _ptr(slide_dev), r1 - r0, W, 0, 0,
                                                    tl_in.ctypes.data_as(_lib.ctypes.c_void_p), B, pin,
                                                    pin, _ptr(patch_buf), 1 | 2 | 4),
                           "cerb_extract_patches")
                plan.run(device_ptr=patch_buf.data_ptr())
                tl_out = np.ascontiguousarray(patch_outputs[start:start + k][:, [1, 0]], dtype=np.int32)
                _lib.check(lib.cerb_scatter_patches(
                    ctx.handle, _lib.ctypes.c_void_p(plan.tensor_ptr(plan.spec.canvas)), k, pout, pout,
                    C, tl_out.ctypes.data_as(_lib.ctypes.c_void_p), _ptr(canvas), H, W),
                    "cerb_scatter_patches")
                self.nr_patches_done += k
            # extract / forward / scatter are queued asynchronously on the ctx stream (the loop above
            # never waits for the GPU); torch must not touch or free these buffers before it is idle
            _lib.check(lib.cerb_ctx_sync(ctx.handle), "cerb_ctx_sync")
            del slide_dev, patch_buf
        if world > 1:
            t0 = time.perf_counter()
            for r, (lo, hi) in enumerate(bands):
                if hi <= lo:
                    continue
                y0 = max(int(patch_outputs[lo:hi, 1].min()), 0)
                y1 = min(int(patch_outputs[lo:hi, 3].max()), H)
                dist.broadcast(canvas[y0:y1], src=r)  # rows are contiguous in [H, W, C]
            torch.cuda.synchronize(dev)
            self.t_exchange = time.perf_counter() - t0
        return canvas

    # ------------------------------------------------------------------ nuclei
    def _process_tile_predictions(self, canvas, tile_bounds, tile_flag, tile_mode, ref_boxes, margin):
        """infer/wsi.py:64-268 for one post-processing tile, on the device canvas. ref_boxes: [n,4]
        boxes of the instances accumulated so far (cross tiles replace some of them). Returns the
        tile's surviving instances as COLUMNS (box, centroid, contour offsets, contour points,
        prob, type - slide coordinates; see dat_writer.InstanceStore) or None, and the INDICES into
        ref_boxes it replaces. (Both halves on the main context, one after the other.)"""
        labelled = self._tile_labels(canvas, tile_bounds)
        return self._tile_tables(self.engine.ctx, labelled, tile_bounds, tile_flag, tile_mode,
                                 ref_boxes, margin)

    def _tile_labels(self, canvas, tile_bounds):
        """First half of a post-processing tile (infer/wsi.py:64-135): nuclei label map of the tile
        crop, left in HBM. None when the tile is empty."""
        eng, ctx, lib = self.engine, self.engine.ctx, self.engine.ctx.lib
        idx = eng.model.idx_dict
        H, W, C = canvas.shape
        tile_bounds = np.asarray(tile_bounds, dtype=np.int64)
        tile_tl = tile_bounds[:2]
        x0, y0 = int(tile_bounds[0]), int(tile_bounds[1])
        x1, y1 = min(int(tile_bounds[2]), W), min(int(tile_bounds[3]), H)  # numpy slicing clips
        if x1 <= x0 or y1 <= y0:
            return None
        crop = canvas[y0:y1, x0:x1].contiguous()
        h, w = crop.shape[:2]
        # the label map of the tile stays in HBM: the device instance table is all the host needs
        type_map_present = "Nuclei-TYPE" in idx
        type_dev = crop[..., idx["Nuclei-TYPE"][0]].contiguous() if type_map_present else None
        labels = torch.empty((h, w), dtype=torch.int32, device=canvas.device)
        any_fg = torch.zeros(1, dtype=torch.int32, device=canvas.device)
        # torch's stream only (crop / type plane are ready); a device-wide synchronisation would
        # also wait for the table thread's kernels of the previous tile
        torch.cuda.current_stream(canvas.device).synchronize()
        t0 = time.perf_counter()
        _lib.check(lib.cerb_postproc_nuclei(ctx.handle, _ptr(crop), 1, h, w, C,
                                            idx["Nuclei-INST"][0], _ptr(labels), _ptr(any_fg), 1 | 2),
                   "cerb_postproc_nuclei")
        _lib.check(lib.cerb_ctx_sync(ctx.handle), "cerb_ctx_sync")
        self.t_dev += time.perf_counter() - t0
        del crop
        if not int(any_fg.item()):
            return None
        return labels, type_dev, (h, w), type_map_present

    def _tile_tables(self, ctx, labelled, tile_bounds, tile_flag, tile_mode, ref_boxes, margin):
        """Second half of a post-processing tile (infer/wsi.py:137-268): instance table of the label
        map (device pass on `ctx` - a context of its own, so that it overlaps the watershed of the
        next tile on the main context) and the boundary de-duplication."""
        if labelled is None:
            return None, []
        labels, type_dev, (h, w), type_map_present = labelled
        tile_bounds = np.asarray(tile_bounds, dtype=np.int64)
        tile_tl = tile_bounds[:2]
        t0 = time.perf_counter()
        table = inst_table(ctx, labels.data_ptr(), type_dev.data_ptr() if type_dev is not None else None,
                           on_device=True, shape=(h, w))
        self.t_table += time.perf_counter() - t0  # device pass + read-back; part of t_host
        del labels, type_dev, labelled
        rows = np.asarray(instinfo_rows(table), dtype=np.int64)
        if len(rows) == 0:
            self.t_host += time.perf_counter() - t0
            return None, []
        inst_boxes = table.box[rows][:, [1, 0, 3, 2]]  # tile coordinates, [x0, y0, x1, y1]
        sel, sel_ref = select_tile_instances(inst_boxes, tile_bounds, tile_flag, tile_mode, margin,
                                             ref_boxes if tile_mode == 3 else None)
        keep = np.ones(len(rows), dtype=bool)
        keep[np.asarray(sel, dtype=np.int64)] = False
        cols = self._table_columns(table, rows[keep], tile_tl, type_map_present)
        self.t_host += time.perf_counter() - t0
        return cols, sel_ref

    @staticmethod
    def _table_columns(table, rows, offset_xy, has_type):
        """Rows of a device table as columns in slide coordinates - the arithmetic of
        instinfo.tiatoolbox_dicts ((m10 / m00 + box origin) + offset in float64) without creating a
        Python object per instance."""
        rows = np.asarray(rows, dtype=np.int64)
        off = np.asarray(offset_xy, dtype=np.int64)
        box = table.box[rows][:, [1, 0, 3, 2]].astype(np.int64)          # x0, y0, x1, y1
        m = table.moments[rows].astype(np.float64)
        cen = np.stack([m[:, 1] / m[:, 0], m[:, 2] / m[:, 0]], axis=1)
        cen = cen + box[:, :2]
        if off.any():
            cen = cen + off
            box = box + np.concatenate([off, off])
        starts, stops = table.contour_off[rows], table.contour_off[rows + 1]
        lens = stops - starts
        coff = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        src = np.repeat(starts - coff[:-1], lens) + np.arange(int(coff[-1]), dtype=np.int64)
        xy = table.contour_xy[src].astype(np.int64) + off
        if has_type:
            typ = (table.type[rows, 0] / 4.0).astype(np.int64)  # int(np.float) truncates
            prob = table.type[rows, 1] / (table.moments[rows, 0] + 1.0e-6)
        else:
            typ = prob = None
        return (box, cen, coff, xy, prob, typ)

    def _postproc_nuclei(self, canvas, patch_outputs, pp_tile_shape, margin):
        """infer/wsi.py:640-686. With several ranks every rank holds the merged canvas; the tiles
        of a set are strided over ranks (the reference's ProcessPoolExecutor); the per-tile tables
        travel as a handful of arrays and rank 0 appends them in the reference's tile order.
        Returns a dat_writer.InstanceStore (rank 0) / None."""
        from .dat_writer import InstanceStore
        dist, rank, world = self._dist()
        H, W, _ = canvas.shape
        tile_sets = get_tile_info((W, H), pp_tile_shape, self.patch_output_shape, margin)
        store = InstanceStore(has_type="Nuclei-TYPE" in self.engine.model.idx_dict)
        self.t_dev = self.t_host = self.t_table = 0.0
        from concurrent.futures import ThreadPoolExecutor
        table_ctx = self._table_ctx()
        with ThreadPoolExecutor(max_workers=1) as table_pool:
            self._nuclei_tile_sets(tile_sets, canvas, patch_outputs, margin, store, table_pool, table_ctx)
        return store if rank == 0 else None

    def _table_ctx(self):
        """Context (own stream and instance-table workspace) of the table thread; kept for the life
        of the manager."""
        ctx = self.__dict__.get("_tbl_ctx")
        if ctx is None:
            from ..engine import Context
            ctx = self._tbl_ctx = Context(self.engine.ctx.device, self.engine.ctx.precision)
        return ctx

    def _nuclei_tile_sets(self, tile_sets, canvas, patch_outputs, margin, store, table_pool, table_ctx):
        dist, rank, world = self._dist()
        for set_idx, (set_bounds, set_flags) in enumerate(tile_sets):
            todo = [i for i, tb in enumerate(set_bounds) if len(boxes_intersect(patch_outputs, tb)) > 0]
            # cross tiles (set 3) replace accumulated instances: every tile of the set sees the boxes
            # accumulated before the set started (the reference submits the whole set at once)
            ref_boxes = None
            if set_idx == 3:
                if rank == 0:
                    ref_boxes = store.boxes()
                if world > 1:
                    obj = [ref_boxes]
                    dist.broadcast_object_list(obj, src=0)
                    ref_boxes = obj[0]
            # two-stage pipeline over the tiles of this rank: the watershed of tile k + 1 (main
            # context) runs while a worker thread builds the instance table of tile k on a context
            # of its own; at most two label maps wait in HBM
            futs = []
            for i in todo[rank::world]:
                if len(futs) >= 2:
                    futs[-2][1].result()
                labelled = self._tile_labels(canvas, set_bounds[i])
                futs.append((i, table_pool.submit(self._tile_tables, table_ctx, labelled, set_bounds[i],
                                                  set_flags[i], set_idx, ref_boxes, margin)))
                del labelled
            local = [(i, f.result()) for i, f in futs]
            del futs
            if world > 1:
                gathered = [None] * world if rank == 0 else None
                dist.gather_object(local, gathered, dst=0)
                if rank == 0:
                    local = sorted((x for part in gathered for x in part), key=lambda x: x[0])
            if rank == 0:
                for _, (cols, remove_idx_list) in local:
                    if cols is not None:
                        store.append(*cols)
                    store.remove(remove_idx_list)

    # ------------------------------------------------------------------ gland / lumen
    def _postproc_gland_lumen(self, canvas, wsi_mask, mask_downsample_ratio):
        """infer/wsi.py:720-840. Returns {"Gland": {...}, "Lumen": {...}} (keys present only when
        an instance was found, as in the reference)."""
        from scipy import ndimage
        eng, ctx, lib = self.engine, self.engine.ctx, self.engine.ctx.lib
        idx = eng.model.idx_dict
        H, W, C = canvas.shape
        dev = canvas.device
        mask_lab = ndimage.label(wsi_mask)[0]
        ids = np.unique(mask_lab).tolist()
        tissue_info_list = []
        if len(ids) > 1:
            for region_id in ids[1:]:
                reg = mask_lab == region_id
                rows, cols = np.any(reg, axis=1), np.any(reg, axis=0)
                rmin, rmax = np.where(rows)[0][[0, -1]]
                cmin, cmax = np.where(cols)[0][[0, -1]]
                tissue_info_list.append([rmin, rmax + 1, cmin, cmax + 1])  # misc/utils.py:82-91
        else:
            tissue_info_list.append([0, mask_lab.shape[0], 0, mask_lab.shape[1]])
        dist, rank, world = self._dist()
        out = {}
        per_region = []  # (region index, tissue, instance info) of the regions this rank handles
        ds_factor = 0.5
        for ridx, ti in enumerate(tissue_info_list):
            if ridx % world != rank:
                continue
            rmin = int(round(ti[0] / mask_downsample_ratio))
            rmax = int(round(ti[1] / mask_downsample_ratio))
            cmin = int(round(ti[2] / mask_downsample_ratio))
            cmax = int(round(ti[3] / mask_downsample_ratio))
            rmax, cmax = min(rmax, H), min(cmax, W)  # numpy slicing of the memmap clips
            h, w = rmax - rmin, cmax - cmin
            if h <= 0 or w <= 0:
                continue
            tissue_topleft = [cmin, rmin]
            mask_idx = (mask_lab[ti[0]:ti[1], ti[2]:ti[3]] == ridx + 1)
            if h != mask_idx.shape[0] and w != mask_idx.shape[1]:  # `and`: as the reference (:768)
                mask_idx = cv2.resize(mask_idx.astype("uint8"), (w, h), interpolation=cv2.INTER_NEAREST)
            mask_u8 = np.ascontiguousarray(mask_idx, dtype=np.uint8)
            if mask_u8.shape != (h, w):
                raise ValueError("tissue mask segment %r does not match the prediction window %r "
                                 "(the reference fails here too)" % (mask_u8.shape, (h, w)))
            oh, ow = _cv_round(h * ds_factor), _cv_round(w * ds_factor)
            if oh <= 0 or ow <= 0:
                continue
            inst_maps, type_maps = {}, {}
            for tissue in ("Gland", "Lumen"):
                code = self.decoder_dict[tissue + "-INST"]
                if code not in _SUPPORTED_POSTPROC:
                    raise NotImplementedError("post-processing %r (PostProcInstErodedMap) is outside "
                                              "the hot path (SURVEY.md 8f-4)" % code)
                chans = list(range(*idx[tissue + "-INST"]))
                has_type = (tissue + "-TYPE") in HEAD_NAMES and (tissue + "-TYPE") in idx
                if has_type:
                    chans += list(range(*idx[tissue + "-TYPE"]))
                k = len(chans)
                half = torch.empty((oh, ow, k), dtype=torch.float32, device=dev)
                ch_arr = np.asarray(chans, dtype=np.int32)
                torch.cuda.synchronize(dev)
                _lib.check(lib.cerb_region_half(
                    ctx.handle, _ptr(canvas), H, W, C, rmin, cmin, h, w,
                    mask_u8.ctypes.data_as(_lib.ctypes.c_void_p),
                    ch_arr.ctypes.data_as(_lib.ctypes.c_void_p), k, _ptr(half), oh, ow),
                    "cerb_region_half")
                labels = torch.empty((oh, ow), dtype=torch.int32, device=dev)
                _lib.check(lib.cerb_postproc_gland_lumen(
                    ctx.handle, _ptr(half), 1, oh, ow, k, 0, 0 if tissue == "Gland" else 1,
                    float(ds_factor), _ptr(labels), 1 | 2), "cerb_postproc_gland_lumen")
                # the label maps stay in HBM; only the instance tables reach the host (the
                # reference keeps float64 maps, loader/postproc.py:290,331, hence float64 keys)
                inst_maps[tissue] = labels
                torch.cuda.synchronize(dev)  # the ctx stream is done with `half` before torch reads it
                type_maps[tissue] = half[..., 2].contiguous() if has_type else None
                del half
            torch.cuda.synchronize(dev)
            # remove lumen predictions not inside glands (:802-807)
            _lib.check(lib.cerb_mask_lumen(ctx.handle, _ptr(inst_maps["Lumen"]), _ptr(inst_maps["Gland"]),
                                           oh * ow), "cerb_mask_lumen")
            for tissue in ("Gland", "Lumen"):
                tm = type_maps[tissue]
                pred_inst_info = get_inst_info_dict(inst_maps[tissue].data_ptr(),
                                                    tm.data_ptr() if tm is not None else None, ds_factor,
                                                    ctx=ctx, on_device=True, shape=(oh, ow),
                                                    key_dtype=np.float64)
                for inst_id, inst_info in pred_inst_info.items():
                    # Reference quirk kept for drop-in parity (:815-829): `box` is [[r0,c0],[r1,c1]]
                    # but the (x, y) top-left is added to it, i.e. x to the rows and y to the columns.
                    inst_info["box"] = inst_info["box"] + tissue_topleft
                    inst_info["contour"] = inst_info["contour"] + tissue_topleft
                    inst_info["centroid"] = inst_info["centroid"] + tissue_topleft
                    b = inst_info["box"]
                    inst_info["box"] = np.array([b[0][1], b[0][0], b[1][1], b[1][0]])
                    per_region.append((ridx, tissue, inst_info))
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, per_region)
            per_region = sorted((x for part in gathered for x in part), key=lambda x: x[0])
        for _, tissue, inst_info in per_region:  # region order, Gland before Lumen inside a region
            out.setdefault(tissue, {})[uuid.uuid4().hex] = inst_info
        return out

    # ------------------------------------------------------------------ one slide
    def process_single_file(self, wsi_idx, wsi_basename, output_dir):
        """infer/wsi.py:502-857."""
        eng = self.engine
        wsi_path = self.imgs[wsi_idx]
        mask_path = self.masks[wsi_idx]
        dist, rank, world = self._dist()
        start = time.perf_counter()
        slide = ArraySlide.open(wsi_path, self.wsi_proc_mag)
        if getattr(self, "warm_crop", None):  # tools/wsi_bench.py: untimed warm-up on a corner
            slide.img = slide.img[:self.warm_crop, :self.warm_crop]
        self.wsi_proc_shape = slide.slide_dimensions(self.wsi_proc_mag)[::-1]  # YX
        self.wsi_base_mag = slide.mpp
        self.wsi_base_shape = np.array(slide.img.shape[:2])
        H, W = int(self.wsi_proc_shape[0]), int(self.wsi_proc_shape[1])
        if mask_path is not None and os.path.isfile(mask_path):
            wsi_mask = cv2.imread(mask_path)
            wsi_mask = cv2.cvtColor(wsi_mask, cv2.COLOR_BGR2GRAY)
            wsi_mask[wsi_mask > 0] = 1
        else:
            wsi_mask = np.ones((H, W), dtype=np.uint8)
        mask_downsample_ratio = wsi_mask.shape[0] / H
        if rank == 0 and self.save_mask:  # the reference crashes here (undefined self.wsi_mask)
            cv2.imwrite("%s/mask/%s.png" % (self.output_dir, wsi_basename), wsi_mask * 255)
        if rank == 0 and self.save_thumb:
            cv2.imwrite("%s/thumb/%s.png" % (self.output_dir, wsi_basename),
                        cv2.cvtColor(slide.thumbnail(1.25), cv2.COLOR_RGB2BGR))

        pin, pout = self.patch_input_shape, self.patch_output_shape
        patch_inputs, patch_outputs = get_coordinates((W, H), pin, pout, pout)
        sel = filter_coordinates(wsi_mask, patch_outputs, (H, W))
        patch_inputs, patch_outputs = patch_inputs[sel], patch_outputs[sel]
        self.logger.info("Preparing Input Output Placement: %s" % (time.perf_counter() - start))

        start = time.perf_counter()
        self.nr_patches_done = 0
        canvas = self._infer_slide(slide, patch_inputs, patch_outputs)
        self.logger.info("Inference Time: %s (%d patches on this rank, %d selected)" % (
            time.perf_counter() - start, self.nr_patches_done, len(patch_inputs)))
        self.last_canvas = canvas if getattr(self, "keep_canvas", False) else None
        if os.environ.get("CERB_WSI_CANVAS_SHA"):  # debugging aid: is the merged canvas reproducible?
            import hashlib
            hsh = hashlib.sha1()
            for y in range(0, H, 512):
                hsh.update(canvas[y:y + 512].cpu().numpy().tobytes())
            self.logger.info("Canvas sha1: %s" % hsh.hexdigest())
            print("rank %d canvas sha1 %s" % (rank, hsh.hexdigest()), flush=True)
        wsi_inst_info = {}
        start = time.perf_counter()
        # hard-coded in both IOSegmentorConfigs of the reference (infer/wsi.py:898,909), which
        # ignores --ambiguous_size; `_test_margin` is a test hook only
        margin = int(getattr(self, "_test_margin", 64))
        nuclei_store = self._postproc_nuclei(canvas, patch_outputs, self.postproc_tile_shape, margin)
        wsi_inst_info["Nuclei"] = nuclei_store
        self.logger.info("Nuclei Post Proc Time: %s" % (time.perf_counter() - start))
        lib_, h_ = eng.ctx.lib, eng.ctx.handle
        self.logger.info("Nuclei watershed: %d large tiles, %d redone by the exact whole-tile emulation "
                         "(marker ties); device %.2f s, host instance info %.2f s (of which %.2f s device table pass; "
                         "the table thread overlaps the watershed of the next tile)" % (
                             lib_.cerb_ctx_stat(h_, b"ws_large_images"),
                             lib_.cerb_ctx_stat(h_, b"ws_large_fallbacks"), self.t_dev, self.t_host,
                             self.t_table))

        start = time.perf_counter()
        idx = eng.model.idx_dict
        # every rank holds the merged canvas: the tissue map is written by the LAST rank while rank 0
        # is busy with the instance tables
        if rank == world - 1 and "Patch-Class" in self.model_args["decoder_kwargs"].keys() and "Patch-Class" in idx:
            import scipy.io as sio
            ds = 0.25
            ph, pw = _cv_round(H * ds), _cv_round(W * ds)
            pclass_map = np.empty((ph, pw), dtype=np.float32)
            ctx = eng.ctx
            _lib.check(ctx.lib.cerb_nearest_channel(
                ctx.handle, _ptr(canvas), H, W, canvas.shape[2], idx["Patch-Class"][0], ds,
                pclass_map.ctypes.data_as(_lib.ctypes.c_void_p), ph, pw), "cerb_nearest_channel")
            lores = cv2.resize(wsi_mask, (pw, ph), interpolation=cv2.INTER_NEAREST)
            pclass_map *= lores
            sio.savemat("%s/tissue/%s.mat" % (output_dir, wsi_basename), {"pclass": pclass_map})
        self.logger.info("Tissue Region Post Proc Time: %s" % (time.perf_counter() - start))

        start = time.perf_counter()
        wsi_inst_info.update(self._postproc_gland_lumen(canvas, wsi_mask, mask_downsample_ratio))
        if rank != 0:
            del canvas
            return None
        wsi_inst_info["proc_resolution"] = {"resolution": self.wsi_proc_mag, "units": "mpp"}
        wsi_inst_info["base_resolution"] = {"resolution": self.wsi_base_mag, "units": "mpp"}
        wsi_inst_info["proc_dimensions"] = self.wsi_proc_shape
        wsi_inst_info["base_dimensions"] = self.wsi_base_shape
        # infer/wsi.py:853 writes this dict with joblib.dump, whose per-array framing costs seconds
        # for tens of thousands of small arrays; a protocol-5 pickle is what joblib.load reads back
        # identically (tests/test_gpu_wsi.py loads it with joblib) at a fraction of the time.
        t_out = time.perf_counter()
        from .dat_writer import InstanceStore, write_dat
        write_dat(wsi_inst_info, "%s/dat/%s.dat" % (output_dir, wsi_basename))
        self.n_nuclei = int(nuclei_store.alive().sum())
        if getattr(self, "return_inst_dicts", True):
            # API convenience (tests, notebooks): the reference's dict of dicts; the CLI turns it off -
            # half a million Python dicts per slide are the serial tail this module is built to avoid
            wsi_inst_info = {k: (v.to_dict() if isinstance(v, InstanceStore) else v)
                             for k, v in wsi_inst_info.items()}
        # part of the reference's "Gland & Lumen Post Proc Time" (:853-856); logged on its own too
        self.logger.info("Output File Time: %s" % (time.perf_counter() - t_out))
        self.logger.info("Gland & Lumen Post Proc Time: %s" % (time.perf_counter() - start))
        del canvas
        return wsi_inst_info

    # ------------------------------------------------------------------ driver
    def process_wsi_list(self, run_args):
        """infer/wsi.py:860-986. The reference parses --tile_shape / --chunk_shape / the patch
        shapes and then hard-codes 15000 / 4096 / 448 / 144; the same
        constants are used here unless run_args carries `infer_tile_shape` /
        `postproc_tile_shape` (test hooks)."""
        self._parse_args(run_args)
        for k, code in self.decoder_dict.items():
            if k.endswith("-INST") and code not in _SUPPORTED_POSTPROC + ("IP-ERODED-3", "IP-ERODED-11"):
                raise KeyError(code)
        dist, rank, world = self._dist()
        for sub in ("/dat/", "/tissue/") + (("/thumb/",) if self.save_thumb else ()) + (
                ("/mask/",) if self.save_mask else ()):
            os.makedirs(self.output_dir + sub, exist_ok=True)
        os.makedirs(self.logging_dir, exist_ok=True)
        # hard-coded in the reference (infer/wsi.py:885-915)
        self.patch_input_shape = [448, 448] if not getattr(self, "honour_patch_shapes", False) \
            else self.patch_input_shape
        self.patch_output_shape = [144, 144] if not getattr(self, "honour_patch_shapes", False) \
            else self.patch_output_shape
        # run_args may carry a scalar `postproc_tile_shape` (test hook; the reference hard-codes
        # 4096): keep the scalar, derive the list, so that a second call on this manager works
        self._pp_tile = int(run_args.get("postproc_tile_shape", getattr(self, "_pp_tile", 4096)))
        self.postproc_tile_shape = [self._pp_tile] * 2
        self.imgs = self.input_list
        self.masks = self.mask_list
        from ..postproc import PostProcInstErodedContourMap
        PostProcInstErodedContourMap.bind(self.engine.ctx)
        results = {}
        for wsi_idx, wsi_path in enumerate(self.imgs):
            wsi_basename = pathlib.Path(wsi_path).stem
            start = time.perf_counter()
            dt_string = datetime.now().strftime("%d-%m-%Y_%H:%M:%S")
            self.logger = logging.getLogger("cerberus_b200.wsi.%d" % rank)
            fh = logging.FileHandler("%s/%s_%s_std%s.log" % (
                self.logging_dir, wsi_basename, dt_string, "" if rank == 0 else ".rank%d" % rank), mode="w")
            fh.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
            self.logger.addHandler(fh)
            self.logger.setLevel(logging.DEBUG)
            if not os.path.exists(self.output_dir + "/dat/%s.dat" % wsi_basename):
                self.logger.info("Processing %s ..." % wsi_basename)
                results[wsi_basename] = self.process_single_file(wsi_idx, wsi_basename,
                                                                 self.output_dir)
                self.logger.info("Overall Time: %s" % (time.perf_counter() - start))
                self.logger.info("Finish")
            else:
                self.logger.warning("Skip %s- already processed!" % wsi_basename)
            self.logger.handlers.clear()
            fh.close()
            if dist is not None:
                dist.barrier()
        return results
