"""Patch / tile placement and instance de-duplication of the WSI path.

The reference delegates these to tiatoolbox 1.3.1 and shapely 1.7.1 (`environment.yml:34`),
neither of which is vendored or installed. They are restated here from the published
algorithms (SURVEY.md Appendix C lists every call site in `infer/wsi.py`):

  get_coordinates     PatchExtractor.get_coordinates via NucleusInstanceSegmentor.get_coordinates
                      (infer/wsi.py:276, 400)
  filter_coordinates  SemanticSegmentor.filter_coordinates (infer/wsi.py:285, 567)
  get_tile_info       NucleusInstanceSegmentor._get_tile_info (infer/wsi.py:317, 579, 643)
  select_tile_instances / merge of `_process_tile_predictions` (infer/wsi.py:137-268)

PARITY UNPINNED: the originals cannot be imported offline, so these restatements are checked
only against hand-derived cases (tests/test_wsi_host.py). All boxes are [x0, y0, x1, y1]
(end exclusive), shapes are (width, height) unless stated, as in tiatoolbox.

shapely semantics used by the reference and reproduced with integer arithmetic:
  STRtree.query(g)   -> every geometry whose bounding box intersects g's bounding box, touching
                        included (shapely 1.7 returns the candidates without refining them);
  a.contains(b)      -> for boxes: b inside the closed box a (b non-degenerate).
"""
import numpy as np


def boxes_intersect(boxes, sel):
    """Indices of `boxes` whose closed bounding box intersects the closed box `sel`
    (STRtree.query semantics, shapely 1.7: envelope test, touching counts)."""
    boxes = np.asarray(boxes).reshape(-1, 4)
    x0, y0, x1, y1 = sel
    hit = (boxes[:, 0] <= x1) & (boxes[:, 2] >= x0) & (boxes[:, 1] <= y1) & (boxes[:, 3] >= y0)
    return np.nonzero(hit)[0]


def boxes_within(boxes, sel):
    """Indices of `boxes` lying inside the closed box `sel` (shapely `sel.contains(box)`)."""
    boxes = np.asarray(boxes).reshape(-1, 4)
    x0, y0, x1, y1 = sel
    hit = (boxes[:, 0] >= x0) & (boxes[:, 2] <= x1) & (boxes[:, 1] >= y0) & (boxes[:, 3] <= y1)
    return np.nonzero(hit)[0]


def get_coordinates(image_shape, patch_input_shape, patch_output_shape, stride_shape):
    """tiatoolbox PatchExtractor.get_coordinates with its default (no bound filtering).
    image_shape: (width, height). Returns (patch_inputs, patch_outputs) int64 [n,4]; the output
    grid starts at 0 with `stride_shape` steps up to ceil(size / out) * out, x fastest; the
    input box is the output box grown by (in - out) // 2 on the top/left (may be negative)."""
    image_shape = np.asarray(image_shape, dtype=np.int64)
    pin = np.asarray(patch_input_shape, dtype=np.int64)
    pout = np.asarray(patch_output_shape, dtype=np.int64)
    stride = np.asarray(stride_shape, dtype=np.int64)
    x_end = int(np.ceil(image_shape[0] / pout[0]) * pout[0])
    y_end = int(np.ceil(image_shape[1] / pout[1]) * pout[1])
    xs = np.arange(0, x_end, stride[0], dtype=np.int64)
    ys = np.arange(0, y_end, stride[1], dtype=np.int64)
    gx, gy = np.meshgrid(xs, ys)  # 'xy' indexing: x varies fastest
    out_tl = np.stack([gx.flatten(), gy.flatten()], axis=-1)
    out_br = out_tl + pout[None]
    in_tl = out_tl - ((pin - pout) // 2)[None]
    in_br = in_tl + pin[None]
    return np.concatenate([in_tl, in_br], -1), np.concatenate([out_tl, out_br], -1)


def filter_coordinates(mask, bounds, proc_shape_yx):
    """tiatoolbox SemanticSegmentor.filter_coordinates: keep a box iff the mask has a non-zero
    pixel inside ceil(scale * box), scale = mask rows / slide rows at the processing resolution
    (the y ratio is used for both axes, as upstream). numpy slicing semantics apply to boxes
    that leave the mask (negative starts wrap as in the original)."""
    mask = np.asarray(mask)
    scale = mask.shape[0] / float(proc_shape_yx[0])
    flags = np.zeros(len(bounds), dtype=bool)
    for i, b in enumerate(bounds):
        sx, sy, ex, ey = np.ceil(scale * np.asarray(b)).astype(np.int32)
        roi = mask[sy:ey, sx:ex]
        flags[i] = np.sum(roi > 0) > 0
    return flags


def get_tile_info(image_shape, tile_shape, patch_output_shape, margin):
    """NucleusInstanceSegmentor._get_tile_info. image_shape: (width, height). Returns a list of
    [boxes, flags] sets: 0 = the non-overlapping tile grid, 1 = vertical strips over vertical
    seams, 2 = horizontal strips over horizontal seams, 3 = squares at seam crossings. flags are
    [top, bottom, left, right] removal flags (1 = instances in the `margin` band along that side
    are dropped; 0 on slide borders). A slide that fits in one tile yields only set 0."""
    image_shape = np.asarray(image_shape, dtype=np.int64)
    pout = np.asarray(patch_output_shape, dtype=np.int64)
    tile_shape = (np.floor(np.asarray(tile_shape) / pout) * pout).astype(np.int64)
    _, boxes = get_coordinates(image_shape, tile_shape, tile_shape, tile_shape)
    if np.all(image_shape <= tile_shape):
        return [[boxes, np.zeros([boxes.shape[0], 4], dtype=np.int32)]]
    w, h = int(image_shape[0]), int(image_shape[1])

    def unset_removal_flag(bxs, flags):
        edges = [(0, 0, w, 0), (0, h, w, h), (0, 0, 0, h), (w, 0, w, h)]  # top bottom left right
        for idx, e in enumerate(edges):
            flags[boxes_intersect(bxs, e), idx] = 0
        return flags

    br = boxes[:, 2:]
    tr = np.stack([boxes[:, 2], boxes[:, 1]], -1)
    bl = np.stack([boxes[:, 0], boxes[:, 3]], -1)
    flags = unset_removal_flag(boxes, np.ones([boxes.shape[0], 4], dtype=np.int32))
    info = [[boxes, flags]]
    m = int(margin)
    # vertical strips: tiles whose right side is a seam
    sel = np.nonzero(flags[:, 3])[0]
    vb = np.concatenate([tr[sel] - np.array([m, 0])[None], br[sel] + np.array([m, 0])[None]], -1)
    vf = np.zeros([vb.shape[0], 4], dtype=np.int32)
    vf[:, [0, 1]] = 1
    info.append([vb, unset_removal_flag(vb, vf)])
    # horizontal strips: tiles whose bottom side is a seam
    sel = np.nonzero(flags[:, 1])[0]
    hb = np.concatenate([bl[sel] - np.array([0, m])[None], br[sel] + np.array([0, m])[None]], -1)
    hf = np.zeros([hb.shape[0], 4], dtype=np.int32)
    hf[:, [2, 3]] = 1
    info.append([hb, unset_removal_flag(hb, hf)])
    # squares where four tiles meet
    sel = np.nonzero(np.prod(flags[:, [1, 3]], axis=-1))[0]
    cb = np.concatenate([br[sel] - 2 * m, br[sel] + 2 * m], -1)
    info.append([cb, np.zeros([cb.shape[0], 4], dtype=np.int32)])
    return info


def select_tile_instances(inst_boxes, tile_bounds, tile_flag, tile_mode, margin, ref_boxes=None):
    """The selection rules of `_process_tile_predictions` (infer/wsi.py:153-262).

    inst_boxes: [n,4] boxes of the instances found in this tile, TILE coordinates.
    Returns (remove_in_tile, remove_in_ref): indices into inst_boxes of instances to drop from
    this tile's result, and indices into ref_boxes (WSI coordinates, the accumulated result) of
    instances the new tile replaces (tile_mode 3 only)."""
    tile_bounds = np.asarray(tile_bounds, dtype=np.int64)
    tile_tl, tile_br = tile_bounds[:2], tile_bounds[2:]
    w, h = (tile_br - tile_tl).tolist()
    m = int(margin)
    boundary_lines = [(0, 0, w, 1), (0, h - 1, w, h), (0, 0, 1, h), (w - 1, 0, w, h)]
    margin_boxes = [(0, 0, w, m), (0, h - m, w, h), (0, 0, m, h), (w - m, 0, w, h)]
    sel = []
    if tile_mode in (0, 3):
        # instances lying entirely inside a flagged margin band
        for idx, box in enumerate(margin_boxes):
            if tile_flag[idx] or tile_mode == 3:
                cand = boxes_intersect(inst_boxes, box)
                inside = boxes_within(np.asarray(inst_boxes).reshape(-1, 4)[cand], box)
                sel.extend(cand[inside].tolist())
    elif tile_mode in (1, 2):
        # everything touching a flagged margin band, or the boundary line of an unflagged side
        for idx, flag in enumerate(tile_flag):
            box = margin_boxes[idx] if flag else boundary_lines[idx]
            sel.extend(boxes_intersect(inst_boxes, box).tolist())
    else:
        raise ValueError("Unknown tile mode %r." % (tile_mode,))
    remove_in_ref = []
    if tile_mode == 3 and ref_boxes is not None and len(ref_boxes) > 0:
        lines = np.array([[[m, m], [w - m, m]], [[m, h - m], [w - m, h - m]],
                          [[m, m], [m, h - m]], [[w - m, m], [w - m, h - m]]], dtype=np.int64)
        lines = lines + tile_tl[None, None]
        for ln in lines:
            remove_in_ref.extend(boxes_intersect(ref_boxes, ln.flatten().tolist()).tolist())
    return sel, remove_in_ref
