"""ORACLE - test infrastructure only. CPU restatement of the WSI tail of the reference
(infer/wsi.py:137-268 nuclei tiles + de-duplication, :690-715 Patch-Class map, :720-840 gland /
lumen regions) applied to a merged float32 prediction canvas [H, W, C]. The arithmetic is the
oracle's (oracle/postproc_oracle.py, OpenCV, scipy); the tile placement / selection rules are the
restated tiatoolbox / shapely helpers of cerberus_b200/infer/wsi_geometry.py (host logic, parity
unpinned - the originals are not installable offline). Only tests import this module."""
import cv2
import numpy as np
from scipy import ndimage

from cerberus_b200.infer.wsi_geometry import boxes_intersect, get_tile_info, select_tile_instances
from oracle.instinfo_oracle import get_inst_info_dict, get_instance_info
from oracle import postproc_oracle as po


def nuclei_tables(canvas, idx_dict, patch_outputs, pp_tile_shape, pout, margin):
    """Returns the list of instance dicts (WSI coordinates), order-independent."""
    H, W, _ = canvas.shape
    sets = get_tile_info((W, H), pp_tile_shape, pout, margin)
    acc = {}
    uid = 0
    for mode, (bounds, flags) in enumerate(sets):
        results = []
        for ti, tb in enumerate(bounds):
            if len(boxes_intersect(patch_outputs, tb)) == 0:
                continue
            x0, y0, x1, y1 = [int(v) for v in tb]
            crop = canvas[y0:y1, x0:x1]
            # infer/wsi.py:147-149: raw_map = concat(inst (2 ch), type (1 ch)); idx {"Nuclei-INST": [0,2], "Nuclei-TYPE": [2,4]}
            lo = idx_dict["Nuclei-INST"][0]
            raw = np.concatenate([crop[..., lo:lo + 2], crop[..., idx_dict["Nuclei-TYPE"][0]:idx_dict["Nuclei-TYPE"][0] + 1]], -1)
            inst_map, type_map = po.post_process(raw, {"Nuclei-INST": [0, 2], "Nuclei-TYPE": [2, 4]}, "Nuclei")
            info = get_instance_info(inst_map.astype(np.int32), type_map)
            if not info:
                results.append(({}, []))
                continue
            boxes = np.array([v["box"] for v in info.values()])
            ref_uids = list(acc.keys())
            ref_boxes = np.array([acc[u]["box"] for u in ref_uids]) if (mode == 3 and ref_uids) else None
            sel, sel_ref = select_tile_instances(boxes, tb, flags[ti], mode, margin, ref_boxes)
            keys = list(info.keys())
            drop = set(keys[i] for i in sel)
            new = {}
            for k, v in info.items():
                if k in drop:
                    continue
                v["box"] = v["box"] + np.concatenate([tb[:2]] * 2)
                v["centroid"] = v["centroid"] + tb[:2]
                v["contour"] = v["contour"] + tb[:2]
                new["u%d" % uid] = v
                uid += 1
            results.append((new, [ref_uids[i] for i in sel_ref]))
        for new, rem in results:
            acc.update(new)
            for u in rem:
                acc.pop(u, None)
    return list(acc.values())


def gland_lumen_tables(canvas, idx_dict, wsi_mask):
    H, W, _ = canvas.shape
    ratio = wsi_mask.shape[0] / H
    lab = ndimage.label(wsi_mask)[0]
    ids = np.unique(lab).tolist()
    regions = []
    if len(ids) > 1:
        for r in ids[1:]:
            m = lab == r
            rows, cols = np.any(m, 1), np.any(m, 0)
            rmin, rmax = np.where(rows)[0][[0, -1]]
            cmin, cmax = np.where(cols)[0][[0, -1]]
            regions.append([rmin, rmax + 1, cmin, cmax + 1])
    else:
        regions.append([0, lab.shape[0], 0, lab.shape[1]])
    out = {"Gland": [], "Lumen": []}
    for ridx, ti in enumerate(regions):
        rmin, rmax = int(round(ti[0] / ratio)), int(round(ti[1] / ratio))
        cmin, cmax = int(round(ti[2] / ratio)), int(round(ti[3] / ratio))
        mask_idx = lab[ti[0]:ti[1], ti[2]:ti[3]] == ridx + 1
        inst, typ = {}, {}
        for tissue in ("Gland", "Lumen"):
            maps, new_idx, ch = [], {}, 0
            for ot in ("INST", "TYPE"):
                key = tissue + "-" + ot
                if key in idx_dict and key in ("Gland-INST", "Gland-TYPE", "Lumen-INST"):
                    m = np.array(canvas[rmin:rmax, cmin:cmax, idx_dict[key][0]:idx_dict[key][1]])
                    if m.shape[0] != mask_idx.shape[0] and m.shape[1] != mask_idx.shape[1]:
                        mask_idx = cv2.resize(mask_idx.astype("uint8"), (m.shape[1], m.shape[0]),
                                              interpolation=cv2.INTER_NEAREST)
                    if mask_idx.ndim == 2:
                        mask_idx = np.expand_dims(mask_idx, -1)
                    m = m * mask_idx
                    maps.append(m)
                    new_idx[key] = [ch, ch + m.shape[-1]]
                    ch += m.shape[-1]
            tile = np.concatenate(maps, -1).astype(np.float32)
            tile = cv2.resize(tile, (0, 0), fx=0.5, fy=0.5)
            inst[tissue], typ[tissue] = po.post_process(tile, new_idx, tissue, 0.5)
        g = inst["Gland"].copy()
        g[g > 0] = 1
        inst["Lumen"] = g * inst["Lumen"]
        for tissue in ("Gland", "Lumen"):
            info = get_inst_info_dict(inst[tissue], typ[tissue], 0.5)
            for v in info.values():
                v["box"] = v["box"] + [cmin, rmin]
                v["contour"] = v["contour"] + [cmin, rmin]
                v["centroid"] = v["centroid"] + [cmin, rmin]
                b = v["box"]
                v["box"] = np.array([b[0][1], b[0][0], b[1][1], b[1][0]])
                out[tissue].append(v)
    return out
