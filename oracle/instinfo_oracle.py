"""TEST INFRASTRUCTURE - CPU restatement of loader/postproc.py:12-98 (get_inst_info_dict) and
misc/utils.py:82-91 (get_bounding_box): per-instance box / centroid / contour / majority type on
the host with OpenCV, exactly as the reference does. It is the checker of the device path
(cerberus_b200/instinfo.py over cerb_inst_info); only tests/, smoke() and bench.py's CPU legs may
import it. Pinned against the unmodified reference function in tests/golden/tile_postproc.npz
(oracle/gen_golden.py runs the reference's own get_inst_info_dict through _post_process_patches)
and tests/golden/instinfo.npz."""
import cv2
import numpy as np


def get_bounding_box(img):
    rows = np.any(img, axis=1)
    cols = np.any(img, axis=0)
    rmin, rmax = np.where(rows)[0][[0, -1]]
    cmin, cmax = np.where(cols)[0][[0, -1]]
    rmax += 1
    cmax += 1
    return [rmin, rmax, cmin, cmax]


def get_inst_info_dict(inst_map, type_map, ds_factor=1.0):
    inst_id_list = np.unique(inst_map)[1:]  # reference quirk: drops the smallest value
    info = {}
    for inst_id in inst_id_list:
        single = inst_map == inst_id
        rmin, rmax, cmin, cmax = get_bounding_box(single)
        bbox = np.array([[rmin, cmin], [rmax, cmax]])
        single = single[bbox[0][0]:bbox[1][0], bbox[0][1]:bbox[1][1]].astype(np.uint8)
        moment = cv2.moments(single)
        contour = cv2.findContours(single, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        contour = np.squeeze(contour[0][0].astype("int32"))
        if contour.shape[0] < 3:
            continue
        if len(contour.shape) != 2:
            continue
        centroid = np.array([moment["m10"] / moment["m00"], moment["m01"] / moment["m00"]])
        contour[:, 0] += bbox[0][1]
        contour[:, 1] += bbox[0][0]
        centroid[0] += bbox[0][1]
        centroid[1] += bbox[0][0]
        info[inst_id] = {"box": bbox, "centroid": centroid, "contour": contour}

    if type_map is not None:
        for inst_id in list(info.keys()):
            rmin, cmin, rmax, cmax = (info[inst_id]["box"]).flatten()
            crop = inst_map[rmin:rmax, cmin:cmax] == inst_id
            inst_type = type_map[rmin:rmax, cmin:cmax][crop]
            type_list, type_pixels = np.unique(inst_type, return_counts=True)
            type_list = sorted(zip(type_list, type_pixels), key=lambda x: x[1], reverse=True)
            inst_type = type_list[0][0]
            if inst_type == 0 and len(type_list) > 1:
                inst_type = type_list[1][0]
            type_dict = {v[0]: v[1] for v in type_list}
            info[inst_id]["type"] = int(inst_type)
            info[inst_id]["type_prob"] = float(type_dict[inst_type] / (np.sum(crop) + 1.0e-6))

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


def get_instance_info(pred_inst, pred_type=None):
    """tiatoolbox HoVerNet.get_instance_info (infer/wsi.py:150) on the host with OpenCV: box is
    flat [x0, y0, x1, y1]."""
    info = {}
    ids = np.unique(pred_inst)[1:]
    if len(ids) == 0:
        return info
    # The original builds `pred_inst == inst_id` over the WHOLE tile for every instance
    # (O(instances x pixels): minutes for a 4032^2 tile); one find_objects pass gives the same
    # boxes, and the per-instance mask is then cut from the box only.
    from scipy import ndimage
    lab = pred_inst if pred_inst.dtype.kind in "iu" else pred_inst.astype(np.int64)
    if lab.min() < 0:
        raise ValueError("negative instance ids")
    slices = ndimage.find_objects(lab)
    for inst_id in ids:
        sl = slices[int(inst_id) - 1]
        box = np.array([sl[1].start, sl[0].start, sl[1].stop, sl[0].stop])
        crop = (lab[sl] == inst_id).astype(np.uint8)
        moment = cv2.moments(crop)
        contour = cv2.findContours(crop, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        contour = np.squeeze(contour[0][0].astype(np.int32))
        if contour.shape[0] < 3 or len(contour.shape) != 2:
            continue
        centroid = np.array([moment["m10"] / moment["m00"], moment["m01"] / moment["m00"]])
        contour = contour + box[:2][None]
        centroid = centroid + box[:2]
        info[inst_id] = {"box": box, "centroid": centroid, "contour": contour, "prob": None,
                         "type": None}
    if pred_type is not None:
        for inst_id in list(info.keys()):
            c0, r0, c1, r1 = info[inst_id]["box"]
            m = pred_inst[r0:r1, c0:c1] == inst_id
            t = pred_type[r0:r1, c0:c1][m]
            tl, tp = np.unique(t, return_counts=True)
            pairs = sorted(zip(tl, tp), key=lambda x: x[1], reverse=True)
            inst_type = pairs[0][0]
            if inst_type == 0 and len(pairs) > 1:
                inst_type = pairs[1][0]
            d = {v[0]: v[1] for v in pairs}
            info[inst_id]["type"] = int(inst_type)
            info[inst_id]["prob"] = float(d[inst_type] / (np.sum(m) + 1.0e-6))
    return info
