"""Instance tables: device path (cerb_inst_info) vs the reference's per-instance
OpenCV loop (oracle/instinfo_oracle.py = loader/postproc.py:12-98 restated) on the label maps of
the device post-processing.   python tools/instinfo_bench.py [SIZE]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    from cerberus_b200 import instinfo, synth
    from cerberus_b200.engine import Context
    from cerberus_b200.postproc import post_process_batch
    from oracle import instinfo_oracle as oi
    ctx = Context(0, "f16")
    out = {"size": size}
    for tissue, up in (("Nuclei", 1), ("Nuclei", 2), ("Gland", 1)):
        n = size if up == 1 else size // 4
        canvas = np.zeros((1, n, n, 2), np.float32)
        canvas[0] = synth.postproc_field(n, n, tissue, seed=1)
        lab = post_process_batch(ctx, canvas, 0, tissue, 1.0)[0][0]
        typ = (np.arange(n)[None, :] // 64 % 5 + np.zeros((n, 1))).astype(np.float32)
        instinfo.inst_table(ctx, lab, typ, up=up)  # warm-up (workspace allocation)
        t0 = time.perf_counter()
        table = instinfo.inst_table(ctx, lab, typ, up=up)
        t_table = time.perf_counter() - t0
        t0 = time.perf_counter()
        info = instinfo.get_inst_info_dict(lab, typ, ctx=ctx, up=up)
        t_dev = time.perf_counter() - t0
        a, t = lab, typ
        if up != 1:
            import cv2
            a = cv2.resize(lab, (0, 0), fx=up, fy=up, interpolation=cv2.INTER_NEAREST)
            t = cv2.resize(typ, (0, 0), fx=up, fy=up, interpolation=cv2.INTER_NEAREST)
        t0 = time.perf_counter()
        ref = oi.get_inst_info_dict(a, t) if len(table.ids) < 3000 else oi.get_instance_info(a, t)
        t_ref = time.perf_counter() - t0
        out["%s_up%d" % (tissue, up)] = {
            "image": [n * up, n * up], "instances": int(len(table.ids)),
            "contour_points": int(table.contour_off[-1]),
            "device_table_ms": round(t_table * 1e3, 3), "device_dict_ms": round(t_dev * 1e3, 3),
            "opencv_loop_ms": round(t_ref * 1e3, 3),
            "opencv_loop": "get_inst_info_dict" if len(table.ids) < 3000 else
                           "find_objects variant (the literal loop is O(instances x pixels))",
            "same_instances": len(info) == len(ref)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
