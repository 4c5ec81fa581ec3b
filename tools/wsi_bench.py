"""WSI-mode measurement (BASELINE config 4 shape, scaled): synthetic array-backed slide + random
tissue mask through cerberus_b200.infer.wsi.InferManager, stage times from its log.
  python tools/wsi_bench.py [SIZE] [BATCH] [PP_TILE]        (torchrun for N GPUs)"""
import json
import os
import re
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import yaml  # noqa: E402


def single_rank_checks(m, tmp, size, pp_tile):
    """WSI_BENCH_VERIFY=1 on one GPU: (1) a second inference pass reproduces the canvas bit for bit,
    (2) the nuclei labelling of the first post-processing tile is reproducible, (3) the device
    instance table of that tile equals the OpenCV loop (oracle/instinfo_oracle.py)."""
    import torch
    from cerberus_b200 import _lib, instinfo
    from cerberus_b200.infer.wsi_geometry import filter_coordinates, get_coordinates
    from cerberus_b200.infer.wsi_reader import ArraySlide
    from oracle import instinfo_oracle as oi
    sl = ArraySlide.open(tmp + "/wsi/slide.npy", 0.5)
    msk = (cv2.cvtColor(cv2.imread(tmp + "/msk/slide.png"), cv2.COLOR_BGR2GRAY) > 0).astype(np.uint8)
    pi, po = get_coordinates((size, size), [448, 448], [144, 144], [144, 144])
    sel = filter_coordinates(msk, po, (size, size))
    again = m._infer_slide(sl, pi[sel], po[sel])
    d = (again != m.last_canvas)
    out = {"canvas_pixels_differing_between_two_passes": int(d.any(-1).sum())}
    del d, again
    ctx, lib = m.engine.ctx, m.engine.ctx.lib
    idx = m.engine.model.idx_dict
    t = (pp_tile // 144) * 144
    crop = m.last_canvas[:t, :t].contiguous()
    h, w, C = crop.shape
    labs = []
    for _ in range(2):
        labels = np.empty((h, w), dtype=np.int32)
        any_fg = np.zeros(1, dtype=np.int32)
        torch.cuda.synchronize()
        _lib.check(lib.cerb_postproc_nuclei(ctx.handle, _lib.ctypes.c_void_p(crop.data_ptr()), 1, h, w, C,
                                            idx["Nuclei-INST"][0],
                                            labels.ctypes.data_as(_lib.ctypes.c_void_p),
                                            any_fg.ctypes.data_as(_lib.ctypes.c_void_p), 1), "nuclei")
        labs.append(labels)
    out["tile0_label_pixels_differing_between_two_runs"] = int((labs[0] != labs[1]).sum())
    # the same tile on a FRESH context (no call history) and through the CPU oracle
    from cerberus_b200.engine import Context
    from cerberus_b200.postproc import post_process_batch
    from oracle import postproc_oracle as po
    c0 = idx["Nuclei-INST"][0]
    crop_host = crop[..., c0:c0 + 2].contiguous().cpu().numpy()
    fresh = Context(ctx.device, "f16")
    lab_fresh, _ = post_process_batch(fresh, crop_host[None], 0, "Nuclei")
    fresh.close()
    out["tile0_label_pixels_differing_fresh_context"] = int((lab_fresh[0] != labs[0]).sum())
    t0 = time.perf_counter()
    ref = po.proc_nuclei(crop_host)
    out["tile0_oracle_s"] = round(time.perf_counter() - t0, 2)
    out["tile0_label_pixels_differing_from_oracle"] = int((ref.astype(np.int64) != labs[0].astype(np.int64)).sum())
    out["tile0_fresh_vs_oracle"] = int((ref.astype(np.int64) != lab_fresh[0].astype(np.int64)).sum())
    typ = crop[..., idx["Nuclei-TYPE"][0]].cpu().numpy()
    t0 = time.perf_counter()
    a = instinfo.get_instance_info(labs[0], typ, ctx=ctx)
    t1 = time.perf_counter()
    b = oi.get_instance_info(labs[0], typ)
    t2 = time.perf_counter()
    bad = [k for k in b if k not in a or not (np.array_equal(a[k]["box"], b[k]["box"]) and
                                             np.array_equal(a[k]["contour"], b[k]["contour"]) and
                                             np.array_equal(a[k]["centroid"], b[k]["centroid"]) and
                                             a[k]["type"] == b[k]["type"] and a[k]["prob"] == b[k]["prob"])]
    out["tile0_instances"] = {"device": len(a), "opencv_loop": len(b), "mismatching": len(bad),
                              "first_bad": [int(k) for k in bad[:5]],
                              "device_s": round(t1 - t0, 3), "opencv_loop_s": round(t2 - t1, 3)}
    return out


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 4608
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    pp_tile = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from cerberus_b200 import synth
    from cerberus_b200.infer.wsi import InferManager
    tmp = os.environ.get("WSI_BENCH_DIR") or tempfile.mkdtemp(prefix="wsi_bench_")
    if world > 1:
        obj = [tmp]
        dist.broadcast_object_list(obj, src=0)
        tmp = obj[0]
    if rank == 0:
        os.makedirs(tmp + "/wsi", exist_ok=True)
        os.makedirs(tmp + "/msk", exist_ok=True)
        # H&E-like tiles; above 5120 px a 2560 x 2560 block is repeated (the generator is CPU-bound:
        # 100 s for the 6241 tiles of a 20000^2 slide; the work per patch does not depend on content)
        n = min(-(-size // 256), 10)
        tiles = synth.synthetic_tiles(n * n, 256, 256, seed=5)
        slide = tiles.reshape(n, n, 256, 256, 3).transpose(0, 2, 1, 3, 4).reshape(n * 256, n * 256, 3)
        reps = -(-size // (n * 256))
        if reps > 1:
            slide = np.tile(slide, (reps, reps, 1))
        np.save(tmp + "/wsi/slide.npy", np.ascontiguousarray(slide[:size, :size]))
        del slide
        # seeded blobs, ~40 % coverage, at 1/10 resolution
        rng = np.random.RandomState(0)
        m = cv2.GaussianBlur(rng.rand(size // 10, size // 10).astype(np.float32), (0, 0), size / 160.0)
        mask = (m > np.quantile(m, 0.6)).astype(np.uint8) * 255
        cv2.imwrite(tmp + "/msk/slide.png", mask)
        # the synthetic checkpoint is calibrated with CPU convolutions whose summation order depends
        # on the OpenMP thread count (torchrun exports OMP_NUM_THREADS=1): pin it, so that runs with
        # different launchers use bit-identical weights and their instance tables can be compared
        nt = torch.get_num_threads()
        torch.set_num_threads(1)
        synth.write_model_dir(tmp + "/model", seed=0)
        torch.set_num_threads(nt)
    if world > 1:
        dist.barrier()
    st = yaml.full_load(open(tmp + "/model/settings.yml"))
    m = InferManager(checkpoint_path=tmp + "/model/weights.tar",
                     decoder_dict=st["dataset_kwargs"]["req_target_code"], model_args=st["model_kwargs"],
                     device=local_rank, precision=os.environ.get("CERB_PRECISION", "f16"))
    run_args = {
        "nr_inference_workers": 0, "nr_post_proc_workers": 0, "batch_size": batch,
        "input_list": [tmp + "/wsi/slide.npy"], "mask_list": [tmp + "/msk/slide.png"],
        "output_dir": tmp + "/out", "patch_input_shape": 448, "patch_output_shape": 144,
        "save_thumb": False, "save_mask": False, "mask_dir": tmp + "/msk/", "postproc_list": [],
        "msk_dir": tmp + "/msk/", "tile_shape": 2048, "chunk_shape": 15000, "ambiguous_size": 64,
        "cache_path": tmp + "/cache", "logging_dir": tmp + "/log", "wsi_proc_mag": 0.5,
        "postproc_tile_shape": pp_tile,
    }
    m.keep_canvas = os.environ.get("WSI_BENCH_VERIFY", "0") == "1"
    m.return_inst_dicts = False  # the CLI's mode: instance tables stay arrays up to the .dat file
    if os.environ.get("WSI_BENCH_WARM", "1") == "1":
        # untimed pass over a small corner of the same slide: plan build, buffer allocation, NCCL
        # communicator set-up - one-off costs of the process, not of the slide
        warm = dict(run_args, output_dir=tmp + "/out_warm")
        m.warm_crop = 2304
        m.process_wsi_list(warm)
        m.warm_crop = None
        if world > 1:
            dist.barrier()
    t0 = time.perf_counter()
    res = m.process_wsi_list(run_args)
    dt = time.perf_counter() - t0
    verify = None
    if m.keep_canvas and world > 1 and rank == 0:
        # the all-reduced canvas of N ranks must equal the canvas one rank computes alone
        from cerberus_b200.infer.wsi_geometry import filter_coordinates, get_coordinates
        from cerberus_b200.infer.wsi_reader import ArraySlide
        sl = ArraySlide.open(tmp + "/wsi/slide.npy", 0.5)
        msk = (cv2.cvtColor(cv2.imread(tmp + "/msk/slide.png"), cv2.COLOR_BGR2GRAY) > 0).astype(np.uint8)
        pi, po = get_coordinates((size, size), [448, 448], [144, 144], [144, 144])
        sel = filter_coordinates(msk, po, (size, size))
        m.force_single = True
        alone = m._infer_slide(sl, pi[sel], po[sel])
        m.force_single = False
        d = (alone - m.last_canvas).abs()
        verify = {"max_abs_diff": float(d.max()), "pixels_differing": int((d.amax(-1) > 0).sum())}
    if m.keep_canvas and world == 1:
        verify = single_rank_checks(m, tmp, size, pp_tile)
    if rank == 0:
        logs = sorted(f for f in os.listdir(tmp + "/log") if "rank" not in f)
        text = open(os.path.join(tmp, "log", logs[-1])).read()
        stages = {k: float(v) for k, v in re.findall(r"INFO - ([A-Za-z& ]+ Time): ([0-9.]+)", text)}
        npatch = int(re.search(r"(\d+) selected", text).group(1))
        ws = re.search(r"Nuclei watershed: (.*)", text)
        r = res["slide"]
        import hashlib
        from cerberus_b200.infer.dat_writer import InstanceStore

        def table_sha1(v):
            h = hashlib.sha1()
            if isinstance(v, InstanceStore):
                for col in v.columns():
                    if col is not None:
                        h.update(np.ascontiguousarray(col).tobytes())
                return h.hexdigest(), int(v.alive().sum())
            for d in v.values():
                for f in ("box", "centroid", "contour"):
                    h.update(np.ascontiguousarray(d[f]).tobytes())
                h.update(repr((d.get("type"), d.get("type_prob"))).encode())
            return h.hexdigest(), len(v)

        tables = {k: table_sha1(v) for k, v in r.items() if k in ("Nuclei", "Gland", "Lumen")}
        line = {"slide": [size, size], "n_gpus": world, "batch": batch, "postproc_tile": pp_tile,
                "patches_448_to_144": npatch, "wall_s": dt, "stages_s": stages,
                "nuclei_detail": ws.group(1) if ws else None, "multi_rank_canvas_check": verify,
                "patches_per_s_inference": npatch / stages["Inference Time"],
                "tiles256_equivalent_per_s": npatch / stages["Inference Time"] * (448 * 448) / (256 * 256),
                "instances": {k: t[1] for k, t in tables.items()},
                "instance_table_sha1": {k: t[0] for k, t in tables.items()},
                "canvas_exchange_s": getattr(m, "t_exchange", None),
                "dat_bytes": os.path.getsize(tmp + "/out/dat/slide.dat"),
                "precision": m.precision}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
