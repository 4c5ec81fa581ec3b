#!/usr/bin/env python
"""bench.py — tiles/s of the Cerberus tiled multi-task inference hot path on B200.

Contract (see DESIGN.md section "Measurement"):
  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                         reference arm: the CPU path

One "step" = one batch of `--batch` synthetic 256x256x3 uint8 tiles through the full six-head
model (+ on-device post-processing unless --no-postproc). One rank per GPU; ranks process
independent batches (no data-path collective; weights are NCCL-broadcast once at start-up).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE = 256
GFLOP_PER_TILE_6HEAD = 121.128  # algorithmic conv FLOPs, 2*M*N*K over every nn.Conv2d


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--precision", default="f16", choices=["f16", "f16x2"])
    ap.add_argument("--no-postproc", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-batch", type=int, default=4)
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tensor_burst": d["bf16_tflops"], "tensor_sustained": d["bf16_tflops_sustained"],
                "hbm": d["hbm_gbs"], "source": "measured"}
    return {"tensor_burst": 1590.0, "tensor_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_rate(args, seconds, batch):
    """The reference's CPU path (oracle restatement: models/run_desc.py:439-502 +
    loader/postproc.py post_process) on this host's cores. Returns (tiles/s, cores, sample)."""
    import torch
    from cerberus_b200 import synth
    from oracle import net_oracle
    margs = synth.model_args()
    sd = synth.make_state_dict(seed=0)
    tiles = synth.synthetic_tiles(batch, TILE, TILE, seed=123)
    torch.set_num_threads(host_cores())  # torchrun exports OMP_NUM_THREADS=1
    cores = torch.get_num_threads()
    post = None
    if not args.no_postproc:
        try:
            from oracle import pipeline_oracle
            post = pipeline_oracle.postprocess_step
        except Exception:
            post = None

    def one():
        step, _ = net_oracle.infer_step(sd, tiles, TILE, margs["decoder_kwargs"],
                                        margs["considered_tasks"])
        if post is not None:
            post(step, margs)

    one()  # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        one()
        n += batch
        if time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    sample = "%d tiles (batches of %d) of the bench workload, oracle infer_step%s, torch %d threads" % (
        n, batch, " + post_process" if post is not None else "", cores)
    return n / dt, cores, sample


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cerberus_b200 import synth
    from oracle import net_oracle
    margs = synth.model_args()
    sd = synth.make_state_dict(seed=0)
    b = args.ref_batch
    tiles = synth.synthetic_tiles(b, TILE, TILE, seed=123)
    torch.set_num_threads(host_cores())  # torchrun exports OMP_NUM_THREADS=1
    post = None
    if not args.no_postproc:
        try:
            from oracle import pipeline_oracle
            post = pipeline_oracle.postprocess_step
        except Exception:
            post = None

    def one():
        step, _ = net_oracle.infer_step(sd, tiles, TILE, margs["decoder_kwargs"],
                                        margs["considered_tasks"])
        if post is not None:
            post(step, margs)

    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    val = b * args.steps / dt
    cores = torch.get_num_threads()
    sample = "each step = %d tiles of the bench workload on the host CPU (oracle port of the reference path)" % b
    line = {
        "impl": "reference", "metric": "tiles/sec (256x256x3) end-to-end incl. post-proc",
        "value": val, "unit": "tiles/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, b, post is not None),
        "cpu_baseline": {"value": val, "unit": "tiles/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, batch, postproc):
    return {
        "workload": "batch=%d synthetic 256x256x3 tiles, full 6-head Cerberus (ResNet34 encoder, "
                    "5 seg decoders + patch-cls)%s, forward only (no grad)" % (
                        batch, " + instance post-processing (nuclei watershed, gland/lumen)" if postproc else ""),
        "tile": [TILE, TILE, 3], "batch_per_gpu": batch, "heads": 6,
        "precision": args.precision,
        "l2": "per-step working set (GBs of activations) >> 126 MB L2; inputs rotate over 4 batches",
    }


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cerberus_b200 import synth
    from cerberus_b200.engine import Context, ForwardPlan, canvas_to_step_outputs
    from cerberus_b200.plan import PackedModel, PlanSpec

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    margs = synth.model_args()
    # rank 0 folds + packs the checkpoint; the packed blob is NCCL-broadcast once
    from cerberus_b200.dist import broadcast_packed_model
    model = PackedModel(synth.make_state_dict(seed=0), margs) if rank == 0 else None
    model = broadcast_packed_model(model, margs, rank, world, torch.device("cuda", local_rank))

    B = args.batch
    from cerberus_b200.engine import Engine
    eng = Engine(None, None, device=local_rank, precision=args.precision, packed=model)
    ctx = eng.ctx
    plan = eng.plan_for(B, TILE, TILE, TILE, TILE)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    post = None if args.no_postproc else True

    n_in = 4
    host_batches = [synth.synthetic_tiles(B, TILE, TILE, seed=1000 * rank + i) for i in range(n_in)]
    pinned = [torch.from_numpy(b).pin_memory() for b in host_batches]
    dev_batches = [p.cuda() for p in pinned]
    torch.cuda.synchronize()

    pipe = None
    if post is not None:
        from cerberus_b200.pipeline import TilePipeline
        pipe = TilePipeline(eng, B, TILE, TILE)
        if os.environ.get("CERB_WS_MODE"):  # robustness experiment: force the exact fallback
            pipe.pctx.set_option("ws_mode", int(os.environ["CERB_WS_MODE"]))
        post = pipe  # post-processing runs on the pipeline's second context

    def step_resident(i):
        if pipe is not None:  # forward of batch i overlaps the post-processing of batch i-1
            pipe.submit(device_ptr=dev_batches[i % n_in].data_ptr(), download=False)
        else:
            plan.run(device_ptr=dev_batches[i % n_in].data_ptr())

    def launches_now():
        return pipe.launch_count if pipe is not None else ctx.launch_count

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_resident(i)
    ctx.sync()
    if pipe is not None:
        pipe.flush()
    l0 = launches_now()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        step_resident(i)
    if pipe is not None:
        pipe.join_streams()  # the end event on the compute stream then covers the post stream
    e1.record(stream)
    host_enqueue_ms = 1e3 * (time.perf_counter() - t_host0) / args.steps
    ctx.sync()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = launches_now() - l0
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    per_rank_ms = [ms]
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank_ms = [float(x.item()) / args.steps for x in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    else:
        per_rank_ms = [ms / args.steps]
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---- e2e: host buffers in, host results out, through the public API. Every step uploads its
    # uint8 batch from host memory and downloads its result (the label maps); the copies of
    # neighbouring steps overlap the compute (cerberus_b200.pipeline.TilePipeline).
    host_np = [p.numpy() for p in pinned]
    if pipe is not None:
        for i in range(max(2, args.warmup)):
            pipe.submit(host_np[i % n_in])
        pipe.flush()
        barrier()
        t0 = time.perf_counter()
        got = 0
        for i in range(args.steps):
            if pipe.submit(host_np[i % n_in]) is not None:
                got += 1
        got += len(pipe.flush())
        barrier()
        dt = time.perf_counter() - t0
        assert got == args.steps, (got, args.steps)
        h2d, d2h = int(pipe.h2d_bytes), int(pipe.d2h_bytes)
    else:
        for i in range(max(1, args.warmup // 2)):
            plan.run(host_np[i % n_in])
            plan.read_canvas()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            plan.run(host_np[i % n_in])
            plan.read_canvas()
        barrier()
        dt = time.perf_counter() - t0
        h2d, d2h = int(pinned[0].numel()), int(B * TILE * TILE * model.canvas_c * 4)
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * B * args.steps / float(t.item())

    # ---- roofline of the dominant kernel: per-op CUDA events on the ctx stream (cerb_plan_profile)
    # Dominant = conv64_kernel on the full-resolution 64->64 3x3 layers (largest single share of
    # the step); the aggregate over every conv launch is reported next to it.
    peaks = load_peaks()
    roof = None
    try:
        from cerberus_b200 import _lib as L_
        from cerberus_b200.engine import profile_ops
        prof = profile_ops(plan, dev_batches[0].data_ptr(), reps=3)
        spec = plan.spec
        dom_ms, dom_fl, dom_n = 0.0, 0.0, 0
        all_ms, n_conv = 0.0, 0
        for (kind, ms_), op in zip(prof, spec.ops):
            if kind != "conv":
                continue
            all_ms += ms_
            n_conv += 1
            tid = op["in0"] if op["aux_classes"] else op["out"]
            _, n_, h_, w_, _, _ = spec.tensors[tid]
            if (op["in_c"] == 64 and op["cout"] == 64 and op["kh"] == 3 and op["stride"] == 1
                    and h_ == TILE and w_ == TILE and not op["stem"] and not op["aux_classes"]):
                dom_ms += ms_
                dom_n += 1
                dom_fl += 2.0 * n_ * h_ * w_ * 64 * 576
        flops = spec.conv_flops()
        achieved = dom_fl / (dom_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tpath) and B == 32:
            tj = json.load(open(tpath)).get("%s 256x256 64->64 3x3 batch 32" % (
                "conv64x_kernel" if ctx.conv64_mode == 3 else "conv64_kernel"))
            if tj:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        agg = flops / (all_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["tensor_sustained"],
                "unit": "TFLOP/s", "frac": achieved / peaks["tensor_sustained"], "traffic": traffic,
                "kernel": ("conv64x_kernel (even/odd N=128 formulation)" if ctx.conv64_mode == 3
                           else "conv64_kernel") + ", 256x256 64->64 3x3, batch %d: %d launches/step, %.4f ms avg, "
                          "%.1f GFLOP (algorithmic) per launch" % (B, dom_n, dom_ms / max(dom_n, 1),
                                                                   dom_fl / max(dom_n, 1) / 1e9),
                "peak_source": peaks["source"] + " bf16 cuBLAS sustained (kernel timed inside a long step)",
                "frac_of_burst": achieved / peaks["tensor_burst"],
                "traffic_note": "DRAM bytes per launch from ncu --set full (profiles/r1_traffic.json); "
                                "algorithmic bytes 536.9 MB (fp16 in + out)",
                "all_convs": {"achieved": agg, "frac": agg / peaks["tensor_sustained"],
                              "launches_per_step": n_conv, "ms_per_step": all_ms,
                              "gflop_per_step": flops / 1e9},
                "step_breakdown_ms": {k: sum(m for kk, m in prof if kk == k)
                                      for k in sorted(set(k for k, _ in prof))}}
    except Exception as e:  # profiling is evidence, never a reason to lose the line
        roof = {"bound": "tensor", "achieved": None, "peak": peaks["tensor_sustained"],
                "unit": "TFLOP/s", "frac": None, "traffic": None, "error": str(e)}

    ws_stats = None
    if pipe is not None:
        pl = pipe.pctx
        imgs = int(pl.lib.cerb_ctx_stat(pl.handle, b"ws_images"))
        ws_stats = {"images": imgs,
                    "exact_fallback_tie": int(pl.lib.cerb_ctx_stat(pl.handle, b"ws_tie_fallbacks")),
                    "exact_fallback_capacity": int(pl.lib.cerb_ctx_stat(pl.handle, b"ws_capacity_fallbacks")),
                    "note": "nuclei watershed, all steps incl. warm-up and e2e: images handled by the "
                            "component-parallel path vs redone by the exact whole-tile emulation"}
    if world > 1 and pipe is not None:  # which ranks were slowed by exact-watershed images
        mine = torch.tensor([float(ws_stats["exact_fallback_tie"])], device="cuda", dtype=torch.float64)
        allw = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allw, mine)
        ws_stats["exact_fallback_tie_per_rank"] = [int(x.item()) for x in allw]
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_reference_rate(args, args.cpu_seconds, args.ref_batch)
        cpu = {"value": v, "unit": "tiles/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "tiles/sec (256x256x3) end-to-end incl. post-proc" if not args.no_postproc
            else "tiles/sec (256x256x3) forward only (BASELINE config 2; --no-postproc)",
            "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": workload_config(args, B, post is not None),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "tiles/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "host_enqueue_ms_per_step": host_enqueue_ms,
            "ms_per_step_per_rank": per_rank_ms,
            "watershed": ws_stats,
            "roofline": roof,
            "cpu_baseline": cpu,
            "tflops_forward": value * GFLOP_PER_TILE_6HEAD / 1e3,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
