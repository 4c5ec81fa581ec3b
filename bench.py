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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-gpu"],
                    help="reference = the reference's CPU path (the driver's second arm); torch-gpu = EXTRA CONTEXT "
                         "only: the unmodified reference network run by PyTorch / cuDNN on cuda:0 (fp16, "
                         "channels_last), forward only")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--precision", default="f16", choices=["f16", "f16x2"])
    ap.add_argument("--no-postproc", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-second-precision", action="store_true",
                    help="skip timing the other precision mode (f16x2 when --precision f16)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block (CPU oracle, ~10 s)")
    ap.add_argument("--parity-tiles", type=int, default=4)
    ap.add_argument("--config", default="3", choices=["1", "3", "tile448"],
                    help="BASELINE.json config: 3 (default; 2 with --no-postproc) = batch of 256^2 tiles, six "
                         "heads; 1 = ONE 256^2 image, encoder + Nuclei head, batch 1, through the "
                         "run_infer_tile.py plumbing (InferManager), CPU reference beside it; tile448 = a "
                         "directory of PNGs through process_file_list at the CLI defaults 448/144")
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


def cpu_step_fn(args, batch):
    """One CPU step of the bench workload on `batch` tiles. Prefers the UNMODIFIED reference
    modules (oracle/_ref, staged by oracle/make_ref.py; kind "reference"), falls back to the
    oracle restatement (kind "port"). Returns (callable, kind, description)."""
    import torch
    from cerberus_b200 import synth
    margs = synth.model_args()
    sd = synth.make_state_dict(seed=0)
    tiles = synth.synthetic_tiles(batch, TILE, TILE, seed=123)
    torch.set_num_threads(host_cores())  # torchrun exports OMP_NUM_THREADS=1
    from oracle import ref_runner
    if ref_runner.available() and not os.environ.get("CERB_BENCH_PORT"):
        ref = ref_runner.ReferenceTilePath(sd, margs)
        if args.no_postproc:
            return (lambda: ref.infer_step(tiles, TILE)), "reference", \
                "unmodified reference create_model + infer_step (models/run_desc.py:439-502)"
        return (lambda: ref.step_with_labels(tiles)), "reference", \
            "unmodified reference create_model + infer_step + PostProcInstErodedContourMap.post_process " \
            "(skimage watershed / remove_small_objects restated in C)"
    from oracle import net_oracle, pipeline_oracle

    def one():
        step, _ = net_oracle.infer_step(sd, tiles, TILE, margs["decoder_kwargs"],
                                        margs["considered_tasks"])
        if not args.no_postproc:
            pipeline_oracle.postprocess_step(step, margs)

    return one, "port", "oracle restatement of infer_step%s" % ("" if args.no_postproc else " + post_process")


def cpu_reference_rate(args, seconds, batch):
    """The reference's CPU path on this host's cores, bounded to `seconds` of work.
    Returns (tiles/s, cores, sample, kind)."""
    import torch
    one, kind, what = cpu_step_fn(args, batch)
    cores = torch.get_num_threads()
    one()  # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        one()
        n += batch
        if time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    sample = "%d tiles (batches of %d) of the bench workload: %s, torch %d threads" % (n, batch, what, cores)
    return n / dt, cores, sample, kind


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b = args.ref_batch
    one, kind, what = cpu_step_fn(args, b)
    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    val = b * args.steps / dt
    cores = torch.get_num_threads()
    sample = "each step = %d tiles of the bench workload on the host CPU: %s" % (b, what)
    line = {
        "impl": "reference", "metric": "tiles/sec (256x256x3) end-to-end incl. post-proc",
        "value": val, "unit": "tiles/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, b, not args.no_postproc),
        "cpu_baseline": {"value": val, "unit": "tiles/s", "cores": cores, "kind": kind, "sample": sample},
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


class Arm:
    """One precision mode of the CUDA path on this rank: engine + plan + streaming pipeline."""

    def __init__(self, model, local_rank, precision, batch, postproc):
        import torch
        from cerberus_b200.engine import Engine
        self.eng = Engine(None, None, device=local_rank, precision=precision, packed=model)
        self.ctx = self.eng.ctx
        self.precision = precision
        self.B = batch
        self.plan = self.eng.plan_for(batch, TILE, TILE, TILE, TILE)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=local_rank)
        self.pipe = None
        if postproc:
            from cerberus_b200.pipeline import TilePipeline
            self.pipe = TilePipeline(self.eng, batch, TILE, TILE)
            if os.environ.get("CERB_WS_MODE"):  # robustness experiment: force the exact fallback
                self.pipe.pctx.set_option("ws_mode", int(os.environ["CERB_WS_MODE"]))

    def launches(self):
        return self.pipe.launch_count if self.pipe is not None else self.ctx.launch_count

    def step_resident(self, dev_ptr):
        if self.pipe is not None:  # forward of batch i overlaps the post-processing of batch i-1
            self.pipe.submit(device_ptr=dev_ptr, download=False)
        else:
            self.plan.run(device_ptr=dev_ptr)

    def settle(self):
        self.ctx.sync()
        if self.pipe is not None:
            self.pipe.flush()

    def close(self):
        if self.pipe is not None:
            self.pipe.close()
        self.eng.close()


def time_resident(arm, dev_batches, steps, warmup, barrier, forward_only=False):
    """K steps with the uint8 batches already in HBM; CUDA events on the ctx stream.
    Returns (ms_total, launches, host_enqueue_ms_per_step)."""
    import torch
    n_in = len(dev_batches)
    run = (lambda i: arm.plan.run(device_ptr=dev_batches[i % n_in].data_ptr())) if forward_only \
        else (lambda i: arm.step_resident(dev_batches[i % n_in].data_ptr()))
    for i in range(warmup):
        run(i)
    arm.settle()
    l0 = arm.launches()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(arm.stream)
    t0 = time.perf_counter()
    for i in range(steps):
        run(i)
    if arm.pipe is not None and not forward_only:
        arm.pipe.join_streams()  # the end event on the compute stream then covers the post stream
    e1.record(arm.stream)
    enq = 1e3 * (time.perf_counter() - t0) / steps
    arm.settle()
    barrier()
    return e0.elapsed_time(e1), arm.launches() - l0, enq


def time_e2e(arm, host_np, steps, warmup, barrier):
    """K steps through the public API with HOST buffers: every step uploads its uint8 batch and
    downloads its result (label maps; the fp32 canvas when there is no post-processing). The
    copies of neighbouring steps overlap the compute (cerberus_b200.pipeline.TilePipeline).
    Returns (seconds, h2d_bytes_per_step, d2h_bytes_per_step)."""
    n_in = len(host_np)
    pipe, plan = arm.pipe, arm.plan
    if pipe is not None:
        for i in range(max(2, warmup)):
            pipe.submit(host_np[i % n_in])
        pipe.flush()
        barrier()
        t0 = time.perf_counter()
        got = 0
        for i in range(steps):
            if pipe.submit(host_np[i % n_in]) is not None:
                got += 1
        got += len(pipe.flush())
        barrier()
        dt = time.perf_counter() - t0
        assert got == steps, (got, steps)
        return dt, int(pipe.h2d_bytes), int(pipe.d2h_bytes)
    for i in range(max(1, warmup // 2)):
        plan.run(host_np[i % n_in])
        plan.read_canvas()
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        plan.run(host_np[i % n_in])
        plan.read_canvas()
    barrier()
    dt = time.perf_counter() - t0
    return dt, int(host_np[0].size), int(arm.B * TILE * TILE * arm.eng.model.canvas_c * 4)


def conv_roofline(arm, dev_ptr, peaks):
    """Roofline of the dominant kernel: per-op CUDA events on the ctx stream (cerb_plan_profile).
    Dominant = the 64->64 3x3 kernel on the full-resolution layers (largest single share of the
    step); the aggregate over every conv launch and the per-kind breakdown are reported with it."""
    from cerberus_b200.engine import profile_ops
    plan, ctx, B = arm.plan, arm.ctx, arm.B
    prof = profile_ops(plan, dev_ptr, reps=3)
    spec = plan.spec
    dom_ms, dom_fl, dom_n = 0.0, 0.0, 0
    all_ms, n_conv = 0.0, 0
    enc_ms, enc_fl = 0.0, 0.0
    in_encoder = True
    fused_ms, fused_n = 0.0, 0
    prev = (None, 0.0, None)
    for (kind, ms_), op in zip(prof, spec.ops):
        # the library folds an UPADD into the 64->64 convolution that is its only reader (option
        # fuse_upadd): the UPADD step then takes no time and the convolution does both
        folded = (kind == "conv" and prev[0] == "upadd" and prev[2]["out"] == op["in0"] and prev[1] < 0.004)
        prev = (kind, ms_, op)
        if kind == "pclass":
            in_encoder = False
        if kind != "conv":
            if in_encoder:
                enc_ms += ms_
            continue
        all_ms += ms_
        n_conv += 1
        tid = op["in0"] if op["aux_classes"] else op["out"]
        _, n_, h_, w_, _, _ = spec.tensors[tid]
        cin = 3 if op["stem"] else op["in_c"]
        fl = 2.0 * n_ * h_ * w_ * op["cout"] * op["kh"] * op["kw"] * cin
        if in_encoder:
            enc_ms += ms_
            enc_fl += fl
        if (op["in_c"] == 64 and op["cout"] == 64 and op["kh"] == 3 and op["stride"] == 1
                and h_ == TILE and w_ == TILE and not op["stem"] and not op["aux_classes"]):
            if folded:  # same kernel + the upsample-add in its producer: reported apart
                fused_ms += ms_
                fused_n += 1
                continue
            dom_ms += ms_
            dom_n += 1
            dom_fl += fl
    flops = spec.conv_flops()
    achieved = dom_fl / (dom_ms * 1e-3) / 1e12
    traffic = None
    kname = "conv64x_kernel" if ctx.conv64_mode == 3 else "conv64_kernel"
    tsrc = None
    for tf in ("r2_traffic.json", "r1_traffic.json"):  # ncu --set full captures, newest first
        tpath = os.path.join(ROOT, "profiles", tf)
        if os.path.exists(tpath) and B == 32 and arm.precision == "f16":
            tj = json.load(open(tpath)).get("%s 256x256 64->64 3x3 batch 32" % kname)
            if tj:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                tsrc = tf
                break
    agg = flops / (all_ms * 1e-3) / 1e12
    enc = enc_fl / (enc_ms * 1e-3) / 1e12 if enc_ms > 0 else None
    return {"bound": "tensor", "achieved": achieved, "peak": peaks["tensor_burst"],
            "unit": "TFLOP/s", "frac": achieved / peaks["tensor_burst"], "traffic": traffic,
            "kernel": ("%s (even/odd N=128 formulation)" % kname if ctx.conv64_mode == 3 else kname)
            + ", 256x256 64->64 3x3, batch %d: %d launches/step, %.4f ms avg, %.1f GFLOP "
              "(algorithmic) per launch" % (B, dom_n, dom_ms / max(dom_n, 1), dom_fl / max(dom_n, 1) / 1e9),
            "peak_source": peaks["source"] + " bf16 cuBLAS burst (MEASURED_PEAKS.json bf16_tflops): "
                                             "the kernel is timed alone by per-op events",
            "frac_of_sustained": achieved / peaks["tensor_sustained"],
            "traffic_note": "DRAM bytes per launch from ncu --set full (profiles/%s); "
                            "algorithmic bytes 536.9 MB (fp16 in + out)" % tsrc,
            "fused_upadd_conv": ({"launches_per_step": fused_n, "ms_avg": fused_ms / fused_n,
                                  "note": "the same 64->64 kernel at 256x256 with skip + bilinear_x2(low) built in its "
                                          "producer (no stand-alone upadd pass, no 268 MB sum tensor): compare with "
                                          "one plain launch + one upadd pass"} if fused_n else None),
            "all_convs": {"achieved": agg, "frac": agg / peaks["tensor_burst"],
                          "launches_per_step": n_conv, "ms_per_step": all_ms,
                          "gflop_per_step": flops / 1e9},
            "encoder_stack": {"achieved": enc, "frac": enc / peaks["tensor_burst"] if enc else None,
                              "ms": enc_ms, "gflop": enc_fl / 1e9,
                              "note": "prep + stem + maxpool + layer1-4 (SURVEY 8a a3-a9), per-op events"},
            "step_breakdown_ms": {k: sum(m for kk, m in prof if kk == k)
                                  for k in sorted(set(k for k, _ in prof))}}, prof


def hbm_rooflines(arm, prof, peaks, barrier, reps=10):
    """HBM-bound stages: extract (uint8 tile -> fp16 NHWC stem input, prep_kernel) and
    the post-processing chain, against the measured copy bandwidth. Algorithmic bytes per tile:
    196,608 + 393,216 (extract); 2,359,296 + 786,432 (post-proc)."""
    import torch
    B = arm.B
    out = {}
    prep_ms = sum(m for k, m in prof if k == "prep")
    if prep_ms > 0:
        by = B * (196608 + 393216)
        a = by / (prep_ms * 1e-3) / 1e9
        out["extract"] = {"bound": "hbm", "achieved": a, "peak": peaks["hbm"], "unit": "GB/s",
                          "frac": a / peaks["hbm"], "traffic": None, "ms": prep_ms,
                          "kernel": "prep_kernel, %d tiles per launch (%.1f MB algorithmic)" % (B, by / 1e6)}
    if arm.pipe is not None:
        pipe = arm.pipe
        pstream = torch.cuda.ExternalStream(pipe.pctx.stream, device=arm.ctx.device)
        post = pipe.post[0]
        canvas = pipe.canvas_copy[(pipe.k - 1) % pipe.depth]  # canvas of the last submitted batch
        post.run(arm.plan, canvas_ptr=canvas)
        pipe.pctx.sync()
        l0 = pipe.pctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(pstream)
        for _ in range(reps):
            post.run(arm.plan, canvas_ptr=canvas)
        e1.record(pstream)
        pipe.pctx.sync()
        ms = e0.elapsed_time(e1) / reps
        by = B * (2359296 + 786432)
        a = by / (ms * 1e-3) / 1e9
        out["postproc"] = {"bound": "hbm", "achieved": a, "peak": peaks["hbm"], "unit": "GB/s",
                           "frac": a / peaks["hbm"], "traffic": None, "ms": ms,
                           "launches": (pipe.pctx.launch_count - l0) // reps,
                           "kernel": "nuclei + gland + lumen post-processing chain of one batch of %d "
                                     "images, run alone on the post stream (%.1f MB algorithmic)"
                                     % (B, by / 1e6)}
    return out


def parity_report(arms, sd, margs, n_tiles=4):
    """BASELINE.md section 5 in the same run: logit error of every timed precision vs the fp32
    reference forward and vs the net.half() restatement, and the end-to-end label agreement
    when each side uses its own forward. CPU side: oracle/ (checker only)."""
    from cerberus_b200 import synth
    from oracle import parity_report as pr
    tiles = synth.synthetic_tiles(n_tiles, TILE, TILE, seed=4242)
    ora = pr.oracle_side(sd, margs, tiles)
    rep = {"tiles": n_tiles, "tolerance_north_star": 1e-3,
           "reference": "oracle/net_oracle.py (fp32, pinned to the unmodified reference by "
                        "tests/golden/forward_*.npz) + oracle post-processing"}
    for arm in arms:
        lg, lab = pr.device_side(arm.eng, tiles)
        rep[arm.precision] = pr.compare(lg, lab, ora)
    return rep


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cerberus_b200 import synth
    from cerberus_b200.plan import PackedModel

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    margs = synth.model_args()
    # rank 0 folds + packs the checkpoint; the packed blob is NCCL-broadcast once
    from cerberus_b200.dist import broadcast_packed_model
    sd = synth.make_state_dict(seed=0) if rank == 0 else None
    model = PackedModel(sd, margs) if rank == 0 else None
    model = broadcast_packed_model(model, margs, rank, world, torch.device("cuda", local_rank))

    B = args.batch
    postproc = not args.no_postproc
    arm = Arm(model, local_rank, args.precision, B, postproc)

    n_in = 4
    host_batches = [synth.synthetic_tiles(B, TILE, TILE, seed=1000 * rank + i) for i in range(n_in)]
    pinned = [torch.from_numpy(b).pin_memory() for b in host_batches]
    dev_batches = [p.cuda() for p in pinned]
    host_np = [p.numpy() for p in pinned]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def over_ranks_max(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # Roofline evidence FIRST, on a GPU that has not been under load yet: the per-op events time each
    # kernel alone, and the denominator is the BURST tensor peak (MEASURED_PEAKS.json) - after a
    # second of continuous load the clocks of a power-capped B200 drop and the same profile reads
    # 10-15 % lower (clocks sampled during the profile are reported with it).
    peaks = load_peaks()
    roof, hbm, prof = None, None, None
    try:
        for i in range(args.warmup):
            arm.plan.run(device_ptr=dev_batches[i % n_in].data_ptr())
        arm.ctx.sync()
        psampler = ClockSampler(local_rank)
        psampler.start()
        roof, prof = conv_roofline(arm, dev_batches[0].data_ptr(), peaks)
        roof["clocks_during_profile"] = psampler.stop()
    except Exception as e:  # profiling is evidence, never a reason to lose the line
        roof = {"bound": "tensor", "achieved": None, "peak": peaks["tensor_burst"],
                "unit": "TFLOP/s", "frac": None, "traffic": None, "error": str(e)}

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches, host_enqueue_ms = time_resident(arm, dev_batches, args.steps, args.warmup, barrier)
    clocks = sampler.stop()
    per_rank_ms = [ms / args.steps]
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank_ms = [float(x.item()) / args.steps for x in allt]
    ms_max = over_ranks_max(ms)
    value = world * B * args.steps / (ms_max * 1e-3)

    dt, h2d, d2h = time_e2e(arm, host_np, args.steps, args.warmup, barrier)
    e2e_val = world * B * args.steps / over_ranks_max(dt)

    # forward only (BASELINE config 2) in the same run
    fwd = None
    if postproc:
        fms, _, _ = time_resident(arm, dev_batches, args.steps, args.warmup, barrier, forward_only=True)
        fms = over_ranks_max(fms)
        fv = world * B * args.steps / (fms * 1e-3)
        fwd = {"value": fv, "unit": "tiles/s", "ms_per_step": fms / args.steps,
               "tflops": fv * GFLOP_PER_TILE_6HEAD / 1e3, "note": "BASELINE config 2: forward only, "
               "batch resident in HBM, CUDA events"}

    try:
        if fwd is not None:
            fwd["frac_of_burst"] = fwd["tflops"] / world / peaks["tensor_burst"]
        if prof is not None:
            hbm = hbm_rooflines(arm, prof, peaks, barrier)
    except Exception as e:
        hbm = {"error": str(e)}

    ws_stats = None
    if arm.pipe is not None:
        pl = arm.pipe.pctx
        imgs = int(pl.lib.cerb_ctx_stat(pl.handle, b"ws_images"))
        ws_stats = {"images": imgs,
                    "exact_fallback_tie": int(pl.lib.cerb_ctx_stat(pl.handle, b"ws_tie_fallbacks")),
                    "exact_fallback_capacity": int(pl.lib.cerb_ctx_stat(pl.handle, b"ws_capacity_fallbacks")),
                    "note": "nuclei watershed, all steps incl. warm-up and e2e: images handled by the "
                            "component-parallel path vs redone by the exact whole-tile emulation"}
        if world > 1:  # which ranks were slowed by exact-watershed images
            mine = torch.tensor([float(ws_stats["exact_fallback_tie"])], device="cuda", dtype=torch.float64)
            allw = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allw, mine)
            ws_stats["exact_fallback_tie_per_rank"] = [int(x.item()) for x in allw]

    # ---- the other precision mode, timed in the same invocation (fewer steps)
    other = None
    arms = [arm]
    other_prec = "f16x2" if args.precision == "f16" else "f16"
    if not args.no_second_precision:
        arm2 = Arm(model, local_rank, other_prec, B, postproc)
        arms.append(arm2)
        k2, w2 = max(3, args.steps // 2), max(3, args.warmup)
        ms2, l2, _ = time_resident(arm2, dev_batches, k2, w2, barrier)
        ms2 = over_ranks_max(ms2)
        dt2, _, _ = time_e2e(arm2, host_np, k2, w2, barrier)
        dt2 = over_ranks_max(dt2)
        other = {"value": world * B * k2 / (ms2 * 1e-3), "unit": "tiles/s", "ms_per_step": ms2 / k2,
                 "steps": k2, "warmup": w2, "gpu_launches": int(l2),
                 "e2e": {"value": world * B * k2 / dt2, "unit": "tiles/s"},
                 "dtype": other_prec}

    parity = None
    cpu = None
    if rank == 0 and world == 1:
        if not args.no_parity:
            try:
                parity = parity_report(arms, sd, margs, args.parity_tiles)
            except Exception as e:
                parity = {"error": "%s: %s" % (type(e).__name__, e)}
        if not args.no_cpu_baseline:
            v, cores, sample, kind = cpu_reference_rate(args, args.cpu_seconds, args.ref_batch)
            cpu = {"value": v, "unit": "tiles/s", "cores": cores, "kind": kind, "sample": sample}

    if rank == 0:
        line = {
            "metric": "tiles/sec (256x256x3) end-to-end incl. post-proc" if postproc
            else "tiles/sec (256x256x3) forward only (BASELINE config 2; --no-postproc)",
            "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": workload_config(args, B, postproc),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "tiles/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "host_enqueue_ms_per_step": host_enqueue_ms,
            "ms_per_step_per_rank": per_rank_ms,
            "watershed": ws_stats,
            "roofline": roof,
            "roofline_hbm": hbm,
            "forward_only": fwd,
            other_prec: other,
            "parity": parity,
            "cpu_baseline": cpu,
            "tflops_forward": value * GFLOP_PER_TILE_6HEAD / 1e3,
        }
        print(json.dumps(line))
    for a in arms:
        a.close()
    if world > 1:
        dist.destroy_process_group()


def _tile_manager(tmp, tasks, precision, batch, in_shape, out_shape, postproc_list):
    """Model directory on disk -> InferManager, exactly as run_infer_tile.py builds it."""
    import yaml
    from cerberus_b200 import synth
    from cerberus_b200.infer.tile import InferManager
    mdir = os.path.join(tmp, "model_%s" % "_".join(tasks or ["all"]).replace("#", ""))
    if not os.path.exists(os.path.join(mdir, "weights.tar")):
        synth.write_model_dir(mdir, considered_tasks=tasks, seed=0)
    st = yaml.full_load(open(os.path.join(mdir, "settings.yml")))
    m = InferManager(checkpoint_path=os.path.join(mdir, "weights.tar"),
                     decoder_dict=st["dataset_kwargs"]["req_target_code"], model_args=st["model_kwargs"],
                     precision=precision)
    m.patch_input_shape, m.patch_output_shape, m.patch_output_overlap = in_shape, out_shape, 0
    m.batch_size, m.postproc_list = batch, postproc_list
    return m, st


def run_config1(args):
    """BASELINE config 1: ONE 256x256x3 synthetic image, ResNet34 encoder + Nuclei decoder head,
    batch 1, patch 256 -> 256, through the run_infer_tile.py plumbing: InferManager(model dir) ->
    patch grid -> device extract -> forward -> stitch -> __proc_nuclei -> instance tables. A step
    is one image. `value`: image uploaded per step, results left in HBM; `e2e`: host image in,
    host label map + type-less instance dict out (what the CLI hands to its .mat writer). The CPU
    reference (unmodified modules, oracle/_ref) runs the same image through _prepare_patching ->
    infer_step -> _post_process_patches beside it."""
    import tempfile

    import torch
    from cerberus_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.cuda.set_device(0)
    tmp = tempfile.mkdtemp(prefix="cerb_cfg1_")
    img = synth.synthetic_tiles(1, TILE, TILE, seed=0)[0]
    out = {}
    gflop = 54.891  # encoder + Nuclei decoder / head
    for prec in ("f16", "f16x2"):
        m, st = _tile_manager(tmp, ["Nuclei"], prec, 1, TILE, TILE, ["nuclei"])
        ctx = m.engine.ctx
        for _ in range(max(3, args.warmup)):
            res = m.process_image(img, "t")
        ctx.sync()
        l0 = ctx.launch_count
        sampler = ClockSampler(0)
        sampler.start()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            m.process_images([("t", img)], to_host=False)
        ctx.sync()
        dt_res = (time.perf_counter() - t0) / args.steps
        launches = ctx.launch_count - l0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = m.process_image(img, "t")
        dt_e2e = (time.perf_counter() - t0) / args.steps
        clocks = sampler.stop()
        # forward alone (device events): the part the tensor roofline applies to
        plan = m.engine.plan_for(1, TILE, TILE, TILE, TILE)
        stream = torch.cuda.ExternalStream(ctx.stream, device=0)
        dev = torch.from_numpy(img[None].copy()).cuda()
        plan.run(device_ptr=dev.data_ptr())
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            plan.run(device_ptr=dev.data_ptr())
        e1.record(stream)
        ctx.sync()
        fwd_ms = e0.elapsed_time(e1) / args.steps
        out[prec] = {"value": 1.0 / dt_res, "ms_per_step": 1e3 * dt_res, "e2e": 1.0 / dt_e2e,
                     "e2e_ms": 1e3 * dt_e2e, "launches": int(launches), "forward_ms": fwd_ms,
                     "forward_tflops": gflop / fwd_ms, "instances": len(res[3]["Nuclei"]), "clocks": clocks}
        m.release_device_buffers()
        m.engine.close()
    # parity of this very image: device label map vs reference on own forwards
    cpu = None
    parity = None
    from oracle import ref_runner
    if not args.no_cpu_baseline and ref_runner.available():
        torch.set_num_threads(host_cores())
        margs = synth.model_args(["Nuclei"])
        sd = synth.make_state_dict(["Nuclei"], seed=0)
        ref = ref_runner.ReferenceTilePath(sd, margs)
        code = dict(synth.DEFAULT_REQ_TARGET_CODE)
        r = ref.process_image(img, TILE, TILE, code, ["nuclei"])  # warm-up
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            r = ref.process_image(img, TILE, TILE, code, ["nuclei"])
            best = min(best, time.perf_counter() - t0)
        cpu = {"value": 1.0 / best, "unit": "tiles/s", "cores": torch.get_num_threads(), "kind": "reference",
               "sample": "the same image, best of 5: unmodified _prepare_patching -> infer_step (the "
                         "reference infers the single patch TWICE, infer/tile.py:90-103) -> "
                         "_post_process_patches incl. get_inst_info_dict; %.1f ms" % (1e3 * best)}
        m2, _ = _tile_manager(tmp, ["Nuclei"], "f16x2", 1, TILE, TILE, ["nuclei"])
        mine = m2.process_image(img, "t")
        a, b = np.asarray(mine[2]["Nuclei"]).astype(np.int64), np.asarray(r[2]["Nuclei"]).astype(np.int64)
        parity = {"f16x2_label_map_identical_to_reference": bool(np.array_equal(a, b)),
                  "pixel_mismatch": float((a != b).mean()), "instances": [int(a.max()), int(b.max())],
                  "instance_ids_equal": list(mine[3]["Nuclei"].keys()) == list(r[3]["Nuclei"].keys())}
        m2.release_device_buffers()
        m2.engine.close()
    peaks = load_peaks()
    main = out[args.precision]
    line = {
        "metric": "tiles/sec (256x256x3) end-to-end incl. post-proc", "value": main["value"],
        "unit": "tiles/s", "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "BASELINE config 1: single 256x256x3 synthetic image, ResNet34 encoder + 1 "
                               "nuclei-seg decoder head, batch=1, run_infer_tile.py plumbing "
                               "(InferManager.process_image, patch 256 -> 256) incl. nuclei post-processing "
                               "and instance tables", "tile": [TILE, TILE, 3], "batch_per_gpu": 1, "heads": 1,
                   "precision": args.precision,
                   "l2": "latency-bound single-image case: the working set fits L2; nothing to flush"},
        "clocks": main["clocks"],
        "e2e": {"value": main["e2e"], "unit": "tiles/s", "h2d_bytes_per_step": int(img.nbytes),
                "d2h_bytes_per_step": int(TILE * TILE * 4)},
        "gpu_launches": main["launches"],
        "roofline": {"bound": "tensor", "achieved": main["forward_tflops"], "peak": peaks["tensor_burst"],
                     "unit": "TFLOP/s", "frac": main["forward_tflops"] / peaks["tensor_burst"], "traffic": None,
                     "kernel": "whole forward of ONE tile (54.891 GFLOP, %.3f ms): 64 regions of 16x16 px "
                               "on 148 SMs - launch / latency bound by construction" % main["forward_ms"]},
        "per_precision": out, "parity": parity, "cpu_baseline": cpu,
    }
    print(json.dumps(line))


def run_tile448(args):
    """The CLI path at its defaults: a directory of PNGs through InferManager.process_file_list
    (PNG decode -> 448/144 patches batched ACROSS files -> forward -> stitch -> post-processing ->
    instance tables -> .mat / overlay files), one rank per GPU, files sharded over ranks."""
    import shutil
    import tempfile

    import cv2
    import torch
    import torch.distributed as dist
    from cerberus_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    tmp = os.environ.get("CERB_BENCH_TMP") or os.path.join(tempfile.gettempdir(), "cerb_tile448")
    n_files, size = int(os.environ.get("CERB_TILE448_FILES", "32")) * world, 1000
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)
        os.makedirs(tmp + "/in")
        for i in range(n_files):
            img = synth.synthetic_tiles(1, 1008, 1008, seed=500 + i)[0][:size, :size]
            cv2.imwrite("%s/in/img%03d.png" % (tmp, i), cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
        synth.write_model_dir(tmp + "/model_all", seed=0)
    if world > 1:
        dist.barrier()
    import yaml
    from cerberus_b200.infer.tile import InferManager
    st = yaml.full_load(open(tmp + "/model_all/settings.yml"))
    m = InferManager(checkpoint_path=tmp + "/model_all/weights.tar",
                     decoder_dict=st["dataset_kwargs"]["req_target_code"], model_args=st["model_kwargs"],
                     precision=args.precision, device=local_rank)
    nw = int(os.environ.get("CERB_TILE448_WORKERS", "4"))
    m.time_stages = os.environ.get("CERB_TILE448_STAGES") == "1"
    run_args = {"nr_inference_workers": nw, "nr_post_proc_workers": nw, "batch_size": args.batch,
                "input_dir": tmp + "/in", "output_dir": tmp + "/out_warm", "patch_input_shape": 448,
                "patch_output_shape": 144, "patch_output_overlap": 0,
                "postproc_list": ["gland", "lumen", "nuclei", "patch-class"]}
    os.makedirs(run_args["output_dir"], exist_ok=True)
    import io
    from contextlib import redirect_stdout
    with redirect_stdout(io.StringIO()):
        m.process_file_list(dict(run_args))  # warm-up: plans, buffers
        if world > 1:
            dist.barrier()
        m.nr_patches_inferred = 0
        m.stage_seconds = {"extract": 0.0, "forward": 0.0, "finish": 0.0}
        m.finish_seconds = {}
        run_args["output_dir"] = tmp + "/out"
        os.makedirs(run_args["output_dir"], exist_ok=True)
        l0 = m.engine.ctx.launch_count
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.process_file_list(dict(run_args))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt, float(m.nr_patches_inferred)], device="cuda", dtype=torch.float64)
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt, patches = float(mx[0].item()), float(t[1].item())
    else:
        patches = float(t[1].item())
    if rank == 0:
        n_out = len([f for f in os.listdir(tmp + "/out/nuclei_mat") if f.endswith(".mat")])
        px = n_files * size * size
        line = {"metric": "tiles/sec (256x256x3) end-to-end incl. post-proc",
                "value": px / 65536.0 / dt, "unit": "tiles/s", "n_gpus": world, "steps": 1, "warmup": 1,
                "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic",
                "config": {"workload": "CLI defaults: %d synthetic %dx%d PNGs through "
                                       "InferManager.process_file_list, patch 448 -> 144 (each output pixel "
                                       "costs 9.7x the network work of the in == out bench), batch %d filled "
                                       "across files, six heads + post-processing + instance tables + .mat / "
                                       "overlay files; value = image pixels / 65536 per second"
                                       % (n_files, size, size, args.batch),
                           "files": n_files, "patches_448_inferred": int(patches),
                           "patches_per_s": patches / dt, "files_per_s": n_files / dt,
                           "precision": args.precision, "mat_files_written": n_out,
                           "loader_and_writer_threads": nw,
                           "rank0_stage_seconds": {k: round(v, 3) for k, v in m.stage_seconds.items()},
                           "rank0_finish_seconds": {k: round(v, 3) for k, v in getattr(m, "finish_seconds", {}).items()}},
                "e2e": {"value": px / 65536.0 / dt, "unit": "tiles/s",
                        "h2d_bytes_per_step": int(px * 3), "d2h_bytes_per_step": int(px * 4 * 6)},
                "gpu_launches": int(m.engine.ctx.launch_count - l0),
                "tflops_forward": patches / dt * 370.953 / 1e3}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_torch_gpu(args):
    """EXTRA CONTEXT ("second baseline"; not an arm the driver runs): the
    UNMODIFIED reference network (oracle/_ref: models.net_desc.create_model, random-init synthetic
    checkpoint) executed by PyTorch / cuDNN on cuda:0 - `.half()`, channels_last,
    cudnn.benchmark - forward only on the bench batch, timed with CUDA events. Compare with
    `forward_only` of the default line. Nothing of this repository's engine is on this path."""
    import torch
    from cerberus_b200 import synth
    from oracle import ref_runner
    if not ref_runner.available():
        print(json.dumps({"impl": "torch_gpu_context", "unavailable": "oracle/_ref is not staged"}))
        return
    margs = synth.model_args()
    ref = ref_runner.ReferenceTilePath(synth.make_state_dict(seed=0), margs)
    torch.backends.cudnn.benchmark = True
    net = ref.net.half().cuda().to(memory_format=torch.channels_last)
    tiles = synth.synthetic_tiles(args.batch, TILE, TILE, seed=123)
    x = torch.from_numpy(tiles).cuda().permute(0, 3, 1, 2).half().contiguous(memory_format=torch.channels_last)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            net(x)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            net(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({
        "impl": "torch_gpu_context", "metric": "tiles/sec (256x256x3) forward only",
        "value": args.batch / (ms * 1e-3), "unit": "tiles/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": "batch=%d synthetic 256x256x3 tiles, forward of the unmodified reference "
                               "network (six heads) under PyTorch %s / cuDNN: net.half(), channels_last, "
                               "cudnn.benchmark, no post-processing; the reference's own forward syncs once "
                               "per call (net_desc.py:171)" % (args.batch, torch.__version__)},
        "note": "context only: compare with `forward_only.ms_per_step` of the default bench line"}))


def main():
    args = parse()
    if args.impl == "torch-gpu":
        run_torch_gpu(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.config == "1":
        run_config1(args)
    elif args.config == "tile448":
        run_tile448(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
