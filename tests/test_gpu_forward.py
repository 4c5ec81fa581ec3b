"""GPU parity of the whole forward / run_step against the oracle (oracle/net_oracle.py, a CPU
fp32 restatement pinned to the reference by tests/golden/forward_*.npz) and against the
golden vectors themselves.

Tolerance (BASELINE.json north_star): 1e-3 max-abs on float head logits. It is met in the
split-precision mode (CERB_PREC_F16X2); the fp16 throughput mode is measured and bounded
loosely here, its error is reported by bench.py / DESIGN.md (landmine 2)."""
import os

import numpy as np
import pytest
import torch

from cerberus_b200 import synth
from cerberus_b200.engine import Engine
from oracle import net_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOGIT_TOL = 1e-3


def _oracle_logits(sd, args, tiles):
    x = torch.from_numpy(tiles).float().permute(0, 3, 1, 2).contiguous()
    out = net_oracle.forward(sd, x, args["decoder_kwargs"], args["considered_tasks"])
    return {k: v.permute(0, 2, 3, 1).contiguous().numpy() for k, v in out.items()}


@pytest.mark.parametrize("name", ["six_256", "six_448", "nuclei_256", "r18_256"])
def test_logits_match_reference_golden_f16x2(name, built_lib):
    g = np.load(os.path.join(GOLD, "forward_%s.npz" % name))
    tasks = [str(t) for t in g["tasks"]]
    backbone = str(g["backbone"]) if "backbone" in g else "resnet34"  # r18_*: resnet18 encoder
    args = synth.model_args(tasks, backbone=backbone)
    sd = synth.make_state_dict(tasks, seed=int(g["ckpt_seed"]), backbone=backbone)
    n, size, out = int(g["n"]), int(g["size"]), int(g["out"])
    tiles = synth.synthetic_tiles(n, size, size, seed=int(g["tile_seed"]))
    eng = Engine(sd, args, precision="f16x2")
    plan = eng.plan_for(n, size, size, out, out, want_logits=True)
    plan.run(tiles)
    got = plan.read_logits()
    ref = _oracle_logits(sd, args, tiles)
    for k in ref:
        gk = got[k].reshape(ref[k].shape)
        err = float(np.abs(gk - ref[k]).max())
        assert err <= LOGIT_TOL, "%s: max-abs logit error %g vs oracle" % (k, err)
        # the reference's own numbers (sub-sampled golden)
        gold = g["logits_sub/" + k]  # NCHW, [..., 3::8, 3::8]
        mine = np.transpose(gk, (0, 3, 1, 2))
        mine = mine[..., 3::8, 3::8] if mine.shape[-1] > 1 else mine
        err_g = float(np.abs(mine - gold).max())
        assert err_g <= LOGIT_TOL, "%s: max-abs logit error %g vs reference golden" % (k, err_g)
    # step outputs (models/run_desc.py:480-502 contract)
    step = eng.run_step(torch.from_numpy(tiles), out)
    assert len(step) == n
    ref_step, _ = net_oracle.infer_step(sd, tiles, out, args["decoder_kwargs"], tasks)
    for i in range(n):
        assert list(step[i].keys()) == list(ref_step[i].keys())
        for k, v in step[i].items():
            r = ref_step[i][k]
            assert v.shape == r.shape and v.dtype == r.dtype, (k, v.shape, v.dtype, r.shape, r.dtype)
            if k.endswith("-INST"):
                assert float(np.abs(v - r).max()) <= LOGIT_TOL
            else:
                # argmax maps: identical except where the two top logits are within tolerance
                assert float((v != r).mean()) <= 2e-3, "%s mismatch %g" % (k, float((v != r).mean()))
    eng.close()


# fp16 throughput mode (the precision bench.py times). Stated bounds, not the 1e-3 gate:
#   * max-abs logit error vs the fp32 reference forward <= 0.12 (CPU emulation of the mode,
#     tools/precision_study.py: 0.055 with error-diffused weight rounding; 0.23 with plain
#     round-to-nearest weights; `net.half()` itself sits 0.2 from fp32) and no further from fp32
#     than the net.half() restatement is;
#   * end-to-end, each side using its own forward: foreground pixel
#     mismatch of the instance maps <= 1 % per tissue.
F16_LOGIT_BOUND = 0.12
F16_FG_MISMATCH_BOUND = 0.01


def test_fp16_mode_error_is_bounded(built_lib, six_head_sd):
    from oracle import parity_report as pr
    args = synth.model_args()
    tiles = synth.synthetic_tiles(3, 256, 256, seed=7)
    ora = pr.oracle_side(six_head_sd, args, tiles)
    eng = Engine(six_head_sd, args, precision="f16")
    logits, labels = pr.device_side(eng, tiles)
    rep = pr.compare(logits, labels, ora)
    print("fp16 mode parity report:", rep)
    assert rep["logits_max_abs_vs_fp32"] <= F16_LOGIT_BOUND
    assert rep["logits_max_abs_vs_fp32"] <= rep["half_reference_max_abs_vs_fp32"]
    for t, r in rep["labels_own_forward"].items():
        assert r["foreground_pixel_mismatch"] <= F16_FG_MISMATCH_BOUND, (t, r)
    eng.close()


def test_split_mode_labels_match_reference_end_to_end(built_lib, six_head_sd):
    """Parity mode, each side on its own forward: logits within 1e-3 and (because the floats
    differ by < 1e-3 only) label maps equal except where a probability sits within the logit
    tolerance of a post-processing threshold."""
    from oracle import parity_report as pr
    args = synth.model_args()
    tiles = synth.synthetic_tiles(3, 256, 256, seed=7)
    ora = pr.oracle_side(six_head_sd, args, tiles, with_half=False)
    eng = Engine(six_head_sd, args, precision="f16x2")
    logits, labels = pr.device_side(eng, tiles)
    rep = pr.compare(logits, labels, ora)
    print("split mode parity report:", rep)
    assert rep["logits_max_abs_vs_fp32"] <= LOGIT_TOL
    for t, r in rep["labels_own_forward"].items():
        assert r["foreground_pixel_mismatch"] <= 1e-3, (t, r)
    eng.close()


def test_plan_is_deterministic(built_lib, six_head_sd):
    args = synth.model_args()
    tiles = synth.synthetic_tiles(3, 256, 256, seed=11)
    eng = Engine(six_head_sd, args, precision="f16")
    plan = eng.plan_for(3, 256, 256, 256, 256)
    plan.run(tiles)
    a = plan.read_canvas().copy()
    plan.run(tiles)
    b = plan.read_canvas()
    assert np.array_equal(a, b)
    eng.close()


def test_device_pipeline_matches_oracle_pipeline(built_lib, six_head_sd):
    """BASELINE config 3: forward + on-device post-processing for a batch of independent tiles.
    The label maps must be bit-exact with the oracle pipeline (infer/tile.py:116-191 restated)
    fed with the SAME float canvas (thresholds make labels discontinuous in the
    floats, so exactness is gated on identical post-processing inputs)."""
    from cerberus_b200.pipeline import DevicePostProc
    from cerberus_b200.engine import canvas_to_step_outputs
    from oracle import pipeline_oracle
    args = synth.model_args()
    n = 6
    tiles = synth.synthetic_tiles(n, 256, 256, seed=21)
    eng = Engine(six_head_sd, args, precision="f16")
    plan = eng.plan_for(n, 256, 256, 256, 256)
    plan.run(tiles)
    post = DevicePostProc(eng.ctx, eng.model, n, 256, 256)
    labels = post.run_to_host(plan)
    step = canvas_to_step_outputs(plan.read_canvas(), eng.model)
    ref = pipeline_oracle.postprocess_step(step, args)
    total_inst = 0
    for i in range(n):
        for t in ("Nuclei", "Gland", "Lumen"):
            assert np.array_equal(labels[t][i].astype(np.int64), ref[i][t].astype(np.int64)), (t, i)
            total_inst += int(ref[i][t].max())
    assert total_inst > 100  # the synthetic tiles give the post-processing real work
    # whole-tile exact emulation (ws_mode 1) agrees with the component-parallel fast path
    eng.ctx.set_option("ws_mode", 1)
    labels2 = {t: v.copy() for t, v in post.run_to_host(plan).items()}
    for t in labels2:
        assert np.array_equal(labels2[t], labels[t])
    post.close()
    eng.close()


def test_tile_pipeline_streams_batches_in_order(built_lib, six_head_sd):
    """TilePipeline (overlapped H2D / compute / D2H) returns, one submit late, exactly what the
    synchronous path produces for each batch."""
    from cerberus_b200.pipeline import DevicePostProc, TilePipeline
    args = synth.model_args()
    n = 4
    batches = [synth.synthetic_tiles(n, 256, 256, seed=30 + i) for i in range(4)]
    eng = Engine(six_head_sd, args, precision="f16")
    plan = eng.plan_for(n, 256, 256, 256, 256)
    post = DevicePostProc(eng.ctx, eng.model, n, 256, 256)
    want = []
    for b in batches:
        plan.run(b)
        want.append({t: v.copy() for t, v in post.run_to_host(plan).items()})
    post.close()
    pipe = TilePipeline(eng, n, 256, 256)
    got = []
    for b in batches:
        r = pipe.submit(b)
        if r is not None:
            got.append({t: v.copy() for t, v in r.items()})
    for r in pipe.flush():
        got.append({t: v.copy() for t, v in r.items()})
    assert len(got) == len(want)
    for g, w_ in zip(got, want):
        for t in w_:
            assert np.array_equal(g[t], w_[t]), t
    pipe.close()
    eng.close()


@pytest.mark.parametrize("variant", ["head_in_1x1", "head_behind_conv64"])
def test_fused_head_on_tensor_core_matches_fp32_head(built_lib, six_head_sd, variant):
    """fp16 mode: the fused heads run the 1x1 96->C (and, behind the last decoder conv, also the
    hidden 1x1 64->96) on the tensor core with fp16-rounded hidden tiles / head weights and fp32
    accumulation. They must agree with the unfused path (tensors in HBM as fp16, CERB_OP_HEAD in
    fp32 FMAs) to fp16-rounding accuracy, on every head, for a batch that leaves CTAs idle."""
    from cerberus_b200.engine import Context, ForwardPlan
    from cerberus_b200.plan import PackedModel, PlanSpec
    args = synth.model_args()
    model = PackedModel(six_head_sd, args)
    tiles = synth.synthetic_tiles(3, 256, 256, seed=5)
    ctx = Context(0, "f16")
    if variant == "head_behind_conv64":
        ctx.set_option("conv64_mode", 1)  # the fused tail lives in conv64.cu (halo layout 1)
    out = []
    for fuse in (True, False):
        spec = PlanSpec(model, 3, 256, 256, 256, 256, want_logits=True,
                        fuse_head=fuse and variant == "head_in_1x1",
                        fuse_tail=fuse and variant == "head_behind_conv64")
        plan = ForwardPlan(ctx, model, 3, 256, 256, 256, 256, spec=spec)
        plan.run(tiles)
        out.append(({k: v.copy() for k, v in plan.read_logits().items()}, plan.read_canvas().copy()))
        plan.close()
    ctx.close()
    (lg_f, cv_f), (lg_u, cv_u) = out
    for k in lg_u:
        if k == "Patch-Class":
            continue
        err = float(np.abs(lg_f[k] - lg_u[k]).max())
        scale = float(np.abs(lg_u[k]).max())
        print("%s %s: fused-vs-unfused max-abs %.4g (logit scale %.3g)" % (variant, k, err, scale))
        assert err <= 4e-3 * max(scale, 1.0), (k, err, scale)
    # probabilities written to the canvas agree as well
    inst = [i for k, (a, b) in model.idx_dict.items() if k.endswith("-INST") for i in range(a, b)]
    assert float(np.abs(cv_f[..., inst] - cv_u[..., inst]).max()) <= 5e-3


def test_fused_head_crops_and_partial_tiles(built_lib, six_head_sd):
    """448 -> 144 centre crop (reference defaults, run_infer_tile.py:1-23) through the fused tail:
    448 is not a multiple of the 16-row tile, so border tiles are partial."""
    args = synth.model_args()
    tiles = synth.synthetic_tiles(1, 448, 448, seed=3)
    outs = []
    for env in ("1", "0"):
        import os
        os.environ["CERB_FUSE_TAIL"] = env
        os.environ["CERB_CONV64_MODE"] = "1" if env == "1" else "3"
        eng = Engine(six_head_sd, args, precision="f16")
        plan = eng.plan_for(1, 448, 448, 144, 144)
        plan.run(tiles)
        outs.append(plan.read_canvas().copy())
        eng.close()
    os.environ.pop("CERB_FUSE_TAIL", None)
    os.environ.pop("CERB_CONV64_MODE", None)
    inst = [i for k, (a, b) in eng.model.idx_dict.items() if k.endswith("-INST") for i in range(a, b)]
    assert outs[0].shape == outs[1].shape == (1, 144, 144, eng.model.canvas_c)
    # the two runs also use different 64->64 kernels (accumulation order): fp16-mode noise
    assert float(np.abs(outs[0][..., inst] - outs[1][..., inst]).max()) <= 2e-2


def test_model_api_forward_equals_plan_api(built_lib, six_head_sd):
    """cerb_model_create + cerb_forward (op graph built inside the library, csrc/model.cu) vs the
    plan-level API fed by cerberus_b200/plan.py: same kernels, bit-identical canvas; 448 -> 144
    crop as well."""
    from cerberus_b200.engine import CModel
    args = synth.model_args()
    eng = Engine(six_head_sd, args, precision="f16")
    cm = CModel(eng.ctx, eng.model)
    for (n, size, out, seed) in ((3, 256, 256, 5), (1, 448, 144, 6)):
        tiles = synth.synthetic_tiles(n, size, size, seed=seed)
        plan = eng.plan_for(n, size, size, out, out)
        plan.run(tiles)
        want = plan.read_canvas().copy()
        got = cm.forward(tiles, out, out, eng.model.canvas_c)
        assert got.shape == want.shape and np.array_equal(got, want)
    cm.close()
    eng.close()


def test_level_sync_order_gives_the_same_canvas(built_lib, six_head_sd, monkeypatch):
    """Opt-in decoder order (level by level, ONE grouped UPADD per level that reads the shared
    skip tensor once): same arithmetic per output, so the canvas must be bit-identical."""
    args = synth.model_args()
    tiles = synth.synthetic_tiles(3, 256, 256, seed=17)
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("CERB_LEVEL_SYNC", flag)
        eng = Engine(six_head_sd, args, precision="f16")
        plan = eng.plan_for(3, 256, 256, 256, 256)
        assert any(op["kind"] == 4 and op["cout"] > 1 for op in plan.spec.ops) == (flag == "1")
        plan.run(tiles)
        outs.append(plan.read_canvas().copy())
        eng.close()
    assert np.array_equal(outs[0], outs[1])
