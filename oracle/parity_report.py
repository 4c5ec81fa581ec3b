"""ORACLE — test infrastructure only (imported by bench.py's `parity` block and by tests/).

Measures, for one precision mode of the CUDA path, what BASELINE.md section 5 asks to be
reported next to every throughput number:
  * max-abs head-logit error vs the fp32 reference forward (oracle/net_oracle.forward, pinned
    to the unmodified reference by tests/golden/forward_*.npz) and vs the `net.half()`
    restatement (net_oracle.forward_half);
  * end-to-end label agreement "when each side uses its own forward" (SURVEY.md 8d config 3):
    CUDA forward + device post-processing vs fp32 oracle forward (models/run_desc.py:439-502)
    + oracle post-processing (loader/postproc.py:352-407, infer/tile.py:187-191).
"""
from collections import OrderedDict

import numpy as np
import torch

from oracle import net_oracle, pipeline_oracle


def oracle_side(sd, margs, tiles, with_half=True):
    """CPU: fp32 logits, half logits, step outputs and label maps of the reference path."""
    dk, tasks = margs["decoder_kwargs"], margs["considered_tasks"]
    step, logits = net_oracle.infer_step(sd, tiles, tiles.shape[1], dk, tasks)
    out = {"logits": {k: v.permute(0, 2, 3, 1).contiguous().numpy() for k, v in logits.items()},
           "labels": pipeline_oracle.postprocess_step(step, margs)}
    if with_half:
        x = torch.from_numpy(np.asarray(tiles)).float().permute(0, 3, 1, 2).contiguous()
        half = net_oracle.forward_half(sd, x, dk, tasks)
        out["logits_half"] = {k: v.permute(0, 2, 3, 1).contiguous().numpy() for k, v in half.items()}
        out["half_vs_fp32"] = max(float(np.abs(out["logits_half"][k] - out["logits"][k]).max())
                                  for k in out["logits"])
    return out


def device_side(engine, tiles):
    """CUDA: logits + label maps of one batch through the engine (want_logits plan + device
    post-processing)."""
    from cerberus_b200.pipeline import DevicePostProc
    n, h, w, _ = tiles.shape
    plan = engine.plan_for(n, h, w, h, w, want_logits=True)
    plan.run(tiles)
    logits = plan.read_logits()
    post = DevicePostProc(engine.ctx, engine.model, n, h, w)
    labels = {t: v.copy() for t, v in post.run_to_host(plan).items()}
    post.close()
    return logits, labels


def compare(dev_logits, dev_labels, ora):
    """-> dict for the bench line / test assertions."""
    rep = OrderedDict()
    ref = ora["logits"]
    rep["logits_max_abs_vs_fp32"] = max(
        float(np.abs(dev_logits[k].reshape(ref[k].shape) - ref[k]).max()) for k in ref)
    rep["logits_rms_vs_fp32"] = float(np.sqrt(np.mean(
        [float(((dev_logits[k].reshape(ref[k].shape) - ref[k]) ** 2).mean()) for k in ref])))
    if "logits_half" in ora:
        rh = ora["logits_half"]
        rep["logits_max_abs_vs_half"] = max(
            float(np.abs(dev_logits[k].reshape(rh[k].shape) - rh[k]).max()) for k in rh)
        rep["half_reference_max_abs_vs_fp32"] = ora["half_vs_fp32"]
    lab = OrderedDict()
    n = len(ora["labels"])
    for t in dev_labels:
        fg_mis, id_mis, d_inst, n_ref = 0.0, 0.0, 0, 0
        for i in range(n):
            mine = np.asarray(dev_labels[t][i]).astype(np.int64)
            theirs = np.asarray(ora["labels"][i][t]).astype(np.int64)
            fg_mis += float(((mine > 0) != (theirs > 0)).mean())
            id_mis += float((mine != theirs).mean())
            d_inst += abs(int(mine.max()) - int(theirs.max()))
            n_ref += int(theirs.max())
        lab[t] = {"foreground_pixel_mismatch": fg_mis / n, "label_id_pixel_mismatch": id_mis / n,
                  "instance_count_abs_delta": d_inst, "instances_reference": n_ref}
    rep["labels_own_forward"] = lab
    return rep
