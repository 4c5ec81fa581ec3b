"""CPU: the forward oracle (oracle/net_oracle.py) against golden vectors produced by the
UNMODIFIED reference NetDesc / infer_step (oracle/gen_golden.py). This is what pins the
oracle; the GPU tests then compare the CUDA path with the oracle."""
import os

import numpy as np
import pytest
import torch

from cerberus_b200 import synth
from oracle import net_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["six_256", "six_448", "nuclei_256", "r18_256"])
def test_oracle_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, "forward_%s.npz" % name))
    tasks = [str(t) for t in g["tasks"]]
    backbone = str(g["backbone"]) if "backbone" in g else "resnet34"  # r18_*: resnet18 encoder
    args = synth.model_args(tasks, backbone=backbone)
    sd = synth.make_state_dict(tasks, seed=int(g["ckpt_seed"]), backbone=backbone)
    chk = np.array([float(sd["backbone.layer4.%d.bn2.running_var" % (synth.BACKBONE_BLOCKS[backbone][3] - 1)].double().sum()),
                    float(sd["backbone.layer1.0.bn1.running_mean"].double().sum())])
    assert np.allclose(chk, g["sd_check"], rtol=1e-4), "synthetic checkpoint drifted"
    n, size, out = int(g["n"]), int(g["size"]), int(g["out"])
    tiles = synth.synthetic_tiles(n, size, size, seed=int(g["tile_seed"]))
    step, logits = net_oracle.infer_step(sd, tiles, out, args["decoder_kwargs"], tasks)
    for k, v in logits.items():
        v = v.numpy()
        sub = v[..., 3::8, 3::8] if v.shape[-1] > 1 else v
        assert np.abs(sub - g["logits_sub/" + k]).max() <= 2e-5, k
        assert np.allclose(np.abs(v).mean(axis=(0, 2, 3)), g["logits_absmean/" + k], rtol=1e-4)
    assert len(step) == n
    for k in step[0]:
        full = np.stack([s[k] for s in step])
        assert str(step[0][k].dtype) == str(g["step_dtype/" + k])
        assert tuple(step[0][k].shape) == tuple(g["step_shape/" + k])
        if ("step/" + k) in g:
            assert (full.astype(np.uint8) != g["step/" + k]).mean() <= 1e-4, k
        else:
            assert np.abs(full[:, 2::4, 2::4, :] - g["step_sub/" + k]).max() <= 1e-5, k
