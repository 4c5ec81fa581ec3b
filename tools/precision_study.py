#!/usr/bin/env python
"""CPU emulation of the device numerics: which rounding step costs how much logit error.

Runs the six-head forward of oracle/net_oracle.py with BN folded the way cerberus_b200/pack.py
folds it, and rounds weights / stored activations the way a given device mode would:

  w: f32 | f16 | f16x2(hi+lo)       a: f32 | f16 | f16x2     (accumulation is always fp32)

Prints max-abs / rms logit error per mode vs the fp32 reference forward. Test tooling
(imports oracle/), not product code.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cerberus_b200 import synth  # noqa: E402
from oracle import net_oracle  # noqa: E402


def qw(t, mode):
    """Weights are pre-scaled by a power of two before rounding (pack.weight_shift)."""
    if mode == "f32":
        return t
    if mode == "f16ed":
        return qw_ed(t)
    m = float(t.abs().max())
    sh = int(np.clip(np.floor(np.log2(512.0 / m)), -60, 60)) if m > 0 else 0
    return torch.ldexp(q(torch.ldexp(t, torch.tensor(sh)), mode), torch.tensor(-sh))


def qw_ed(t):
    """fp16 rounding with error diffusion along K (per output channel): the running sum of the
    rounding errors stays below one ulp, so the error seen by a CONSTANT input vanishes."""
    m = float(t.abs().max())
    sh = int(np.clip(np.floor(np.log2(512.0 / m)), -60, 60)) if m > 0 else 0
    w = np.ldexp(t.double().numpy(), sh)
    o = w.shape[0]
    flat = w.reshape(o, -1)  # order: (i, kh, kw)
    out = np.empty_like(flat)
    carry = np.zeros(o)
    for k in range(flat.shape[1]):
        want = flat[:, k] + carry
        r = want.astype(np.float16).astype(np.float64)
        carry = want - r
        out[:, k] = r
    return torch.from_numpy(np.ldexp(out.reshape(w.shape), -sh)).float()


def q(t, mode):
    if mode == "f32":
        return t
    if mode == "f16":
        return t.half().float()
    if mode == "bf16":
        return t.bfloat16().float()
    if mode == "f16x2":
        hi = t.half().float()
        return hi + (t - hi).half().float()
    raise ValueError(mode)


def fold(w, b, bn, eps=1e-5):
    s = bn["weight"].double() / torch.sqrt(bn["running_var"].double() + eps)
    wf = w.double() * s[:, None, None, None]
    b0 = b.double() if b is not None else torch.zeros(w.shape[0], dtype=torch.float64)
    bf = (b0 - bn["running_mean"].double()) * s + bn["bias"].double()
    return wf.float(), bf.float()


def bn_of(sd, p):
    return {k: sd[p + "." + k] for k in ("weight", "bias", "running_mean", "running_var")}


def forward_emul(sd, imgs, decoder_kwargs, tasks, wm="f16", am="f16", res_mode=None, skip_mode=None):
    """res_mode: precision of the tensors that feed residual adds (block outputs);
    skip_mode: precision of encoder features read by the decoders. Default = am."""
    res_mode = res_mode or am
    skip_mode = skip_mode or am

    def conv(x, w, b, stride=1, pad=1, relu=True, res=None, out_mode=None):
        y = F.conv2d(x, qw(w, wm), None, stride, pad) + b[None, :, None, None]
        if res is not None:
            y = y + res
        if relu:
            y = F.relu(y)
        return q(y, out_mode or am)

    with torch.no_grad():
        w, b = fold(sd["backbone.conv1.weight"] / 255.0, None, bn_of(sd, "backbone.bn1"))
        x0 = x = conv(imgs, w, b, 1, 3)
        x = F.max_pool2d(x, 3, 2, 1)
        feats = [x0]
        xr = x  # residual-precision copy
        for li, nb in enumerate(net_oracle.BLOCKS, start=1):
            for bi in range(nb):
                p = "backbone.layer%d.%d" % (li, bi)
                stride = 2 if (li > 1 and bi == 0) else 1
                w1, b1 = fold(sd[p + ".conv1.weight"], None, bn_of(sd, p + ".bn1"))
                w2, b2 = fold(sd[p + ".conv2.weight"], None, bn_of(sd, p + ".bn2"))
                mid = conv(x, w1, b1, stride, 1)
                if (p + ".downsample.0.weight") in sd:
                    wd, bd = fold(sd[p + ".downsample.0.weight"], None, bn_of(sd, p + ".downsample.1"))
                    idn = conv(x, wd, bd, stride, 0, relu=False, out_mode=res_mode)
                else:
                    idn = xr
                y = F.conv2d(mid, qw(w2, wm), None, 1, 1) + b2[None, :, None, None] + idn
                y = F.relu(y)
                xr = q(y, res_mode)
                x = q(y, am)
            feats.append(q(xr, skip_mode) if skip_mode != res_mode else xr)
        x0s = feats[0]
        bottom = feats[-1]
        f4 = conv(q(bottom, am), sd["conv_map.weight"], torch.zeros(256), 1, 0, relu=False)
        outputs = {}
        for d, heads in decoder_kwargs.items():
            if d not in tasks or d == "Patch-Class":
                continue
            prev = f4
            fl = [x0s, feats[1], feats[2], feats[3]]
            for idx in range(1, 5):
                up = F.interpolate(prev, scale_factor=2, mode="bilinear", align_corners=False)
                prev = q(fl[-idx] + up, am)
                for cv in range(2):
                    p = "decoder_head.%s.%d.block.%d" % (d, idx - 1, cv)
                    w, b = fold(sd[p + ".conv.weight"], sd[p + ".conv.bias"], bn_of(sd, p + ".bn"))
                    prev = conv(prev, w, b, 1, 1)
            for clf in heads:
                p = "output_head.%s.%s.x" % (d, clf)
                w, b = fold(sd[p + ".0.block.0.conv.weight"], sd[p + ".0.block.0.conv.bias"],
                            bn_of(sd, p + ".0.block.0.bn"))
                y = conv(prev, w, b, 1, 0)
                y = F.conv2d(y, qw(sd[p + ".1.conv.weight"], wm), sd[p + ".1.conv.bias"])
                outputs[d.split("#")[0] + "-" + clf] = y
    return outputs


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    args = synth.model_args()
    sd = synth.make_state_dict(seed=0)
    tiles = synth.synthetic_tiles(n, 256, 256, seed=42)
    x = torch.from_numpy(tiles).float().permute(0, 3, 1, 2).contiguous()
    ref = net_oracle.forward(sd, x, args["decoder_kwargs"], args["considered_tasks"])
    modes = [("f16ed", "f32", None, None), ("f16ed", "f16", None, None), ("f32", "f32", None, None), ("f16", "f32", None, None), ("f32", "f16", None, None),
             ("f16", "f16", None, None), ("f16x2", "f16", None, None), ("f16x2", "f16", "f16x2", None),
             ("f16x2", "f16", "f16x2", "f16x2"), ("f16", "f16x2", None, None),
             ("f16x2", "f16x2", None, None), ("bf16", "bf16", None, None)]
    for wm, am, rm, sm in modes:
        out = forward_emul(sd, x, args["decoder_kwargs"], args["considered_tasks"], wm, am, rm, sm)
        worst = max(float((out[k] - ref[k]).abs().max()) for k in out)
        rms = float(np.sqrt(np.mean([float(((out[k] - ref[k]) ** 2).mean()) for k in out])))
        print("w=%-6s a=%-6s res=%-6s skip=%-6s  max-abs %.3e  rms %.3e" % (wm, am, rm, sm, worst, rms))


if __name__ == "__main__":
    main()
