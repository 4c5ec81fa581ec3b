"""ORACLE — test infrastructure only. Never imported by the product path (cerberus_b200/).

CPU fp32 restatement (torch.nn.functional, no nn.Module, NCHW like the reference) of the
reference forward + step function, straight from a reference-format state_dict:

  forward()     <- models/net_desc.py:144-200 (NetDesc.forward)
                   models/backbone/resnet.py:81-97 (BasicBlock), :273-286 (_forward_impl)
                   models/utils/conv_layers.py:53-58 (conv -> BN -> ReLU)
                   models/utils/net_layers.py:31-38,45-46 (head, bilinear x2)
  infer_step()  <- models/run_desc.py:439-502

Parity pin: tests/golden/forward_*.npz were produced by oracle/gen_golden.py from the
UNMODIFIED reference NetDesc / infer_step imported from /root/reference; tests/test_oracle_net.py
checks this restatement against them (CPU, `-m "not gpu"`).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BLOCKS = [3, 4, 6, 3]

# models/run_desc.py:472-479
HEAD_NAME_MAP = {
    "Gland": "Gland-INST", "Gland#TYPE": "Gland-TYPE", "Lumen": "Lumen-INST",
    "Nuclei": "Nuclei-INST", "Nuclei#TYPE": "Nuclei-TYPE", "Patch-Class": "Patch-Class",
}


def _bn(x, sd, p, eps=1e-5):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], False, 0.0, eps)


def cropping_center_nchw(x, crop_shape):
    """models/utils/misc_utils.py:6-25 with batch=True."""
    h0 = int((x.shape[2] - crop_shape[0]) * 0.5)
    w0 = int((x.shape[3] - crop_shape[1]) * 0.5)
    return x[:, :, h0:h0 + crop_shape[0], w0:w0 + crop_shape[1]]


def cropping_center_nhwc(x, crop_shape):
    """misc/utils.py:94-104 with batch=True."""
    h0 = int((x.shape[1] - crop_shape[0]) * 0.5)
    w0 = int((x.shape[2] - crop_shape[1]) * 0.5)
    return x[:, h0:h0 + crop_shape[0], w0:w0 + crop_shape[1]]


def forward_half(sd, imgs_nchw_f32, decoder_kwargs, considered_tasks):
    """The same forward as `net.half()` would compute it ("vs net.half()
    reference"): parameters and the input rounded to fp16, every module output (conv, BN, ReLU,
    add, interpolate, pool) rounded to fp16, arithmetic inside a module in fp32 - which is
    what cuDNN / ATen half kernels do (fp32 accumulation, fp16 storage). Returns fp32 logits."""
    return forward(sd, imgs_nchw_f32, decoder_kwargs, considered_tasks, _half=True)


def _half_ops(enabled):
    """(F-like namespace, rounding function): every call's result goes through fp16."""
    if not enabled:
        return F, (lambda t: t), _bn

    def r(t):
        return t.half().float()

    class H:
        conv2d = staticmethod(lambda *a, **k: r(F.conv2d(*a, **k)))
        relu = staticmethod(lambda x: r(F.relu(x)))
        max_pool2d = staticmethod(lambda *a, **k: F.max_pool2d(*a, **k))
        interpolate = staticmethod(lambda *a, **k: r(F.interpolate(*a, **k)))
        adaptive_avg_pool2d = staticmethod(lambda *a, **k: r(F.adaptive_avg_pool2d(*a, **k)))

    return H, r, (lambda x, sd, p, eps=1e-5: r(_bn(x, sd, p, eps)))


def forward(sd, imgs_nchw_f32, decoder_kwargs, considered_tasks, return_feats=False, _half=False):
    """imgs: float32 NCHW in 0..255. Returns OrderedDict head -> NCHW logits (net_desc.py:198)."""
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    if _half:
        sd = {k: (v.half().float() if v.is_floating_point() else v) for k, v in sd.items()}
    return _forward(sd, imgs_nchw_f32, decoder_kwargs, considered_tasks, return_feats, *_half_ops(_half))


def _forward(sd, imgs_nchw_f32, decoder_kwargs, considered_tasks, return_feats, F, r, _bn):  # noqa: N803
    with torch.no_grad():
        x = r(r(imgs_nchw_f32) / 255.0)  # net_desc.py:147
        x = F.conv2d(x, sd["backbone.conv1.weight"], None, 1, 3)  # resnet.py:195-197 (stride 1!)
        x0 = x = F.relu(_bn(x, sd, "backbone.bn1"))
        x = F.max_pool2d(x, 3, 2, 1)  # resnet.py:201
        feats = [x0]
        for li in range(1, 5):  # blocks per stage from the checkpoint (resnet34: 3,4,6,3; resnet18: 2,2,2,2)
            nb = 0
            while ("backbone.layer%d.%d.conv1.weight" % (li, nb)) in sd:
                nb += 1
            for bi in range(nb):
                p = "backbone.layer%d.%d" % (li, bi)
                stride = 2 if (li > 1 and bi == 0) else 1
                out = F.conv2d(x, sd[p + ".conv1.weight"], None, stride, 1)
                out = F.relu(_bn(out, sd, p + ".bn1"))
                out = F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1)
                out = _bn(out, sd, p + ".bn2")
                if (p + ".downsample.0.weight") in sd:
                    idn = F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride, 0)
                    idn = _bn(idn, sd, p + ".downsample.1")
                else:
                    idn = x
                x = F.relu(r(out + idn))
            feats.append(x)
        bottom = feats[-1]
        feat_list = list(feats)
        feat_list[-1] = F.conv2d(bottom, sd["conv_map.weight"])  # net_desc.py:152-153
        outputs = OrderedDict()
        inter = OrderedDict()
        for d, heads in decoder_kwargs.items():
            if d not in considered_tasks:
                continue
            if d == "Patch-Class":
                # net_desc.py:169-180 — note: `bottom_feats` is re-bound by the crop
                if bottom.shape[2] != 9 and bottom.shape[3] != 9:
                    bottom = cropping_center_nchw(bottom, [9, 9])
                pooled = F.adaptive_avg_pool2d(bottom, (1, 1))
                p = "decoder_head.Patch-Class"
                y = F.relu(_bn(pooled, sd, p + ".bn1"))
                y = F.conv2d(y, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"])
                y = F.relu(_bn(y, sd, p + ".bn2"))
                y = F.conv2d(y, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"])
                outputs[d] = y
                continue
            prev = feat_list[-1]
            for idx in range(1, len(feat_list)):  # net_desc.py:184-189
                prev = F.interpolate(prev, scale_factor=2, mode="bilinear", align_corners=False)
                prev = r(feat_list[-(idx + 1)] + prev)
                for cv in range(2):
                    p = "decoder_head.%s.%d.block.%d" % (d, idx - 1, cv)
                    prev = F.conv2d(prev, sd[p + ".conv.weight"], sd[p + ".conv.bias"], 1, 1)
                    prev = F.relu(_bn(prev, sd, p + ".bn"))
                inter["%s.u%d" % (d, 5 - idx)] = prev
            for clf in heads:
                p = "output_head.%s.%s.x" % (d, clf)
                y = F.conv2d(prev, sd[p + ".0.block.0.conv.weight"], sd[p + ".0.block.0.conv.bias"])
                y = F.relu(_bn(y, sd, p + ".0.block.0.bn"))
                y = F.conv2d(y, sd[p + ".1.conv.weight"], sd[p + ".1.conv.bias"])
                outputs[d.split("#")[0] + "-" + clf] = y
    if return_feats:
        return outputs, feats, inter
    return outputs


def infer_step(sd, batch_u8_nhwc, output_shape, decoder_kwargs, considered_tasks):
    """models/run_desc.py:439-502 on CPU. batch: uint8 [N,h,w,3] (numpy or torch).
    Returns (list of per-sample dicts, dict of NCHW logits)."""
    imgs = torch.as_tensor(np.asarray(batch_u8_nhwc)).type(torch.float32)
    imgs = imgs.permute(0, 3, 1, 2).contiguous()
    if not isinstance(output_shape, (list, tuple)):
        output_shape = [output_shape, output_shape]
    logits = forward(sd, imgs, decoder_kwargs, considered_tasks)
    pred = OrderedDict((k, v.permute(0, 2, 3, 1).contiguous()) for k, v in logits.items())
    sub = OrderedDict()
    for task in considered_tasks:
        name = HEAD_NAME_MAP[task]
        out = pred[name]
        if name == "Patch-Class":
            out = torch.argmax(torch.softmax(out, -1), dim=-1, keepdim=True)
            out = F.interpolate(out.type(torch.float32), size=list(output_shape), mode="nearest")
            out = torch.squeeze(out)
            if out.dim() == 2:
                out = torch.unsqueeze(out, 0)
        else:
            out = torch.softmax(out, -1)
            if name.endswith("-INST"):
                out = out[..., 1:]
            out = cropping_center_nhwc(out, output_shape)
        if "TYPE" in name:
            out = torch.argmax(out, dim=-1, keepdim=False)
        sub[name] = out.numpy()
    n = imgs.shape[0]
    return [{k: v[i] for k, v in sub.items()} for i in range(n)], logits
