"""Synthetic model directories and inputs (there are no shipped weights or datasets).

`make_state_dict` emits a checkpoint with exactly the keys/shapes of the reference
`NetDesc(resnet34)` state_dict (models/net_desc.py:23-103; 558 entries for the six-head
model incl. `num_batches_tracked` and the unused `backbone.fc`), so that the reference's
`load_state_dict(strict=True)` (infer/base.py:45) accepts it. Weights are seeded
kaiming-normal (fan_out) like models/utils/__init__.py:10-20, BatchNorm statistics are
calibrated on seeded images so activations stay O(1) through the 50-layer stack
(landmine 1) and the last 1x1 of every head is rescaled so logits are
O(1). `write_model_dir` writes the plugin surface of infer/base.py:28 / run_infer_tile.py:46-49:
`<dir>/weights.tar` = torch.save({"desc": state_dict}) and `<dir>/settings.yml`.
"""
import math
import os
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# models/paramset.yml:45-59 (the only in-repo copy of model_kwargs) + considered_tasks
DEFAULT_DECODER_KWARGS = OrderedDict([
    ("Lumen", OrderedDict(INST=3)), ("Gland", OrderedDict(INST=3)), ("Nuclei", OrderedDict(INST=3)),
    ("Nuclei#TYPE", OrderedDict(TYPE=7)), ("Gland#TYPE", OrderedDict(TYPE=3)),
    ("Patch-Class", OrderedDict(OUT=9)),
])
# order required by the WSI path (infer/wsi.py:610,626-633)
DEFAULT_CONSIDERED_TASKS = ["Nuclei", "Nuclei#TYPE", "Gland", "Gland#TYPE", "Lumen", "Patch-Class"]
# models/paramset.yml:37-43
DEFAULT_REQ_TARGET_CODE = OrderedDict([
    ("Lumen-INST", "IP-ERODED-CONTOUR-3"), ("Gland-INST", "IP-ERODED-CONTOUR-11"),
    ("Nuclei-INST", "IP-ERODED-CONTOUR-3"), ("Nuclei-TYPE", "TP"), ("Gland-TYPE", "TP"),
    ("Patch-Class", "PC"),
])
BLOCKS = [3, 4, 6, 3]
BACKBONE_BLOCKS = {"resnet34": [3, 4, 6, 3], "resnet18": [2, 2, 2, 2]}
FILTERS = [64, 64, 128, 256, 512]
# share of pixels whose (inner + contour) probability exceeds 0.5 on the calibration tiles
FOREGROUND_FRACTION = {"Nuclei": 0.35, "Gland": 0.35, "Lumen": 0.08}


def model_args(considered_tasks=None, decoder_kwargs=None, backbone="resnet34"):
    return {
        "encoder_backbone_name": backbone,
        "decoder_kwargs": decoder_kwargs if decoder_kwargs is not None else DEFAULT_DECODER_KWARGS,
        "considered_tasks": list(considered_tasks) if considered_tasks is not None
        else list(DEFAULT_CONSIDERED_TASKS),
    }


def synthetic_tiles(n, h, w, seed=0):
    """Seeded uint8 RGB tiles [n,h,w,3] with H&E-like structure: a smooth pink stroma field,
    dark purple elliptical "nuclei" (radius 3-8 px, soft edges), a few gland-like rings with a
    pale lumen, and mild pixel noise. A random-weight network does not segment them, but its
    maps inherit their blob structure instead of salt-and-pepper noise, which is what gives
    the post-processing a realistic amount of work."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.empty((n, h, w, 3), dtype=np.uint8)
    scale = (h * w) / (256.0 * 256.0)
    for i in range(n):
        img = np.zeros((h, w, 3), dtype=np.float32)
        for c, base in enumerate((225.0, 170.0, 205.0)):
            fx, fy = rng.uniform(0.01, 0.05, size=2)
            ph = rng.uniform(0, 6.28, size=2)
            img[..., c] = base + 18.0 * np.sin(xx * fx + ph[0]) * np.cos(yy * fy + ph[1])

        def blob(cy, cx, ry, rx, ang, colour, soft, ring=0.0):
            r = int(max(ry, rx) * 1.5) + 2
            y0, y1 = max(0, int(cy) - r), min(h, int(cy) + r + 1)
            x0, x1 = max(0, int(cx) - r), min(w, int(cx) + r + 1)
            if y0 >= y1 or x0 >= x1:
                return
            dy, dx = yy[y0:y1, x0:x1] - cy, xx[y0:y1, x0:x1] - cx
            ca, sa = np.cos(ang), np.sin(ang)
            d = np.sqrt(((dx * ca + dy * sa) / rx) ** 2 + ((-dx * sa + dy * ca) / ry) ** 2)
            a = np.clip((1.0 - d) / soft, 0.0, 1.0)
            if ring > 0.0:
                a = a * np.clip((d - ring) / soft, 0.0, 1.0)
            a = a[..., None]
            img[y0:y1, x0:x1] = img[y0:y1, x0:x1] * (1 - a) + np.asarray(colour, np.float32) * a

        for _ in range(int(rng.randint(2, 5) * scale + 0.5)):  # glands: ring of epithelium + lumen
            cy, cx = rng.uniform(0, h), rng.uniform(0, w)
            ry, rx = rng.uniform(22, 45, size=2)
            ang = rng.uniform(0, 3.14)
            blob(cy, cx, ry, rx, ang, (175.0, 105.0, 170.0), 0.15)
            blob(cy, cx, ry * 0.55, rx * 0.55, ang, (240.0, 232.0, 238.0), 0.2)
        for _ in range(int(rng.randint(70, 110) * scale + 0.5)):  # nuclei
            cy, cx = rng.uniform(0, h), rng.uniform(0, w)
            ry, rx = rng.uniform(3.0, 8.0, size=2)
            col = (rng.uniform(70, 110), rng.uniform(40, 75), rng.uniform(120, 160))
            blob(cy, cx, ry, rx, rng.uniform(0, 3.14), col, 0.35)
        img += rng.normal(0.0, 5.0, size=img.shape).astype(np.float32)
        out[i] = np.clip(img, 0, 255).astype(np.uint8)
    return out


class _Gen:
    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)

    def kaiming(self, o, i, k):
        std = math.sqrt(2.0 / (o * k * k))  # fan_out, relu gain
        return torch.randn(o, i, k, k, generator=self.g) * std

    def uniform(self, n, lo, hi):
        return torch.rand(n, generator=self.g) * (hi - lo) + lo


def make_state_dict(considered_tasks=None, decoder_kwargs=None, seed=0, calib_tiles=None,
                    logit_std=1.5, backbone="resnet34"):
    """Seeded, BN-calibrated checkpoint (CPU, fp32). ~2 s for the six-head model."""
    args = model_args(considered_tasks, decoder_kwargs)
    dk, tasks = args["decoder_kwargs"], args["considered_tasks"]
    gen = _Gen(seed)
    sd = OrderedDict()
    if calib_tiles is None:
        calib_tiles = synthetic_tiles(2, 256, 256, seed=seed + 1000)
    x = torch.from_numpy(calib_tiles).float().permute(0, 3, 1, 2).contiguous() / 255.0

    def bn_calibrated(prefix, pre, c):
        """Sets BN params from the batch statistics of `pre` and returns BN(pre)."""
        gamma = gen.uniform(c, 0.6, 1.4)
        beta = gen.uniform(c, -0.3, 0.3)
        mean = pre.mean(dim=(0, 2, 3))
        var = pre.var(dim=(0, 2, 3), unbiased=False) + 1e-6
        sd[prefix + ".weight"] = gamma
        sd[prefix + ".bias"] = beta
        sd[prefix + ".running_mean"] = mean
        sd[prefix + ".running_var"] = var
        sd[prefix + ".num_batches_tracked"] = torch.tensor(1, dtype=torch.int64)
        return F.batch_norm(pre, mean, var, gamma, beta, False, 0.0, 1e-5)

    def conv(prefix_w, inp, o, k, stride=1, bias_key=None):
        w = gen.kaiming(o, inp.shape[1], k)
        sd[prefix_w] = w
        b = None
        if bias_key is not None:
            b = gen.uniform(o, -0.1, 0.1)
            sd[bias_key] = b
        return F.conv2d(inp, w, b, stride=stride, padding=k // 2)

    with torch.no_grad():
        # ---- encoder (models/backbone/resnet.py:195-211)
        t = conv("backbone.conv1.weight", x, 64, 7)
        x0 = F.relu(bn_calibrated("backbone.bn1", t, 64))
        cur = F.max_pool2d(x0, 3, 2, 1)
        feats = [x0]
        for li, nb in enumerate(BACKBONE_BLOCKS[backbone], start=1):
            c = FILTERS[li]
            for bi in range(nb):
                p = "backbone.layer%d.%d" % (li, bi)
                stride = 2 if (li > 1 and bi == 0) else 1
                y = conv(p + ".conv1.weight", cur, c, 3, stride)
                y = F.relu(bn_calibrated(p + ".bn1", y, c))
                y = conv(p + ".conv2.weight", y, c, 3)
                y = bn_calibrated(p + ".bn2", y, c)
                if stride == 2:
                    idn = conv(p + ".downsample.0.weight", cur, c, 1, 2)
                    idn = bn_calibrated(p + ".downsample.1", idn, c)
                else:
                    idn = cur
                cur = F.relu(y + idn)
            feats.append(cur)
        sd["backbone.fc.weight"] = torch.randn(1000, 512, generator=gen.g) * 0.01
        sd["backbone.fc.bias"] = torch.zeros(1000)
        x4 = feats[4]
        f4 = conv("conv_map.weight", x4, 256, 1)

        # ---- decoders in decoder_kwargs order (nn.ModuleDict insertion order, net_desc.py:58-86)
        for d, heads in dk.items():
            if d not in tasks:
                continue
            if d == "Patch-Class":
                (_, ncls), = heads.items()
                feat = x4
                if feat.shape[2] != 9 and feat.shape[3] != 9:
                    h0 = int((feat.shape[2] - 9) * 0.5)
                    w0 = int((feat.shape[3] - 9) * 0.5)
                    feat = feat[:, :, h0:h0 + 9, w0:w0 + 9]
                pooled = feat.mean(dim=(2, 3), keepdim=True)
                p = "decoder_head.Patch-Class"
                # statistics over a 2-sample batch are degenerate: use spatial statistics of x4
                bn_calibrated(p + ".bn1", x4, 512)
                m, v = sd[p + ".bn1.running_mean"], sd[p + ".bn1.running_var"]
                y = F.relu(F.batch_norm(pooled, m, v, sd[p + ".bn1.weight"], sd[p + ".bn1.bias"],
                                        False, 0.0, 1e-5))
                y = conv(p + ".conv1.weight", y, 256, 1, bias_key=p + ".conv1.bias")
                sd[p + ".bn2.weight"] = gen.uniform(256, 0.6, 1.4)
                sd[p + ".bn2.bias"] = gen.uniform(256, -0.3, 0.3)
                sd[p + ".bn2.running_mean"] = gen.uniform(256, -0.2, 0.2)
                sd[p + ".bn2.running_var"] = gen.uniform(256, 0.5, 1.5)
                sd[p + ".bn2.num_batches_tracked"] = torch.tensor(1, dtype=torch.int64)
                conv(p + ".conv2.weight", y, ncls, 1, bias_key=p + ".conv2.bias")
                continue
            prev = f4
            chans = [(256, 128), (128, 64), (64, 64), (64, 64)]
            for blk in range(4):
                prev = F.interpolate(prev, scale_factor=2, mode="bilinear", align_corners=False)
                prev = feats[3 - blk] + prev
                for cv, oc in enumerate(chans[blk]):
                    p = "decoder_head.%s.%d.block.%d" % (d, blk, cv)
                    w = gen.kaiming(oc, prev.shape[1], 3)
                    b = gen.uniform(oc, -0.1, 0.1)
                    sd[p + ".conv.weight"] = w
                    sd[p + ".conv.bias"] = b
                    pre = F.conv2d(prev, w, b, padding=1)
                    prev = F.relu(bn_calibrated(p + ".bn", pre, oc))
            (clf, ncls), = heads.items()
            p = "output_head.%s.%s.x" % (d, clf)
            w = gen.kaiming(96, 64, 1)
            b = gen.uniform(96, -0.1, 0.1)
            sd[p + ".0.block.0.conv.weight"] = w
            sd[p + ".0.block.0.conv.bias"] = b
            hid = F.relu(bn_calibrated(p + ".0.block.0.bn", F.conv2d(prev, w, b), 96))
            w2 = gen.kaiming(ncls, 96, 1)
            b2 = gen.uniform(ncls, -0.5, 0.5)
            logits = F.conv2d(hid, w2, None)
            scale = logit_std / float(logits.std().clamp_min(1e-6))
            sd[p + ".1.conv.weight"] = w2 * scale
            if clf == "INST" and d in FOREGROUND_FRACTION:
                # bias the background class so that the thresholded foreground covers a realistic
                # share of the tile (a random head would call ~75 % of the pixels "nucleus")
                lg = logits * scale + b2.view(1, -1, 1, 1)
                lo_t, hi_t = -10.0, 20.0
                for _ in range(40):
                    t = 0.5 * (lo_t + hi_t)
                    shifted = lg.clone()
                    shifted[:, 0] += t
                    frac = float((1.0 - torch.softmax(shifted, 1)[:, 0] > 0.5).float().mean())
                    if frac > FOREGROUND_FRACTION[d]:
                        lo_t = t
                    else:
                        hi_t = t
                b2 = b2.clone()
                b2[0] += 0.5 * (lo_t + hi_t)
            sd[p + ".1.conv.bias"] = b2
    return sd


def write_model_dir(path, considered_tasks=None, decoder_kwargs=None, seed=0, state_dict=None,
                    backbone="resnet34"):
    """Writes <path>/weights.tar and <path>/settings.yml (the plugin surface)."""
    import yaml
    os.makedirs(path, exist_ok=True)
    args = model_args(considered_tasks, decoder_kwargs)
    if state_dict is None:
        state_dict = make_state_dict(args["considered_tasks"], args["decoder_kwargs"], seed, backbone=backbone)
    torch.save({"desc": state_dict}, os.path.join(path, "weights.tar"))
    settings = {
        "dataset_kwargs": {"input_shape": 448, "output_shape": 448, "class_input_shape": 144,
                           "req_target_code": dict(DEFAULT_REQ_TARGET_CODE)},
        "model_kwargs": {
            "encoder_backbone_name": backbone,
            "decoder_kwargs": {k: dict(v) for k, v in args["decoder_kwargs"].items()},
            "considered_tasks": list(args["considered_tasks"]),
        },
    }
    with open(os.path.join(path, "settings.yml"), "w") as f:
        yaml.safe_dump(settings, f, sort_keys=False)
    return state_dict


def postproc_field(h, w, tissue, seed=0):
    """Synthetic (inner, contour) probability maps for post-processing:
    f = normalised gaussian_filter(randn, sigma); inner = clip((f-t)*8, 0, 1);
    contour = 0.9 * clip(clip((f-(t-0.06))*8, 0, 1) - inner, 0, 1).
    Returns float32 [h,w,2]."""
    sigma, t = {"Nuclei": (4.0, 0.55), "Gland": (20.0, 0.5), "Lumen": (10.0, 0.6)}[tissue]
    rng = np.random.RandomState(seed)
    f = rng.standard_normal((h, w))
    f = _gaussian_blur(f, sigma)
    f = (f - f.min()) / (f.max() - f.min() + 1e-12)
    inner = np.clip((f - t) * 8.0, 0.0, 1.0)
    ring = np.clip((f - (t - 0.06)) * 8.0, 0.0, 1.0)
    contour = 0.9 * np.clip(ring - inner, 0.0, 1.0)
    return np.stack([inner, contour], axis=-1).astype(np.float32)


def _gaussian_blur(f, sigma):
    """Separable gaussian with reflect borders (numpy only; truncation at 4 sigma)."""
    r = int(4.0 * sigma + 0.5)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    for axis in (0, 1):
        pad = [(0, 0), (0, 0)]
        pad[axis] = (r, r)
        g = np.pad(f, pad, mode="symmetric")
        f = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), axis, g)
    return f


def adversarial_field(h, w, seed=0, big=False):
    """Piecewise-constant (inner, contour) maps drawn from a small value set: plateaus and
    exact threshold hits (0.5, 0.55) everywhere, blobs touching the borders, holes, specks of
    a few pixels. Stresses tie-breaking in the watershed heap, the strict `>` thresholds, the
    size filters and the crop rules. Values are multiples of 1/64 (exact in fp32).
    Returns float32 [h,w,2]."""
    rng = np.random.RandomState(seed)
    inner = np.zeros((h, w), dtype=np.float32)
    cnt = np.zeros((h, w), dtype=np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    vals = np.array([0.25, 0.5, 0.515625, 0.546875, 0.5625, 0.625, 0.75, 1.0], dtype=np.float32)
    n_blobs = rng.randint(6, 14)
    for _ in range(n_blobs):
        cy, cx = rng.randint(-4, h + 4), rng.randint(-4, w + 4)
        if big:
            ry, rx = rng.randint(8, max(9, h // 3)), rng.randint(8, max(9, w // 3))
        else:
            ry, rx = rng.randint(1, max(2, h // 6)), rng.randint(1, max(2, w // 6))
        d = ((yy - cy) / float(ry)) ** 2 + ((xx - cx) / float(rx)) ** 2
        if rng.rand() < 0.3:  # rectangle instead of ellipse
            d = np.maximum(np.abs(yy - cy) / float(ry), np.abs(xx - cx) / float(rx))
        ring = (d <= 1.0) & (d > 0.6)
        core = d <= 0.6
        cnt[ring] = rng.choice([0.25, 0.5, 0.53125, 0.75, 0.90625])
        inner[ring] = rng.choice([0.0, 0.0, 0.25, 0.5])
        inner[core] = rng.choice(vals)
        cnt[core] = rng.choice([0.0, 0.0, 0.0, 0.25])
        if rng.rand() < 0.4:  # a hole inside the core
            hole = d <= rng.choice([0.05, 0.15, 0.3])
            inner[hole] = rng.choice([0.0, 0.25, 0.5])
    for _ in range(rng.randint(5, 20)):  # specks around the size-filter limits
        y0, x0 = rng.randint(0, h), rng.randint(0, w)
        sh, sw = rng.randint(1, 4), rng.randint(1, 5)
        inner[y0:y0 + sh, x0:x0 + sw] = rng.choice([0.5625, 1.0])
    return np.stack([inner, cnt], axis=-1)
