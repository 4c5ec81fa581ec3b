"""Weight folding and packing: reference state_dict -> one contiguous blob for the C ABI.

Follows SURVEY.md Appendix D. Reference modules whose parameters are consumed here:
  models/backbone/resnet.py:32-48,81-97,195-200  (conv / bn / downsample, no conv bias)
  models/utils/conv_layers.py:24-60               (_ConvLayer: conv(bias) -> BN(eps 1e-5) -> ReLU)
  models/net_desc.py:52,64-76                     (conv_map, Patch-Class head)
  models/utils/net_layers.py:31-38                (classification head)
BatchNorm (eval mode) is folded into the preceding convolution in float64, then the result
is split into fp16 hi + fp16 lo planes (lo is only read in CERB_PREC_F16X2 mode).
"""
import numpy as np

BN_EPS = 1e-5
ALIGN = 256


class BlobBuilder:
    def __init__(self):
        self._chunks = []
        self._size = 0

    def add(self, arr):
        arr = np.ascontiguousarray(arr)
        pad = (-self._size) % ALIGN
        if pad:
            self._chunks.append(b"\0" * pad)
            self._size += pad
        off = self._size
        raw = arr.tobytes()
        self._chunks.append(raw)
        self._size += len(raw)
        return off

    def finish(self):
        return np.frombuffer(b"".join(self._chunks), dtype=np.uint8).copy()


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


def fold_bn(weight, bias, bn, eps=BN_EPS):
    """conv -> BN(eval) as a single conv. weight [O,I,kh,kw]; bias [O] or None; bn = dict of
    weight/bias/running_mean/running_var arrays (float64)."""
    s = bn["weight"] / np.sqrt(bn["running_var"] + eps)
    w = weight * s[:, None, None, None]
    b0 = bias if bias is not None else np.zeros(weight.shape[0])
    b = (b0 - bn["running_mean"]) * s + bn["bias"]
    return w, b


def bn_of(sd, prefix):
    return {k: _np(sd[prefix + "." + k]) for k in ("weight", "bias", "running_mean", "running_var")}


def split_f16(x):
    """fp64/fp32 array -> (hi, lo) fp16 planes with hi + lo ~= x to ~22 bits."""
    x = np.asarray(x, dtype=np.float64)
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float64)).astype(np.float16)
    return hi, lo


def round_f16_diffused(w):
    """fp16 rounding of a weight tensor [O, I, kh, kw] with error diffusion along K.

    Round-to-nearest leaves every weight with an independent error of up to half an ulp; a
    post-ReLU input channel has a large positive mean (and is smooth over the 3x3 window), so
    those errors add up COHERENTLY: sum_k dw[o,k] * mean_k is a constant offset per output
    channel that the next 50 layers amplify. Measured on the six-head model
    (tools/precision_study.py): weight rounding alone costs 0.23 max-abs on the logits, the
    fp16 activations only 0.04. Here the rounding error of each weight is carried into the
    next one of the same output channel, walking K as (input channel, tap), so the sum of the
    errors over the nine taps of one input channel - and over the whole row - stays below one
    ulp: 0.23 -> 0.04 at zero run-time cost. The residual w - hi still fits the lo plane."""
    w = np.asarray(w, dtype=np.float64)
    o = w.shape[0]
    flat = w.reshape(o, -1)
    hi = np.empty(flat.shape, dtype=np.float16)
    carry = np.zeros(o, dtype=np.float64)
    for k in range(flat.shape[1]):
        want = flat[:, k] + carry
        r = want.astype(np.float16)
        carry = want - r.astype(np.float64)
        hi[:, k] = r
    return hi.reshape(w.shape)


def split_f16_diffused(w):
    """[O, I, kh, kw] -> (hi, lo): hi by error diffusion along K, lo = fp16(w - hi)."""
    w = np.asarray(w, dtype=np.float64)
    hi = round_f16_diffused(w)
    lo = (w - hi.astype(np.float64)).astype(np.float16)
    return hi, lo


def weight_shift(w):
    """Power-of-two pre-scale: max|w * 2^shift| lands in [256, 512], far from fp16 subnormals
    (so the lo plane keeps its precision) and far from overflow."""
    m = float(np.max(np.abs(w)))
    if m == 0.0 or not np.isfinite(m):
        return 0
    return int(np.clip(np.floor(np.log2(512.0 / m)), -60, 60))


def pack_conv(blob, w, b):
    """w [O,I,kh,kw] float64 (BN folded) -> K-major [O][kh*kw][I] fp16 hi/lo; b -> fp32.
    Returns dict(w_off, w_lo_off, b_off, cout, cin, kh, kw)."""
    o, i, kh, kw = w.shape
    sh = weight_shift(w)
    hi, lo = split_f16_diffused(np.ldexp(w, sh))
    km = lambda a: np.ascontiguousarray(np.transpose(a, (0, 2, 3, 1)).reshape(o, kh * kw * i))  # noqa: E731
    hi, lo = km(hi), km(lo)
    out = {"w_off": blob.add(hi), "w_lo_off": blob.add(lo), "cout": o, "cin": i, "kh": kh, "kw": kw,
           "w_shift": sh}
    out["b_off"] = blob.add(b.astype(np.float32)) if b is not None else -1
    return out


def pack_stem(blob, w, b):
    """Stem 7x7, 3->64. The device reads the image through an overlapping-window view of
    the [N,H,W+8,8] PREP tensor: per filter row one K chunk of 64 = 8 pixels x 8 channels
    (pixel 7 and channels 3..7 carry zero weights). `imgs / 255` (models/net_desc.py:147)
    is folded into the weights."""
    o, i, kh, kw = w.shape
    assert (i, kh, kw) == (3, 7, 7)
    w = w / 255.0
    sh = weight_shift(w)
    whi, wlo = split_f16_diffused(np.ldexp(w, sh))

    def km(a):
        out = np.zeros((o, 7, 8, 8), dtype=np.float16)
        out[:, :, :7, :3] = np.transpose(a, (0, 2, 3, 1))
        return out.reshape(o, 7 * 64)

    hi, lo = km(whi), km(wlo)
    return {"w_off": blob.add(hi), "w_lo_off": blob.add(lo), "b_off": blob.add(b.astype(np.float32)),
            "cout": o, "cin": 3, "kh": 7, "kw": 7, "w_shift": sh}


def strip_module_prefix(sd):
    """infer/base.py:31-45: checkpoints saved from nn.DataParallel carry a 'module.' prefix."""
    if all(k.startswith("module.") for k in sd.keys()):
        return {k[len("module."):]: v for k, v in sd.items()}
    return dict(sd)
