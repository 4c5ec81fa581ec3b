"""GPU parity of the tcgen05 implicit-GEMM convolution against torch fp32 conv2d of the
same fp16-rounded operands (the only floating-point kernel family: torch fp32 reference)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cerberus_b200 import _lib
from cerberus_b200.engine import Context, ForwardPlan
from cerberus_b200.pack import BlobBuilder, pack_conv, pack_stem
from tests.util import MiniModel, MiniSpec, f16, nchw_to_nhwc, nhwc_to_nchw

pytestmark = pytest.mark.gpu

CASES = [
    # n, h, w, cin, cout, k, stride, residual, relu
    (2, 32, 32, 64, 64, 3, 1, False, True),     # layer1-like
    (1, 128, 128, 64, 64, 3, 1, True, True),    # full-width rows, residual
    (2, 64, 64, 64, 128, 3, 2, False, True),    # stride-2 parity views
    (2, 64, 64, 64, 128, 1, 2, False, False),   # downsample 1x1 s2
    (1, 16, 16, 512, 256, 1, 1, False, False),  # conv_map
    (2, 28, 28, 256, 256, 3, 1, True, True),    # non power-of-two map (448 input), partial tiles
    (1, 32, 32, 256, 1280, 3, 1, False, True),  # fused first decoder stage (5 decoders)
    (1, 64, 64, 64, 96, 1, 1, False, True),     # head hidden layer, BN = 96
    (3, 16, 16, 512, 512, 3, 1, True, True),    # layer4: K = 4608
    (1, 48, 80, 128, 64, 3, 1, False, True),    # non-square
    (2, 64, 64, 128, 128, 3, 1, False, True),   # layer2 / decoder u3: halo kernel, several regions per CTA
    (5, 40, 24, 128, 128, 3, 1, True, False),   # partial regions in x and y, residual, no ReLU
    (40, 16, 16, 256, 128, 3, 1, False, True),  # more work items than SMs, 4 chunks
]


def _run_case(precision, n, h, w, cin, cout, k, stride, residual, relu, seed=0):
    rng = np.random.RandomState(seed)
    x = rng.standard_normal((n, h, w, cin)).astype(np.float32)
    wt = (rng.standard_normal((cout, cin, k, k)) * (1.0 / np.sqrt(cin * k * k))).astype(np.float32)
    b = rng.uniform(-0.5, 0.5, cout).astype(np.float32)
    oh = (h + 2 * (k // 2) - k) // stride + 1
    ow = (w + 2 * (k // 2) - k) // stride + 1
    res = rng.standard_normal((n, oh, ow, cout)).astype(np.float32) if residual else None
    if precision == "f16":
        x = f16(x).astype(np.float32)
        wt = f16(wt).astype(np.float32)
        if res is not None:
            res = f16(res).astype(np.float32)
    blob = BlobBuilder()
    layer = pack_conv(blob, wt.astype(np.float64), b.astype(np.float64))
    spec = MiniSpec()
    t_in = spec._tensor("in", n, h, w, cin)
    t_out = spec._tensor("out", n, oh, ow, cout)
    t_res = spec._tensor("res", n, oh, ow, cout) if residual else -1
    spec._conv(layer, t_in, t_out, relu=int(relu), stride=stride, residual=t_res)
    ctx = Context(0, precision)
    plan = ForwardPlan(ctx, MiniModel(blob), 0, 0, 0, 0, 0, spec=spec)

    def put(tid, a):
        hi = a.astype(np.float16)
        plan.write(tid, hi, 0)
        if precision == "f16x2":
            plan.write(tid, (a - hi.astype(np.float32)).astype(np.float16), 1)

    put(t_in, x)
    if residual:
        put(t_res, res)
    plan.run()
    got = plan.read(t_out).astype(np.float32)
    ref = F.conv2d(torch.from_numpy(nhwc_to_nchw(x)).double(), torch.from_numpy(wt).double(),
                   torch.from_numpy(b).double(), stride=stride, padding=k // 2)
    if residual:
        ref = ref + torch.from_numpy(nhwc_to_nchw(res)).double()
    if relu:
        ref = F.relu(ref)
    ref = nchw_to_nhwc(ref.numpy())
    plan.close()
    ctx.close()
    return got, ref


@pytest.mark.parametrize("mode", [-1, 0, 1, 2, 3])
@pytest.mark.parametrize("case", [c for c in CASES if c[3] == 64 and c[4] == 64 and c[5] == 3 and c[6] == 1] +
                         [(2, 50, 36, 64, 64, 3, 1, True, True), (1, 16, 8, 64, 64, 3, 1, False, False)],
                         ids=lambda c: "x".join(map(str, c)))
def test_conv64_every_kernel_variant(case, mode, built_lib, monkeypatch):
    """64->64 3x3: generic kernel (-1), the three halo layouts of conv64.cu and the even/odd
    N = 128 formulation of conv64x.cu (3, the default), incl. partial regions and a residual."""
    monkeypatch.setenv("CERB_CONV64_MODE", str(mode))
    got, ref = _run_case("f16", *case)
    err = np.abs(got - ref)
    tol = 2e-3 * np.abs(ref) + 2e-3
    assert np.all(err <= tol), "mode %d: max err %g" % (mode, err.max())


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_f16(case, built_lib):
    got, ref = _run_case("f16", *case)
    # fp32 accumulation of exact fp16 products; output rounded to fp16 (rel 2^-11)
    err = np.abs(got - ref)
    tol = 2e-3 * np.abs(ref) + 2e-3
    assert np.all(err <= tol), "max err %g at %r" % (err.max(), np.unravel_index(err.argmax(), err.shape))


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_f16x2(case, built_lib):
    got, ref = _run_case("f16x2", *case)
    err = np.abs(got - ref)
    # split operands: ~22-bit products; the tensor-core fp32 accumulator truncates per MMA step
    tol = 5e-5 * np.abs(ref) + 1e-4
    assert np.all(err <= tol), "max err %g at %r" % (err.max(), np.unravel_index(err.argmax(), err.shape))


def test_stem(built_lib):
    """7x7 stride-1 stem through the overlapping-window PREP view (resnet.py:195-200)."""
    rng = np.random.RandomState(3)
    n, h, w = 2, 64, 96
    img = rng.randint(0, 256, size=(n, h, w, 3)).astype(np.uint8)
    wt = (rng.standard_normal((64, 3, 7, 7)) * 0.08).astype(np.float64)
    b = rng.uniform(-0.5, 0.5, 64)
    for precision, tol in (("f16", 4e-3), ("f16x2", 2e-4)):
        blob = BlobBuilder()
        layer = pack_stem(blob, wt, b)
        spec = MiniSpec()
        t_in = spec._tensor("input", n, h, w, 3, _lib.CERB_U8)
        t_prep = spec._tensor("prep", n, h, w + 8, 8)
        t_out = spec._tensor("x0", n, h, w, 64)
        spec._op(_lib.OP_PREP, in0=t_in, out=t_prep)
        spec._conv(layer, t_prep, t_out, relu=1, stem=1)
        ctx = Context(0, precision)
        plan = ForwardPlan(ctx, MiniModel(blob), 0, 0, 0, 0, 0, spec=spec)
        plan.write(t_in, img)
        plan.run()
        got = plan.read(t_out).astype(np.float32)
        ref = F.relu(F.conv2d(torch.from_numpy(nhwc_to_nchw(img.astype(np.float64))) / 255.0,
                              torch.from_numpy(wt), torch.from_numpy(b), padding=3))
        ref = nchw_to_nhwc(ref.numpy())
        err = np.abs(got - ref)
        assert np.all(err <= tol * np.abs(ref) + tol), "%s: max err %g" % (precision, err.max())
        plan.close()
        ctx.close()


@pytest.mark.parametrize("mode", [1, 3])
@pytest.mark.parametrize("shape", [(1, 32, 16), (2, 48, 40), (1, 128, 128), (3, 16, 16), (40, 32, 32)])
def test_conv64_fused_upsample_add(shape, mode, built_lib):
    """conv64 with the UPADD op fused into its producer: conv(skip + bilinear_x2(prev)) with
    align_corners=False (models/net_desc.py:185-188, net_layers.py:45-46); also checks it against
    the unfused upadd kernel + conv (same fp16 rounding of the sum -> identical output)."""
    n, h, w = shape
    rng = np.random.RandomState(5)
    skip = f16(rng.standard_normal((n, h, w, 64))).astype(np.float32)
    prev = f16(rng.standard_normal((n, h // 2, w // 2, 64))).astype(np.float32)
    wt = f16(rng.standard_normal((64, 64, 3, 3)) / 24.0).astype(np.float32)
    b = rng.uniform(-0.5, 0.5, 64).astype(np.float32)
    blob = BlobBuilder()
    layer = pack_conv(blob, wt.astype(np.float64), b.astype(np.float64))
    outs = []
    for fused in (True, False):
        spec = MiniSpec()
        t_skip = spec._tensor("skip", n, h, w, 64)
        t_prev = spec._tensor("prev", n, h // 2, w // 2, 64)
        t_out = spec._tensor("out", n, h, w, 64)
        if fused:
            spec._conv(layer, t_skip, t_out, relu=1, up_prev1=t_prev + 1)
        else:
            t_sum = spec._tensor("sum", n, h, w, 64)
            spec._op(_lib.OP_UPADD, in0=t_skip, in1=t_prev, out=t_sum)
            spec._conv(layer, t_sum, t_out, relu=1)
        ctx = Context(0, "f16")
        # mode 1: producer warps of conv64.cu build the sum from global loads; mode 3 (default
        # kernel, conv64x.cu): fix-up warps add bilinear_x2(prev) to the TMA-loaded skip halo
        ctx.set_option("conv64_mode", mode)
        plan = ForwardPlan(ctx, MiniModel(blob), 0, 0, 0, 0, 0, spec=spec)
        plan.write(t_skip, skip.astype(np.float16))
        plan.write(t_prev, prev.astype(np.float16))
        plan.run()
        outs.append(plan.read(t_out).astype(np.float32))
        plan.close()
        ctx.close()
    assert np.array_equal(outs[0], outs[1]), "fused and unfused paths differ: %g" % np.abs(outs[0] - outs[1]).max()
    up = F.interpolate(torch.from_numpy(nhwc_to_nchw(prev)), scale_factor=2, mode="bilinear", align_corners=False)
    s = (torch.from_numpy(nhwc_to_nchw(skip)) + up).half().double()
    ref = F.relu(F.conv2d(s, torch.from_numpy(wt).double(), torch.from_numpy(b).double(), padding=1))
    ref = nchw_to_nhwc(ref.numpy())
    err = np.abs(outs[0] - ref)
    assert np.all(err <= 4e-3 * np.abs(ref) + 4e-3), err.max()


PAIR_CASES = [
    # n, h, w, cin, cout, k, stride, residual, relu  (3x3 stride 1, Cout % 256 == 0)
    (2, 32, 32, 256, 256, 3, 1, False, True),     # layer3 body
    (2, 28, 28, 256, 256, 3, 1, True, True),      # partial regions (448 input), residual
    (3, 16, 16, 512, 512, 3, 1, True, True),      # layer4: two 256-channel tiles, K = 4608
    (1, 16, 16, 128, 256, 3, 1, False, False),    # one region, 2 chunks, no ReLU
    (37, 16, 16, 256, 256, 3, 1, True, True),     # an odd number of items
    (5, 40, 24, 128, 512, 3, 1, True, False),     # partial regions in x and y
    (80, 16, 16, 256, 256, 3, 1, False, True),    # more items than CTA pairs (dynamic scheduling)
    (2, 64, 64, 128, 128, 3, 1, True, True),      # N = 128 per pair (64 weight rows per CTA): layer2 / u3
    (5, 40, 24, 128, 128, 3, 1, True, False),
    (40, 16, 16, 256, 128, 3, 1, False, True),    # decoder u4 second conv
    (1, 48, 80, 128, 64, 3, 1, False, True),      # N = 64 per pair (32 weight rows per CTA): u3 second conv
    (3, 64, 64, 128, 64, 3, 1, False, True),
]


@pytest.mark.parametrize("dyn", [1, 0])
@pytest.mark.parametrize("case", PAIR_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv3x3_cta_pair_kernel(case, dyn, built_lib, monkeypatch):
    """Wide 3x3 layers on CTA pairs (tcgen05.mma.cta_group::2, csrc/conv3x3c2.cu): M = 256 x N = 256 /
    128 / 64 MMAs whose operands and accumulator are split over the two CTAs of a cluster. Run with
    the pair kernel forced on and compared with the fp64 convolution; the default-mode tests above
    (test_conv_f16) cover whichever kernel the library picks."""
    monkeypatch.setenv("CERB_CONV3_PAIR", "2")  # also the N = 128 / 64 variants (off by default: slower)
    monkeypatch.setenv("CERB_DYN_SCHED", str(dyn))
    got, ref = _run_case("f16", *case)
    err = np.abs(got - ref)
    tol = 2e-3 * np.abs(ref) + 2e-3
    assert np.all(err <= tol), "max err %g at %r" % (err.max(), np.unravel_index(err.argmax(), err.shape))


CHAIN_CASES = [
    # n, h, w, channels, layers
    (3, 32, 32, 256, 5),     # layer3 body: 4 regions per image
    (5, 40, 24, 256, 4),     # partial regions in x and y
    (37, 16, 16, 512, 5),    # layer4 body: two 256-channel tiles per region, K = 4608
    (1, 16, 16, 256, 6),     # ONE item per layer: every item waits for the one before it
    (2, 16, 16, 256, 3),     # fewer items than layers x pairs
    (160, 16, 16, 256, 3),   # more items per layer than CTA pairs
    (2, 64, 64, 128, 4),     # layer2 body: the single-CTA kernel (conv3x3.cu) chains the same way
    (5, 40, 24, 128, 3),
    (1, 16, 16, 128, 5),     # one item per layer
    (200, 16, 16, 128, 3),   # more items per layer than SMs
]


@pytest.mark.parametrize("case", CHAIN_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv3x3_layer_chain_is_one_launch_with_identical_results(case, built_lib, monkeypatch):
    """Consecutive pair-kernel layers of one geometry run as ONE persistent launch whose items wait
    for the previous layer of their image (Conv3c2Params::layers; encoder layer3 / layer4 bodies,
    models/backbone/resnet.py:203-211 with BasicBlock residuals). Same bits as layer-by-layer
    launches, fewer launches, and the last layer within fp16 tolerance of the fp64 chain."""
    n, h, w, ch, n_layers = case
    rng = np.random.RandomState(11)
    x = f16(rng.standard_normal((n, h, w, ch)).astype(np.float32)).astype(np.float32)
    wts = [f16((rng.standard_normal((ch, ch, 3, 3)) * (1.0 / np.sqrt(ch * 9))).astype(np.float32)).astype(np.float32)
           for _ in range(n_layers)]
    bs = [rng.uniform(-0.5, 0.5, ch).astype(np.float32) for _ in range(n_layers)]

    def run(chain):
        monkeypatch.setenv("CERB_CONV3_CHAIN", str(chain))
        blob = BlobBuilder()
        layers = [pack_conv(blob, wt.astype(np.float64), b.astype(np.float64)) for wt, b in zip(wts, bs)]
        spec = MiniSpec()
        tids = [spec._tensor("t%d" % i, n, h, w, ch) for i in range(n_layers + 1)]
        for i, layer in enumerate(layers):
            # BasicBlock pattern: every second conv adds the tensor two steps back; the last has no ReLU
            res = tids[i - 1] if i % 2 == 1 else -1
            spec._conv(layer, tids[i], tids[i + 1], relu=int(i != n_layers - 1), stride=1, residual=res)
        ctx = Context(0, "f16")
        plan = ForwardPlan(ctx, MiniModel(blob), 0, 0, 0, 0, 0, spec=spec)
        plan.write(tids[0], x.astype(np.float16), 0)
        outs = []
        for rep in range(3):  # eager, graph capture, graph replay
            before = ctx.launch_count
            plan.run()
            ctx.sync()
            launches = ctx.launch_count - before
            outs.append([plan.read(t) for t in tids[1:]])
        plan.close()
        ctx.close()
        for o in outs[1:]:
            for a, b in zip(o, outs[0]):
                assert np.array_equal(a, b)
        return outs[0], launches

    chained, l1 = run(1)
    single, l0 = run(0)
    assert l0 == n_layers and l1 == 1, (l0, l1)
    for i, (a, b) in enumerate(zip(chained, single)):
        assert np.array_equal(a, b), "layer %d differs between the chained and the layer-by-layer launch" % i
    # fp64 chain on the device's own fp16 intermediates (each layer checked on its actual input)
    prev = [x] + [a.astype(np.float32) for a in chained]
    for i in range(n_layers):
        ref = F.conv2d(torch.from_numpy(nhwc_to_nchw(prev[i])).double(), torch.from_numpy(wts[i]).double(),
                       torch.from_numpy(bs[i]).double(), padding=1)
        if i % 2 == 1:
            ref = ref + torch.from_numpy(nhwc_to_nchw(prev[i - 1])).double()
        if i != n_layers - 1:
            ref = F.relu(ref)
        ref = nchw_to_nhwc(ref.numpy())
        err = np.abs(prev[i + 1] - ref)
        assert np.all(err <= 2e-3 * np.abs(ref) + 2e-3), (i, err.max())


SPLIT64_CASES = [
    # n, h, w, cin, cout, k, stride, residual, relu   (64 -> 64 3x3 stride 1, split-precision mode)
    (2, 32, 32, 64, 64, 3, 1, False, True),
    (1, 128, 128, 64, 64, 3, 1, True, True),
    (2, 50, 36, 64, 64, 3, 1, True, True),       # partial tiles in x and y, residual
    (1, 16, 8, 64, 64, 3, 1, False, False),      # a single tile
    (3, 40, 24, 64, 64, 3, 1, True, False),
    (40, 16, 16, 64, 64, 3, 1, False, True),     # more tiles than SMs
]


@pytest.mark.parametrize("kernel", [1, 0])
@pytest.mark.parametrize("case", SPLIT64_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv64_split_precision_kernel(case, kernel, built_lib, monkeypatch):
    """Split-precision 64->64 3x3: the halo-reuse kernel (csrc/conv64s.cu, default) and the generic
    kernel against the fp64 convolution at the parity-mode tolerance."""
    monkeypatch.setenv("CERB_CONV64S", str(kernel))
    got, ref = _run_case("f16x2", *case)
    err = np.abs(got - ref)
    tol = 5e-5 * np.abs(ref) + 1e-4
    assert np.all(err <= tol), "max err %g at %r" % (err.max(), np.unravel_index(err.argmax(), err.shape))


@pytest.mark.parametrize("precision", ["f16", "f16x2"])
@pytest.mark.parametrize("shape", [(2, 16, 24, 64), (1, 14, 14, 256), (3, 32, 32, 128), (2, 5, 7, 40)])
def test_upadd_matches_torch_bilinear(shape, precision, built_lib):
    """out = skip + bilinear_x2(prev), align_corners=False (models/utils/net_layers.py:45-46,
    models/net_desc.py:185-188), both precision modes (40 channels: the generic fallback kernel)."""
    n, ph, pw, c = shape
    rng = np.random.RandomState(5)
    prev = rng.standard_normal((n, ph, pw, c)).astype(np.float32)
    skip = rng.standard_normal((n, 2 * ph, 2 * pw, c)).astype(np.float32)
    if precision == "f16":
        prev, skip = f16(prev).astype(np.float32), f16(skip).astype(np.float32)
    spec = MiniSpec()
    t_skip = spec._tensor("skip", n, 2 * ph, 2 * pw, c)
    t_prev = spec._tensor("prev", n, ph, pw, c)
    t_out = spec._tensor("out", n, 2 * ph, 2 * pw, c)
    spec._op(_lib.OP_UPADD, in0=t_skip, in1=t_prev, out=t_out)
    ctx = Context(0, precision)
    plan = ForwardPlan(ctx, MiniModel(BlobBuilder()), 0, 0, 0, 0, 0, spec=spec)

    def put(tid, a):
        hi = a.astype(np.float16)
        plan.write(tid, hi, 0)
        if precision == "f16x2":
            plan.write(tid, (a - hi.astype(np.float32)).astype(np.float16), 1)

    put(t_skip, skip)
    put(t_prev, prev)
    plan.run()
    got = plan.read(t_out).astype(np.float32)
    up = F.interpolate(torch.from_numpy(nhwc_to_nchw(prev)).double(), scale_factor=2, mode="bilinear",
                       align_corners=False)
    ref = nchw_to_nhwc((torch.from_numpy(nhwc_to_nchw(skip)).double() + up).numpy())
    # fp16 mode interpolates in packed half arithmetic (upadd_math.cuh): four roundings of 2^-11 on
    # terms of magnitude <= max|prev| before the final rounding of the sum
    tol = (1e-3 * np.abs(ref) + 2.5e-3 * np.abs(prev).max()) if precision == "f16" \
        else (2e-6 * np.abs(ref) + 2e-6)
    assert np.all(np.abs(got - ref) <= tol), float(np.abs(got - ref).max())
    plan.close()
    ctx.close()


@pytest.mark.parametrize("precision", ["f16", "f16x2"])
@pytest.mark.parametrize("shape", [(2, 32, 32, 64), (1, 30, 50, 64), (3, 17, 23, 64), (33, 8, 8, 64), (2, 16, 16, 128)])
def test_maxpool_matches_torch(shape, precision, built_lib):
    """3x3 stride-2 pad-1 max-pool (models/backbone/resnet.py:201), exact in both precision modes:
    64 channels = the tiled fp16 fast path, 128 channels / split mode = the generic kernel."""
    n, h, w, c = shape
    oh, ow = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    rng = np.random.RandomState(9)
    x = rng.standard_normal((n, h, w, c)).astype(np.float32)
    if precision == "f16":
        x = f16(x).astype(np.float32)
    spec = MiniSpec()
    t_in = spec._tensor("in", n, h, w, c)
    t_out = spec._tensor("out", n, oh, ow, c)
    spec._op(_lib.OP_MAXPOOL, in0=t_in, out=t_out)
    ctx = Context(0, precision)
    plan = ForwardPlan(ctx, MiniModel(BlobBuilder()), 0, 0, 0, 0, 0, spec=spec)
    hi = x.astype(np.float16)
    plan.write(t_in, hi, 0)
    if precision == "f16x2":
        plan.write(t_in, (x - hi.astype(np.float32)).astype(np.float16), 1)
    plan.run()
    got = plan.read(t_out).astype(np.float32)  # hi + lo in split mode
    ref = nchw_to_nhwc(F.max_pool2d(torch.from_numpy(nhwc_to_nchw(x)), 3, 2, 1).numpy())
    if precision == "f16":
        assert np.array_equal(got, ref)
    else:
        assert np.all(np.abs(got - ref) <= 2e-6 * np.abs(ref) + 2e-6)
    plan.close()
    ctx.close()
