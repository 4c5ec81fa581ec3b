"""GPU: on-device post-processing is bit-exact with the reference (golden vectors produced by
the unmodified loader/postproc.py code) and with the oracle on further seeded inputs, through
the C ABI (cerb_postproc_nuclei / cerb_postproc_gland_lumen) and the drop-in class."""
import numpy as np
import pytest

from cerberus_b200 import synth
from cerberus_b200.engine import Context
from cerberus_b200.postproc import PostProcInstErodedContourMap, post_process_batch
from oracle import postproc_oracle as po
from tests.test_oracle_postproc import IDX, golden_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built_lib):
    c = Context(0, "f16")
    PostProcInstErodedContourMap.bind(c)
    yield c
    c.close()


def test_postproc_matches_reference_golden(ctx):
    bad = []
    for key, tissue, ds, field, inst, dtype in golden_cases():
        raw = np.zeros(field.shape[:2] + (6,), np.float32)
        lo = IDX[tissue + "-INST"][0]
        raw[..., lo:lo + 2] = field
        got, _ = PostProcInstErodedContourMap.post_process(raw, IDX, tissue, ds)
        if str(got.dtype) != dtype or not np.array_equal(got.astype(np.int64), inst):
            bad.append((key, str(got.dtype), dtype, int((got.astype(np.int64) != inst).sum())))
    assert not bad, bad


@pytest.mark.parametrize("tissue", ["Nuclei", "Gland", "Lumen"])
def test_postproc_batch_matches_oracle(ctx, tissue):
    """A batch of 8 different 256x256 fields in one call (images are independent)."""
    fields = [synth.postproc_field(256, 256, tissue, seed=100 + s) for s in range(6)]
    fields += [np.tile(synth.adversarial_field(64, 64, seed=s, big=(tissue != "Nuclei")), (4, 4, 1))
               for s in range(2)]
    canvas = np.zeros((8, 256, 256, 9), np.float32)
    ch0 = {"Lumen": 0, "Gland": 2, "Nuclei": 4}[tissue]
    for i, f in enumerate(fields):
        canvas[i, ..., ch0:ch0 + 2] = f
    got, any_fg = post_process_batch(ctx, canvas, ch0, tissue, 1.0)
    for i, f in enumerate(fields):
        if tissue == "Nuclei":
            ref = po.proc_nuclei(f)
        else:
            ref = po.proc_gland_lumen(f, tissue, 1.0)
        assert np.array_equal(got[i].astype(np.int64), ref.astype(np.int64)), (tissue, i)


def test_postproc_large_image_matches_oracle(ctx):
    """1000 x 1200 nuclei field: the sequential heap spills from shared to global memory."""
    f = synth.postproc_field(1000, 1200, "Nuclei", seed=5)
    canvas = np.zeros((1, 1000, 1200, 2), np.float32)
    canvas[0] = f
    got, _ = post_process_batch(ctx, canvas, 0, "Nuclei", 1.0)
    assert np.array_equal(got[0], po.proc_nuclei(f))
    g = synth.postproc_field(700, 900, "Gland", seed=6)
    canvas = np.zeros((1, 700, 900, 2), np.float32)
    canvas[0] = g
    got, _ = post_process_batch(ctx, canvas, 0, "Gland", 1.0)
    assert np.array_equal(got[0].astype(np.int64), po.proc_gland_lumen(g, "Gland", 1.0).astype(np.int64))


def test_idempotence_property(ctx):
    """Size-independent property: feeding the binarised result back (inner = labels > 0,
    contour = 0) cannot create or merge gland instances' supports beyond a dilation."""
    f = synth.postproc_field(512, 512, "Lumen", seed=9)
    canvas = np.zeros((1, 512, 512, 2), np.float32)
    canvas[0] = f
    a, _ = post_process_batch(ctx, canvas, 0, "Lumen", 1.0)
    # every thresholded foreground pixel that survived the size filter is covered by a label
    fg = (f[..., 0] - (f[..., 1] > 0.5)) > 0.5
    kept = po.remove_small_objects(fg, 150)
    assert np.all(a[0][kept] > 0)
    # labels are exactly 1..K
    ids = np.unique(a[0])
    assert np.array_equal(ids, np.arange(ids.size))


def _tie_field(order_matters, seed, size=320, equal_pushes=True):
    """320 x 320 (> 65536 px: the large-image watershed) fields whose every value is unique except
    for engineered marker ties. Clusters = two 10 x 10 marker squares A, B in one mask rectangle.
    order_matters=False: the tied marker pixels sit on the far sides of A and B (and a three-way
    tie A, A, B) - every pop order of the tied entries gives the same labels. Two clusters are
    provably harmless inside the flood (ws_tie_harmless), the third needs the enumeration.
    order_matters=True: A and B are one pixel apart and the tied pixels face each other with the
    globally smallest key: whichever pops first labels the gap pixel."""
    rng = np.random.RandomState(seed)
    H = W = size
    inner = np.zeros((H, W), np.float32)
    cnt = np.zeros((H, W), np.float32)
    hi_vals = (0.6 + rng.permutation(70000) * 2.0 ** -18).astype(np.float32)   # markers, unique
    lo_vals = (0.05 + rng.permutation(100000) * 2.0 ** -18).astype(np.float32)  # gaps, unique
    hi_i = lo_i = 0
    ties = []
    origins = [(20, 30), (20, 180), (120, 60), (200, 30), (220, 200)] if size >= 320 else \
        [(10, 10), (10, 120), (60, 40), (110, 10), (130, 120)]
    for ci, (y0, x0) in enumerate(origins):
        gap = 1 if order_matters else 6
        h, w = 24, 10 + gap + 10 + 8
        cnt[y0:y0 + h, x0:x0 + w] = 0.6
        reg = inner[y0:y0 + h, x0:x0 + w]
        reg[...] = lo_vals[lo_i:lo_i + h * w].reshape(h, w)
        lo_i += h * w
        ay, ax = y0 + 7, x0 + 4
        by, bx = y0 + 7, x0 + 4 + 10 + gap
        for (sy, sx) in ((ay, ax), (by, bx)):
            inner[sy:sy + 10, sx:sx + 10] = hi_vals[hi_i:hi_i + 100].reshape(10, 10)
            hi_i += 100
        if ci >= 3:
            continue  # clusters without a tie
        if order_matters:
            pix = [(ay + 5, ax + 9), (by + 5, bx)]          # facing each other across the gap pixel
            v = np.float32(0.97 + ci * 0.001)               # larger than every other inner value
        else:
            pix = [(ay + 5, ax), (by + 5, bx + 9)]          # far sides
            if ci == 1:
                pix.append((ay, ax + 5))                    # three-way tie: 6 orders
            if ci == 2 and equal_pushes:
                # the pixels the two tied markers push first share a value: not provably harmless
                # (their ages swap with the pop order), so this cluster goes through k_wsg_certify
                inner[ay + 5, ax - 1] = inner[by + 5, bx + 10] = np.float32(0.3)
            v = np.float32(0.93 + ci * 0.001)
        for (y, x) in pix:
            inner[y, x] = v
        ties.append(pix)
    return np.stack([inner, cnt], axis=-1), ties


def test_marker_ties_are_settled_by_enumerating_pop_orders(ctx):
    lib, h = ctx.lib, ctx.handle
    stat = lambda name: lib.cerb_ctx_stat(h, name)  # noqa: E731
    for order_matters in (False, True):
        f, ties = _tie_field(order_matters, seed=3 + order_matters)
        canvas = np.zeros((1, 320, 320, 2), np.float32)
        canvas[0] = f
        before = (stat(b"ws_large_images"), stat(b"ws_large_tied_components"), stat(b"ws_large_fallbacks"))
        got, any_fg = post_process_batch(ctx, canvas, 0, "Nuclei", 1.0)
        after = (stat(b"ws_large_images"), stat(b"ws_large_tied_components"), stat(b"ws_large_fallbacks"))
        ref = po.proc_nuclei(f)
        assert ref.max() == 10 and np.array_equal(got[0], ref), order_matters
        assert after[0] - before[0] == 1
        # components handed to k_wsg_certify: all three facing ties / only the cluster whose tie the
        # flood cannot prove harmless on its own
        assert len(ties) == 3 and after[1] - before[1] == (3 if order_matters else 1), (before, after)
        # harmless ties never reach the whole-image emulation; facing ties must
        assert after[2] - before[2] == (1 if order_matters else 0), (before, after)
    # the facing-tie field really is order dependent: swapping which marker owns the gap pixel
    # changes the oracle's answer, so the fallback above was needed
    f, ties = _tie_field(True, seed=4)
    ref = po.proc_nuclei(f)
    (ya, xa), (yb, xb) = ties[0]
    assert xb - xa == 2 and ref[ya, xa + 1] in (ref[ya, xa], ref[yb, xb]) and ref[ya, xa] != ref[yb, xb]


def test_eroded_map_postproc_matches_reference_golden(ctx):
    """PostProcInstErodedMap (IP-ERODED-3/11 codes) through cerb_postproc_eroded_map."""
    from cerberus_b200.postproc import PostProcInstErodedMap
    from tests.test_oracle_postproc import eroded_golden_cases
    bad = []
    n = 0
    for key, tissue, field, inst, dtype in eroded_golden_cases():
        raw = np.zeros(field.shape[:2] + (3,), np.float32)
        raw[..., 1:2] = field           # channel offset != 0
        raw[..., 2] = 1.0               # a type channel: returned as a [H,W,1] slice
        got, typ = PostProcInstErodedMap.post_process(raw, {tissue + "-INST": [1, 2], tissue + "-TYPE": [2, 3]},
                                                      tissue)
        assert typ.shape == field.shape[:2] + (1,)
        if str(got.dtype) != dtype or not np.array_equal(got.astype(np.int64), inst):
            bad.append((key, int((got.astype(np.int64) != inst).sum())))
        n += 1
    assert n >= 70 and not bad, bad
    # a larger field against the oracle, all three tissues
    f = synth.postproc_field(600, 700, "Gland", seed=11)[..., :1]
    for tissue in ("Gland", "Lumen", "Nuclei"):
        got, _ = PostProcInstErodedMap.post_process(f, {tissue + "-INST": [0, 1]}, tissue)
        ref, _ = po.post_process_eroded(f, {tissue + "-INST": [0, 1]}, tissue)
        assert ref.max() > 0 and np.array_equal(got, ref), tissue
    with pytest.raises(ValueError):
        PostProcInstErodedMap.post_process(np.zeros((8, 8, 2), np.float32), {"Gland-INST": [0, 2]}, "Gland")


def test_small_tile_marker_ties_fall_back_only_when_the_order_can_matter(ctx):
    """<= 65536 px: k_watershed_comp proves most ties harmless inside the flood; facing ties (and
    ties it cannot prove harmless) go to the exact whole-tile emulation. Bit-exact either way."""
    lib, h = ctx.lib, ctx.handle
    stat = lambda name: lib.cerb_ctx_stat(h, name)  # noqa: E731
    for order_matters, equal_pushes, expect in ((False, False, 0), (False, True, 1), (True, False, 1)):
        f, ties = _tie_field(order_matters, seed=7 + order_matters, size=240, equal_pushes=equal_pushes)
        canvas = np.zeros((1, 240, 240, 2), np.float32)
        canvas[0] = f
        before = (stat(b"ws_images"), stat(b"ws_tie_fallbacks"))
        got, _ = post_process_batch(ctx, canvas, 0, "Nuclei", 1.0)
        after = (stat(b"ws_images"), stat(b"ws_tie_fallbacks"))
        ref = po.proc_nuclei(f)
        assert ref.max() == 10 and np.array_equal(got[0], ref), (order_matters, equal_pushes)
        assert after[0] - before[0] == 1 and after[1] - before[1] == expect, (order_matters, equal_pushes, before, after)


def test_oversized_gland_crops_use_the_global_memory_path(ctx):
    """A merged gland whose padded bounding box (~1300 x 1250 px) needs more bit-plane words than
    fit in shared memory: the reference handles any size (loader/postproc.py:292-307), so must the
    device (global-memory planes instead of the former error 11). Two such giants plus ordinary
    glands with holes, at ds 1.0 and ds 0.5, bit-exact vs the oracle."""
    H, W = 1500, 1400
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    rng = np.random.RandomState(3)
    inner = np.zeros((H, W), np.float32)
    # giant ring-shaped gland with lobes (bbox ~1250 x 1200) and a second big one nested in its lumen
    r = np.hypot(yy - 760, xx - 690)
    inner[(r < 600 + 25 * np.sin(xx / 40.0)) & (r > 420)] = 0.97
    inner[np.hypot(yy - 760, xx - 690) < 330] = 0.93
    for _ in range(40):  # holes that must be filled, small glands, specks below the size filter
        cy, cx, rad = rng.randint(60, H - 60), rng.randint(60, W - 60), rng.randint(6, 28)
        inner[np.hypot(yy - cy, xx - cx) < rad] = 0.0 if rng.rand() < 0.5 else 0.9
    contour = np.zeros_like(inner)
    field = np.stack([inner, contour], -1)
    for tissue in ("Gland", "Lumen"):
        for ds in (1.0, 0.5):
            got, _ = post_process_batch(ctx, field[None], 0, tissue, ds)
            ref = po.proc_gland_lumen(field, tissue, ds)
            assert int(ref.max()) >= 3
            assert np.array_equal(got[0].astype(np.int64), ref.astype(np.int64)), (tissue, ds)
