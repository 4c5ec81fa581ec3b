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
