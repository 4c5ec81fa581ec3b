"""CPU: the op graph cerb_model_spec builds inside the library (csrc/model.cu) equals the one
cerberus_b200/plan.py::PlanSpec builds in Python - tensor for tensor, op field for op field -
for the six-head model, a single-decoder model and a model without Patch-Class, at several batch
shapes. (No device call: cerb_model_spec is pure host code.)"""
import ctypes

import pytest

from cerberus_b200 import _lib, synth
from cerberus_b200.plan import PackedModel, PlanSpec, c_model_tables

CASES = [
    (None, 2, 256, 256, 256, 256, False),
    (None, 1, 448, 448, 144, 144, True),
    (None, 3, 64, 96, 48, 80, False),
    (["Nuclei"], 1, 256, 256, 256, 256, True),
    (["Gland", "Gland#TYPE", "Lumen"], 4, 128, 128, 128, 128, False),
    (["Patch-Class"], 2, 256, 256, 16, 16, True),
]


@pytest.fixture(scope="module")
def models():
    out = {}
    for tasks in {tuple(c[0]) if c[0] else None for c in CASES}:
        t = list(tasks) if tasks else None
        args = synth.model_args(t)
        out[tasks] = PackedModel(synth.make_state_dict(args["considered_tasks"], seed=0), args)
    return out


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%dx%dx%d" % ("all" if c[0] is None else "+".join(c[0]), c[1], c[2], c[3]))
def test_library_graph_equals_python_graph(case, models, built_lib):
    tasks, n, h, w, oh, ow, logits = case
    model = models[tuple(tasks) if tasks else None]
    spec = PlanSpec(model, n, h, w, oh, ow, want_logits=logits)
    td_py, ops_py = spec.c_arrays()
    desc, layers = c_model_tables(model)
    cap_t, cap_o = ctypes.c_int(256), ctypes.c_int(512)
    td = (_lib.TensorDesc * 256)()
    ops = (_lib.Op * 512)()
    canvas = ctypes.c_int32(-1)
    lg = (ctypes.c_int32 * (_lib.MAX_DECODERS + 1))()
    _lib.check(built_lib.cerb_model_spec(ctypes.byref(desc), layers, len(layers), n, h, w, oh, ow,
                                         int(logits), td, ctypes.byref(cap_t), ops, ctypes.byref(cap_o),
                                         ctypes.byref(canvas), lg), "cerb_model_spec")
    assert cap_t.value == len(td_py) and cap_o.value == len(ops_py)
    assert canvas.value == spec.canvas
    for i in range(len(td_py)):
        for f, _ in _lib.TensorDesc._fields_:
            assert getattr(td[i], f) == getattr(td_py[i], f), ("tensor", i, f)
    for i in range(len(ops_py)):
        for f, _ in _lib.Op._fields_:
            assert getattr(ops[i], f) == getattr(ops_py[i], f), ("op", i, f, getattr(ops[i], f), getattr(ops_py[i], f))
    # logit tensors: decoders in order, Patch-Class last
    want = [spec.logit_tensors.get(k, -1) for k in
            [__import__("cerberus_b200.plan", fromlist=["HEAD_NAME_MAP"]).HEAD_NAME_MAP[d] for d in model.seg_decoders]]
    assert [lg[i] for i in range(len(want))] == want
    assert lg[_lib.MAX_DECODERS] == spec.logit_tensors.get("Patch-Class", -1)


def test_spec_rejects_bad_shapes(models, built_lib):
    model = models[None]
    desc, layers = c_model_tables(model)
    n_t, n_o = ctypes.c_int(0), ctypes.c_int(0)
    rc = built_lib.cerb_model_spec(ctypes.byref(desc), layers, len(layers), 1, 250, 256, 250, 256, 0,
                                   None, ctypes.byref(n_t), None, ctypes.byref(n_o), None, None)
    assert rc == -2 and b"multiple of 16" in built_lib.cerb_last_error()
    rc = built_lib.cerb_model_spec(ctypes.byref(desc), layers, len(layers) - 1, 1, 256, 256, 256, 256, 0,
                                   None, ctypes.byref(n_t), None, ctypes.byref(n_o), None, None)
    assert rc == -2 and b"layer table lacks" in built_lib.cerb_last_error()


def test_resnet18_graph_equals_python_graph(built_lib):
    """The resnet18 encoder (2, 2, 2, 2 BasicBlocks; models/backbone/resnet.py:292-302): the library
    takes the block count per stage from the layer table."""
    args = synth.model_args(["Nuclei", "Gland#TYPE", "Patch-Class"], backbone="resnet18")
    model = PackedModel(synth.make_state_dict(args["considered_tasks"], seed=0, backbone="resnet18"), args)
    assert model.blocks == [2, 2, 2, 2]
    spec = PlanSpec(model, 2, 256, 256, 256, 256, want_logits=True)
    td_py, ops_py = spec.c_arrays()
    desc, layers = c_model_tables(model)
    cap_t, cap_o = ctypes.c_int(256), ctypes.c_int(512)
    td = (_lib.TensorDesc * 256)()
    ops = (_lib.Op * 512)()
    _lib.check(built_lib.cerb_model_spec(ctypes.byref(desc), layers, len(layers), 2, 256, 256, 256, 256, 1,
                                         td, ctypes.byref(cap_t), ops, ctypes.byref(cap_o), None, None),
               "cerb_model_spec")
    assert cap_t.value == len(td_py) and cap_o.value == len(ops_py)
    n_conv = sum(1 for o in ops_py if o.kind == _lib.OP_CONV)
    assert n_conv == 1 + 2 * 8 + 3 + 1 + 1 + 2 * 7 + 2  # stem, 8 blocks, 3 downsamples, conv_map, first, 2 x 7 decoder convs, 2 heads
    for i in range(len(ops_py)):
        for f, _ in _lib.Op._fields_:
            assert getattr(ops[i], f) == getattr(ops_py[i], f), ("op", i, f)


def _folding(built_lib, spec, precision=_lib.CERB_PREC_F16):
    td, ops = spec.c_arrays()
    folded = (ctypes.c_int32 * len(ops))()
    _lib.check(built_lib.cerb_plan_preview_folding(precision, td, len(td), ops, len(ops), folded),
               "cerb_plan_preview_folding")
    return [int(v) for v in folded]


def test_upadd_folding_rule(models, built_lib):
    """cerb_plan_create folds an UPADD into the 64->64 3x3 convolution that follows it (conv64x.cu's
    fused producer) iff that convolution is the only reader of the sum - decided by pure host code,
    checked here without a device. Six-head model: the 128^2 and 256^2 levels of every segmentation
    decoder (2 x 5), not the 128 / 256-channel levels; the sum tensors are REUSED by the decoders,
    which must not count as a second reader; nothing is folded in the split-precision mode."""
    model = models[None]
    spec = PlanSpec(model, 2, 256, 256, 256, 256)
    folded = _folding(built_lib, spec)
    ups = [i for i, op in enumerate(spec.ops) if op["kind"] == _lib.OP_UPADD]
    assert len(ups) == 1 + 3 * 5
    n_dec = len(model.seg_decoders)
    assert sum(folded) == 2 * n_dec
    for i in ups:
        _, _, h, w, c, _ = spec.tensors[spec.ops[i]["out"]]
        assert folded[i] == int(c == 64), (i, h, w, c)
        if folded[i]:
            nxt = spec.ops[i + 1]
            assert nxt["kind"] == _lib.OP_CONV and nxt["in0"] == spec.ops[i]["out"] and nxt["cout"] == 64
    assert all(f == 0 for i, f in enumerate(folded) if i not in ups)
    assert sum(_folding(built_lib, spec, _lib.CERB_PREC_F16X2)) == 0
    # a second reader of the sum (here: a later op made to read it) keeps the stand-alone pass
    i0 = next(i for i in ups if folded[i])
    spec2 = PlanSpec(model, 2, 256, 256, 256, 256)
    victim = next(k for k in range(i0 + 2, len(spec2.ops)) if spec2.ops[k]["kind"] == _lib.OP_CONV
                  and spec2.ops[k]["in1"] < 0 and spec2.ops[k]["out"] != spec2.ops[i0]["out"])
    spec2.ops[victim]["in1"] = spec2.ops[i0]["out"]
    folded2 = _folding(built_lib, spec2)
    assert folded2[i0] == 0 and sum(folded2) == 2 * n_dec - 1
