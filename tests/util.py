"""Test helpers: ad-hoc plans (single ops) over the C ABI."""
import numpy as np

from cerberus_b200 import _lib
from cerberus_b200.pack import BlobBuilder, pack_conv, pack_stem
from cerberus_b200.plan import PlanSpec


class MiniSpec(PlanSpec):
    """A PlanSpec whose tensors/ops are filled by hand."""

    def __init__(self):
        self.tensors, self.ops, self.named = [], [], {}
        from collections import OrderedDict
        self.logit_tensors = OrderedDict()
        self.canvas = -1


class MiniModel:
    def __init__(self, blob):
        self.blob = blob.finish() if isinstance(blob, BlobBuilder) else blob


def f16(a):
    return np.asarray(a, dtype=np.float32).astype(np.float16)


def nhwc_to_nchw(a):
    return np.ascontiguousarray(np.transpose(a, (0, 3, 1, 2)))


def nchw_to_nhwc(a):
    return np.ascontiguousarray(np.transpose(a, (0, 2, 3, 1)))
