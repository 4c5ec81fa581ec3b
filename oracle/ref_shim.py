"""ORACLE tooling — imports the UNMODIFIED reference from /root/reference in THIS container.

Only oracle/gen_golden.py uses this module (to produce tests/golden/*); nothing that runs on
the GPU box may import it, because /root/reference does not exist there.

Recipe: stub the absent packages (matplotlib, skimage, termcolor,
docopt), shim the APIs removed from numpy 2 / scipy >= 1.12 (np.lib.pad, scipy.interp),
make `.to("cuda")` a no-op so the hard-coded device strings of models/run_desc.py:440 and
infer/base.py:47 run on CPU, and inject the restated scikit-image functions
(oracle/postproc_oracle.py: remove_small_objects, watershed — scikit-image 0.19.2 is an
un-vendored third-party dependency, environment.yml:20) into the skimage stub.
"""
import os
import sys
import types

REF = "/root/reference"


def install(skimage_impl=None, ref=None):
    """`ref`: directory holding the reference packages (default /root/reference; the CPU arm of
    bench.py passes the staged copy oracle/_ref on the GPU box)."""
    ref = ref or REF
    if not os.path.isdir(ref):
        raise RuntimeError("%s is not present: goldens can only be regenerated in the build container" % ref)
    sys.dont_write_bytecode = True
    import numpy as np
    import scipy
    import torch

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "matplotlib" not in sys.modules:
        mpl = stub("matplotlib")
        mpl.pyplot = stub("matplotlib.pyplot", get_cmap=lambda *a, **k: None)
        mpl.cm = stub("matplotlib.cm")
        mpl.lines = stub("matplotlib.lines", Line2D=object)
        stub("termcolor", colored=lambda s, *a, **k: s)
        stub("docopt", docopt=lambda *a, **k: {})
        sk = stub("skimage")
        sk.filters = stub("skimage.filters", rank=None, threshold_otsu=None)
        sk.morphology = stub("skimage.morphology", disk=None, remove_small_holes=None,
                             remove_small_objects=None)
        sk.segmentation = stub("skimage.segmentation", watershed=None)
        sk.color = stub("skimage.color")
        sk.exposure = stub("skimage.exposure")
        sk.measure = stub("skimage.measure")
    if skimage_impl is None:
        from oracle import postproc_oracle as skimage_impl
    sys.modules["skimage.morphology"].remove_small_objects = skimage_impl.remove_small_objects
    sys.modules["skimage.segmentation"].watershed = skimage_impl.watershed
    if "loader.postproc" in sys.modules:  # names were bound at import time
        sys.modules["loader.postproc"].watershed = skimage_impl.watershed
    if not hasattr(np.lib, "pad"):
        np.lib.pad = np.pad
    if not hasattr(scipy, "interp"):
        scipy.interp = np.interp
    if not getattr(torch.Tensor.to, "_cerb_cpu_shim", False):
        orig_to = torch.Tensor.to

        def to(self, *a, **k):
            if a and a[0] == "cuda":
                return self
            return orig_to(self, *a, **k)

        to._cerb_cpu_shim = True
        torch.Tensor.to = to
    if ref not in sys.path:
        sys.path.insert(0, ref)
