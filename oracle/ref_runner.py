"""ORACLE — test infrastructure only: the CPU arm of bench.py running the UNMODIFIED reference.

Loads the reference modules (from /root/reference in the build container, from the staged copy
oracle/_ref/ on the GPU box: oracle/make_ref.py) through the shim of oracle/ref_shim.py and
exposes the tile path as the reference's own code runs it:

    net      = models.net_desc.create_model(**model_args); load_state_dict(strict=True)
    step     = models.run_desc.infer_step(batch, net, out, considered_tasks)     (run_desc.py:439-502)
    labels   = loader.postproc.PostProcInstErodedContourMap.post_process(...)    (postproc.py:383-407)
               + lumen *= gland > 0                                              (infer/tile.py:187-191)
    plumbing = infer.tile._prepare_patching / _post_process_patches              (tile.py:43-212)

scikit-image (un-vendored, absent) is the one restated piece: remove_small_objects / watershed
come from oracle/postproc_oracle.py (compiled C), i.e. FASTER than the reference's Cython/Python,
so the CPU number is if anything flattering to the reference.
"""
import os

import numpy as np

from oracle import ref_shim

_HERE = os.path.dirname(os.path.abspath(__file__))


def ref_dir():
    for d in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isdir(os.path.join(d, "models")) and os.path.isdir(os.path.join(d, "infer")):
            return d
    return None


def available():
    return ref_dir() is not None


class ReferenceTilePath:
    """The reference's own modules on the host CPU."""

    def __init__(self, sd, margs):
        import torch
        ref_shim.install(ref=ref_dir())
        from models.net_desc import create_model
        from models.run_desc import infer_step
        from loader.postproc import PostProcInstErodedContourMap
        import infer.tile as ref_tile
        self.margs = margs
        self.net = create_model(**margs)
        self.net.load_state_dict(sd, strict=True)
        self.net.eval()
        self._infer_step = infer_step
        self._post = PostProcInstErodedContourMap
        self.tile = ref_tile
        self.torch = torch

    def infer_step(self, tiles_u8, out):
        return self._infer_step(self.torch.from_numpy(np.asarray(tiles_u8)), self.net, out,
                                self.margs["considered_tasks"])

    def step_with_labels(self, tiles_u8):
        """Bench workload (in == out tiles): forward + label maps of every tile."""
        from cerberus_b200.plan import canvas_layout
        from oracle.pipeline_oracle import canvas_of
        idx_dict, nr_ch = canvas_layout(self.margs["decoder_kwargs"])
        step = self.infer_step(tiles_u8, tiles_u8.shape[1])
        out = []
        for sample in step:
            raw = canvas_of(sample, idx_dict, nr_ch)
            maps = {}
            for t in ("Nuclei", "Gland", "Lumen"):
                if t + "-INST" in sample:
                    maps[t], _ = self._post.post_process(raw, idx_dict, t)
            if "Gland" in maps and "Lumen" in maps:
                g = maps["Gland"].copy()
                g[g > 0] = 1
                maps["Lumen"] = g * maps["Lumen"]
            out.append(maps)
        return out

    def process_image(self, img, in_size, out_size, postproc_code, postproc_list):
        """run_infer_tile.py plumbing for ONE image, batch 1 (infer/tile.py:294-405 without the
        DataLoader / file writing): _prepare_patching -> infer_step per patch ->
        _post_process_patches."""
        padded, info, src_pos = self.tile._prepare_patching(img, in_size, out_size, 0)
        outs = []
        for k in range(info.shape[0]):
            (y0, x0), (y1, x1) = info[k, 0]
            patch = padded[y0:y1, x0:x1][None]
            pdata = self.infer_step(patch, out_size)[0]
            outs.append((pdata, (info[k, 1, 0], info[k, 1, 1]), 0))
        image_info = {"src_pos": src_pos, "src_shape": img.shape[:2], "src_image": img, "name": "x"}
        return self.tile._post_process_patches(outs, image_info, postproc_code, postproc_list, self.margs)
