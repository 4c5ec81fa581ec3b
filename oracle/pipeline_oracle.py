"""ORACLE — test infrastructure only. CPU restatement of the per-tile tail of the reference
path for tiles whose network output covers the whole tile (in == out):
infer/tile.py:116-191 (_post_process_patches: channel table, stitch of a single patch,
post_process per tissue, lumen *= gland > 0). Used by bench.py's CPU legs (`cpu_baseline`,
`--impl reference`) and by tests."""
import numpy as np

from oracle import postproc_oracle as po

TISSUES = ("Nuclei", "Gland", "Lumen")


def canvas_of(sample, idx_dict, nr_ch):
    """infer/tile.py:136-163 for one patch that covers the image (count == 1)."""
    any_v = next(iter(sample.values()))
    h, w = any_v.shape[:2]
    raw = np.zeros((h, w, nr_ch), dtype=np.float32)
    for k, v in sample.items():
        lo, hi = idx_dict[k]
        raw[..., lo:hi] = v if v.ndim == 3 else v[..., None]
    return raw


def postprocess_step(step, margs):
    """step: list of per-sample dicts (models/run_desc.py:494-502). Returns per-sample
    {tissue: inst_map}."""
    from cerberus_b200.plan import canvas_layout
    idx_dict, nr_ch = canvas_layout(margs["decoder_kwargs"])
    out = []
    for sample in step:
        raw = canvas_of(sample, idx_dict, nr_ch)
        maps = {}
        for t in TISSUES:
            if t + "-INST" in sample:
                maps[t], _ = po.post_process(raw, idx_dict, t, 1.0)
        if "Gland" in maps and "Lumen" in maps:
            maps["Lumen"] = (maps["Gland"] > 0) * maps["Lumen"]
        out.append(maps)
    return out
