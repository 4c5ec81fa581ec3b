"""Multi-GPU plumbing (one process per GPU, torch.distributed): the hot path shards by
independent units (patch batches / images), so the only data-path collective is the one-off
broadcast of the packed weight blob from rank 0 (replaces nn.DataParallel's
per-forward replicate/scatter/gather of infer/base.py:46)."""
import numpy as np
import torch
import torch.distributed as dist

from .plan import PackedModel


def shard_units(n_units, rank, world):
    """Rank-strided assignment of independent units (batches, files): r, r+P, r+2P, ..."""
    return list(range(rank, n_units, world))


def broadcast_packed_model(model, margs, rank, world, device):
    """Rank 0 holds a PackedModel; every other rank receives an identical one.
    `device`: torch.device used for the collective ("cuda:N" with NCCL, "cpu" with gloo)."""
    if world == 1:
        return model
    nbytes = torch.tensor([model.blob.nbytes if rank == 0 else 0], dtype=torch.int64, device=device)
    dist.broadcast(nbytes, 0)
    if rank == 0:
        blob_t = torch.from_numpy(model.blob).to(device)
        meta = [{"layers": model.layers, "idx": model.idx_dict, "canvas_c": model.canvas_c,
                 "seg": model.seg_decoders, "pc": model.has_pclass,
                 "dk": model.decoder_kwargs, "tasks": model.considered_tasks,
                 "blocks": getattr(model, "blocks", [3, 4, 6, 3]), "backbone": getattr(model, "backbone", "resnet34")}]
    else:
        blob_t = torch.empty(int(nbytes.item()), dtype=torch.uint8, device=device)
        meta = [None]
    dist.broadcast(blob_t, 0)
    dist.broadcast_object_list(meta, 0)
    if rank != 0:
        m = meta[0]
        model = PackedModel.__new__(PackedModel)
        model.decoder_kwargs, model.considered_tasks = m["dk"], m["tasks"]
        model.layers, model.idx_dict, model.canvas_c = m["layers"], m["idx"], m["canvas_c"]
        model.seg_decoders, model.has_pclass = m["seg"], m["pc"]
        model.blocks, model.backbone = m["blocks"], m["backbone"]
        model.blob = np.ascontiguousarray(blob_t.cpu().numpy())
    return model


def max_over_ranks(value, world, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
