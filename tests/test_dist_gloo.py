"""CPU, world_size 2 over gloo: the multi-GPU plumbing of the hot path —
rank-strided sharding of independent units and the one-off broadcast of the packed weights."""
import hashlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cerberus_b200 import synth
from cerberus_b200.dist import broadcast_packed_model, max_over_ranks, shard_units
from cerberus_b200.plan import PackedModel, PlanSpec


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    margs = synth.model_args(["Nuclei", "Patch-Class"])
    model = None
    if rank == 0:
        model = PackedModel(synth.make_state_dict(margs["considered_tasks"], seed=0), margs)
    model = broadcast_packed_model(model, margs, rank, world, torch.device("cpu"))
    spec = PlanSpec(model, 4, 256, 256, 256, 256)
    mine = shard_units(11, rank, world)
    t = max_over_ranks(1.0 + rank, world, torch.device("cpu"))
    rec = {"sha": hashlib.sha1(model.blob.tobytes()).hexdigest(), "ops": len(spec.ops),
           "flops": spec.conv_flops(), "units": mine, "tmax": t, "idx": dict(model.idx_dict)}
    torch.save(rec, os.path.join(out_dir, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_share_weights_and_split_units(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(str(tmp_path), "r0.pt"))
    r1 = torch.load(os.path.join(str(tmp_path), "r1.pt"))
    assert r0["sha"] == r1["sha"] and r0["ops"] == r1["ops"] and r0["flops"] == r1["flops"]
    assert r0["idx"] == r1["idx"]
    assert sorted(r0["units"] + r1["units"]) == list(range(11))
    assert not set(r0["units"]) & set(r1["units"])
    assert r0["tmax"] == r1["tmax"] == 2.0


def test_shard_units_edge_cases():
    assert shard_units(0, 0, 4) == []
    assert shard_units(3, 3, 4) == []
    assert shard_units(9, 1, 4) == [1, 5]
    allu = sorted(u for r in range(8) for u in shard_units(100, r, 8))
    assert allu == list(range(100))
