#!/usr/bin/env python
"""Multi-GPU check of the C-ABI weight broadcast (cerb_nccl_* + cerb_bcast_weights), run as
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bcast_check.py
Rank 0 packs the checkpoint and creates the model; the other ranks receive description, layer
table and blob over NCCL (communicator created through the library from an ncclUniqueId that
travels over a gloo broadcast), then every rank runs the same batch and the canvases must be
bit-identical. Prints one JSON line on rank 0."""
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cerberus_b200 import _lib, synth  # noqa: E402
from cerberus_b200.engine import CModel, Context  # noqa: E402
from cerberus_b200.plan import PackedModel  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    lib = _lib.load()
    ctx = Context(local, "f16")
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (ctypes.c_uint8 * 128)()
        _lib.check(lib.cerb_nccl_unique_id(buf), "cerb_nccl_unique_id")
        uid = torch.from_numpy(np.frombuffer(buf, dtype=np.uint8).copy())
    dist.broadcast(uid, 0)
    idb = (ctypes.c_uint8 * 128).from_buffer_copy(uid.numpy().tobytes())
    comm = ctypes.c_void_p()
    _lib.check(lib.cerb_nccl_comm_create(ctx.handle, world, rank, idb, ctypes.byref(comm)), "comm_create")
    args = synth.model_args()
    canvas_c = 9
    if rank == 0:
        packed = PackedModel(synth.make_state_dict(seed=0), args)
        cm = CModel(ctx, packed)
        handle = cm.handle
    else:
        handle = ctypes.c_void_p()
    _lib.check(lib.cerb_bcast_weights(ctx.handle, ctypes.byref(handle), comm, 0, rank), "cerb_bcast_weights")
    if rank != 0:
        cm = CModel(ctx, None, handle=handle)
    tiles = synth.synthetic_tiles(2, 256, 256, seed=77)
    canvas = cm.forward(tiles, 256, 256, canvas_c)
    t = torch.from_numpy(canvas)
    allc = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allc, t)
    same = all(bool(torch.equal(allc[0], c)) for c in allc)
    if rank == 0:
        print(json.dumps({"check": "cerb_bcast_weights", "world": world, "canvases_identical": same,
                          "canvas_abs_sum": float(np.abs(canvas).sum())}))
    _lib.check(lib.cerb_nccl_comm_destroy(comm), "comm_destroy")
    cm.close()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if same else 1)


if __name__ == "__main__":
    main()
