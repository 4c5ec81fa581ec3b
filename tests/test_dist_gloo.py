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


class _FakeCanvas:
    shape = (700, 1000, 9)


def _wsi_worker(rank, world, port, out_dir):
    """WSI mode: the post-processing tiles of a set are strided over ranks and merged on rank 0
    (cerberus_b200/infer/wsi.py::_postproc_nuclei); the device call is replaced by a deterministic
    stand-in so that the sharding / gather / merge logic runs on CPU."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from cerberus_b200.infer.wsi import InferManager
    from cerberus_b200.infer.wsi_geometry import get_coordinates, select_tile_instances
    m = object.__new__(InferManager)
    m.patch_output_shape = [144, 144]
    calls = []

    def fake_tile(canvas, tile_bounds, tile_flag, tile_mode, ref_boxes, margin):
        tb = np.asarray(tile_bounds)
        calls.append(tuple(int(v) for v in tb))
        rng = np.random.RandomState(int(tb.sum()) % 100000 + 7 * tile_mode)
        w, h = int(tb[2] - tb[0]), int(tb[3] - tb[1])
        xy = np.stack([rng.randint(0, max(w - 12, 1), 40), rng.randint(0, max(h - 12, 1), 40)], -1)
        boxes = np.concatenate([xy, xy + rng.randint(4, 12, (40, 2))], -1)
        sel, sel_ref = select_tile_instances(boxes, tb, tile_flag, tile_mode, margin,
                                             ref_boxes if tile_mode == 3 else None)
        keep = np.array([k not in set(sel) for k in range(len(boxes))])
        kb = boxes[keep] + np.concatenate([tb[:2]] * 2)
        n = len(kb)
        # columns of dat_writer.InstanceStore: box, centroid, contour offsets, contour points, prob, type
        cols = (kb, kb[:, :2].astype(np.float64), np.arange(n + 1) * 2, np.repeat(kb[:, :2], 2, axis=0),
                np.full(n, 0.5), np.full(n, tile_mode))
        return cols, sel_ref  # indices into the rows accumulated so far

    class _Model:
        idx_dict = {"Nuclei-TYPE": [6, 7]}

    class _Eng:
        model = _Model()

    m.engine = _Eng()
    # the two halves of a post-processing tile (watershed on the main context / instance table on
    # the table thread's context) and the table context itself are replaced
    m._tile_labels = lambda canvas, tb: tb
    m._tile_tables = lambda ctx, labelled, tb, flag, mode, ref, margin: fake_tile(None, tb, flag, mode, ref, margin)
    m._table_ctx = lambda: None
    _, pout = get_coordinates((1000, 700), [448, 448], [144, 144], [144, 144])
    store = m._postproc_nuclei(_FakeCanvas(), pout, [300, 300], 64)
    keys = []
    if store is not None:
        box, _, _, _, _, typ = store.columns()
        keys = sorted("m%d_%s" % (t, "_".join(map(str, b))) for b, t in zip(box.tolist(), typ.tolist()))
    torch.save({"keys": keys, "n_calls": len(calls)}, os.path.join(out_dir, "w%d_r%d.pt" % (world, rank)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_wsi_postproc_tiles_shard_over_ranks_and_merge_like_one_rank(tmp_path):
    _wsi_worker(0, 1, _free_port(), str(tmp_path))
    mp.spawn(_wsi_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    one = torch.load(os.path.join(str(tmp_path), "w1_r0.pt"))
    r0 = torch.load(os.path.join(str(tmp_path), "w2_r0.pt"))
    r1 = torch.load(os.path.join(str(tmp_path), "w2_r1.pt"))
    assert len(one["keys"]) > 100
    assert r0["keys"] == one["keys"]            # rank 0 ends up with the single-rank result
    assert r1["keys"] == []                      # other ranks only contribute
    assert r0["n_calls"] + r1["n_calls"] == one["n_calls"] and r1["n_calls"] > 0
