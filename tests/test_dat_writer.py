"""CPU: the `.dat` written by the array serialiser (csrc/dat_writer.cu via
cerberus_b200/infer/dat_writer.py) loads - with pickle AND joblib - as exactly the dict the
reference's object-by-object construction gives (infer/wsi.py:150,265,853)."""
import pickle

import joblib
import numpy as np

from cerberus_b200.infer.dat_writer import InstanceStore, unique_ids, write_dat


def _random_store(rng, n, has_type=True):
    s = InstanceStore(has_type)
    total = 0
    while total < n:
        m = int(rng.randint(1, 700))
        lens = rng.randint(0, 40, size=m)
        lens[rng.rand(m) < 0.05] = 300  # long contours: BININT shapes beyond one byte
        off = np.concatenate([[0], np.cumsum(lens)])
        s.append(rng.randint(-5, 70000, size=(m, 4)), rng.rand(m, 2) * 1e4, off,
                 rng.randint(0, 70000, size=(int(off[-1]), 2)),
                 rng.rand(m) if has_type else None, rng.randint(0, 7, size=m) if has_type else None)
        total += m
    return s


def _assert_same(a, b):
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert list(a[k].keys()) == list(b[k].keys()) == ["box", "centroid", "contour", "prob", "type"]
        for f in ("box", "centroid", "contour"):
            assert a[k][f].dtype == b[k][f].dtype and a[k][f].shape == b[k][f].shape, (k, f)
            assert np.array_equal(a[k][f], b[k][f]), (k, f)
        assert a[k]["prob"] == b[k]["prob"] and type(a[k]["prob"]) is type(b[k]["prob"])
        assert a[k]["type"] == b[k]["type"] and type(a[k]["type"]) is type(b[k]["type"])


def test_dat_roundtrip_equals_plain_pickle(tmp_path, built_lib):
    rng = np.random.RandomState(0)
    nuc = _random_store(rng, 3500)
    nuc.remove(rng.choice(len(nuc), 400, replace=False))  # cross tiles drop earlier instances
    gl = _random_store(rng, 30, has_type=False)
    uids = unique_ids(int(nuc.alive().sum()))
    want = {"Nuclei": nuc.to_dict(uids), "Gland": {"a": {"box": np.arange(4), "type": 1, "type_prob": 0.5}},
            "Lumen": gl, "proc_resolution": {"resolution": 0.5, "units": "mpp"},
            "proc_dimensions": np.array([20000, 20000]), "empty": InstanceStore()}
    info = dict(want)
    info["Nuclei"] = nuc
    path = str(tmp_path / "x.dat")
    # same uuids for the comparison
    import cerberus_b200.infer.dat_writer as dw
    seq = iter([uids, unique_ids(len(gl))])
    orig = dw.unique_ids
    dw.unique_ids = lambda n: next(seq) if n else []
    try:
        write_dat(info, path)
    finally:
        dw.unique_ids = orig
    for loader in (lambda p: pickle.load(open(p, "rb")), joblib.load):
        got = loader(path)
        assert list(got.keys()) == list(want.keys())
        _assert_same(got["Nuclei"], want["Nuclei"])
        assert len(got["Lumen"]) == len(gl) and got["empty"] == {}
        lum = next(iter(got["Lumen"].values()))
        assert lum["prob"] is None and lum["type"] is None
        assert got["proc_resolution"] == want["proc_resolution"]
        assert np.array_equal(got["proc_dimensions"], want["proc_dimensions"])
        assert got["Gland"]["a"]["type_prob"] == 0.5
    # arrays read back own their data and are writable, like any unpickled ndarray
    a = got["Nuclei"][uids[0]]["contour"]
    assert a.flags.writeable and a.flags.c_contiguous


def test_store_transport_and_removal():
    rng = np.random.RandomState(1)
    s = _random_store(rng, 900)
    dead = rng.choice(len(s), 100, replace=False)
    s.remove(dead)
    t = InstanceStore.unpack(pickle.loads(pickle.dumps(s.pack())))
    uids = unique_ids(len(t))
    _assert_same(s.to_dict(uids), t.to_dict(uids))
    assert len(t) == len(s) - 100
