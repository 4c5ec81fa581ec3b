"""The WSI instance tables as arrays, and their `.dat` file (infer/wsi.py:853).

The reference accumulates `wsi_inst_info[tissue][uuid] = {box, centroid, contour, prob, type}`
one Python dict per instance and writes the whole thing with joblib.dump. A 20000^2 slide holds
half a million nuclei: building and pickling 1.6 million small arrays is the serial tail that
capped the multi-GPU speed-up of round 1 (8.2 s of a 17.7 s slide, all on rank 0). Here the
instances stay in a few big arrays from the device tables (csrc/instinfo.cu) to the file:

  InstanceStore   append-only columns (box / centroid / contour / prob / type) + an alive mask
                  (the cross tiles of infer/wsi.py:196-268 remove earlier instances); per-rank
                  stores travel as arrays, not as pickled dicts
  write_dat       the pickle stream of the dict is produced directly from the columns by
                  cerb_pickle_instances (csrc/dat_writer.cu); pickle.load / joblib.load read the
                  same dict back as from a plain pickle.dump of `store.to_dict()`
"""
import ctypes
import io
import os
import pickle

import numpy as np

from .. import _lib

_KEYS = ("box", "centroid", "contour", "prob", "type")


class InstanceStore:
    """Instances of one tissue in tiatoolbox's get_instance_info layout (infer/wsi.py:150):
    box int64 [x0, y0, x1, y1], centroid float64 [x, y], contour int64 [k, 2], prob float | None,
    type int | None. Rows keep insertion order; `remove` clears rows by global index."""

    def __init__(self, has_type=True):
        self.has_type = has_type
        self._box, self._cen, self._off, self._xy, self._prob, self._type = [], [], [], [], [], []
        self._n = 0
        self._npts = 0
        self._dead = []

    def __len__(self):
        return self._n

    def append(self, box, centroid, contour_off, contour_xy, prob=None, type_=None):
        """One tile's rows. contour_off: int64 [m+1] offsets into contour_xy (relative to 0)."""
        m = len(box)
        if m == 0:
            return np.zeros(0, dtype=np.int64)
        self._box.append(np.ascontiguousarray(box, dtype=np.int64).reshape(m, 4))
        self._cen.append(np.ascontiguousarray(centroid, dtype=np.float64).reshape(m, 2))
        off = np.asarray(contour_off, dtype=np.int64)
        self._off.append(off[1:] - off[0] + self._npts)
        xy = np.ascontiguousarray(contour_xy, dtype=np.int64).reshape(-1, 2)
        self._xy.append(xy)
        if self.has_type:
            self._prob.append(np.ascontiguousarray(prob, dtype=np.float64).reshape(m))
            self._type.append(np.ascontiguousarray(type_, dtype=np.int64).reshape(m))
        first = self._n
        self._n += m
        self._npts += len(xy)
        return np.arange(first, first + m, dtype=np.int64)

    def remove(self, idx):
        if len(idx):
            self._dead.append(np.asarray(idx, dtype=np.int64))

    def boxes(self):
        """int64 [n, 4] of every row appended so far (dead rows included: indices are global)."""
        return np.concatenate(self._box) if self._box else np.zeros((0, 4), dtype=np.int64)

    def alive(self):
        a = np.ones(self._n, dtype=bool)
        for d in self._dead:
            a[d] = False
        return a

    def columns(self):
        """Compacted (alive rows only) columns: box, centroid, contour_off [m+1], contour_xy, prob, type."""
        if self._n == 0:
            z = np.zeros
            return (z((0, 4), np.int64), z((0, 2), np.float64), z(1, np.int64), z((0, 2), np.int64),
                    z(0, np.float64) if self.has_type else None, z(0, np.int64) if self.has_type else None)
        box, cen = np.concatenate(self._box), np.concatenate(self._cen)
        end = np.concatenate(self._off)
        xy = np.concatenate(self._xy)
        start = np.concatenate([[0], end[:-1]])
        prob = np.concatenate(self._prob) if self.has_type else None
        typ = np.concatenate(self._type) if self.has_type else None
        a = self.alive()
        if not a.all():
            keep = np.nonzero(a)[0]
            lens = (end - start)[keep]
            new_off = np.concatenate([[0], np.cumsum(lens)])
            # ragged gather of the surviving contours
            src = np.repeat(start[keep] - new_off[:-1], lens) + np.arange(int(new_off[-1]), dtype=np.int64)
            xy = xy[src]
            box, cen = box[keep], cen[keep]
            if self.has_type:
                prob, typ = prob[keep], typ[keep]
            off = new_off
        else:
            off = np.concatenate([[0], end])
        return box, cen, off.astype(np.int64), xy, prob, typ

    def to_dict(self, uids=None):
        """The reference's dict (one Python dict per instance): the slow, object-by-object form."""
        box, cen, off, xy, prob, typ = self.columns()
        n = len(box)
        uids = uids if uids is not None else unique_ids(n)
        out = {}
        for j in range(n):
            out[uids[j]] = {"box": box[j], "centroid": cen[j], "contour": xy[off[j]:off[j + 1]],
                            "prob": float(prob[j]) if self.has_type else None,
                            "type": int(typ[j]) if self.has_type else None}
        return out

    # ---- transport between ranks: a dict of arrays (cheap to pickle / send)
    def pack(self):
        box, cen, off, xy, prob, typ = self.columns()
        return {"has_type": self.has_type, "box": box, "cen": cen, "off": off, "xy": xy, "prob": prob, "type": typ}

    @classmethod
    def unpack(cls, d):
        s = cls(d["has_type"])
        s.append(d["box"], d["cen"], d["off"], d["xy"], d["prob"], d["type"])
        return s


def unique_ids(n):
    """n random 128-bit hex keys (the reference draws uuid.uuid4().hex per instance, infer/wsi.py:265)."""
    raw = os.urandom(16 * n).hex()
    return [raw[32 * i:32 * i + 32] for i in range(n)]


_unique_ids_default = unique_ids  # tests substitute unique_ids to pin the keys


def _preamble():
    """Defines memo slots 1..11 for the objects every record refers to, then pops them."""
    rec = np.empty(0).__reduce__()[0]
    out = io.BytesIO()

    def put(i):
        out.write(b"q" + bytes([i]) + b"0")  # BINPUT i, POP

    def glob(mod, name, i):
        out.write(b"c" + mod.encode() + b"\n" + name.encode() + b"\n")
        put(i)

    def ustr(s):
        b = s.encode()
        return b"X" + len(b).to_bytes(4, "little") + b

    def dtype(code, i):
        # numpy.dtype(code, False, True) + state (3, '<', None, None, None, -1, -1, 0)
        out.write(b"cnumpy\ndtype\n" + ustr(code) + b"\x89\x88\x87R(K\x03" + ustr("<") +
                  b"NNNJ\xff\xff\xff\xffJ\xff\xff\xff\xffK\x00tb")
        put(i)

    glob(rec.__module__, rec.__name__, 1)
    glob("numpy", "ndarray", 2)
    out.write(b"K\x00\x85")
    put(3)
    out.write(b"C\x01b")
    put(4)
    dtype("i8", 5)
    dtype("f8", 6)
    for k, key in enumerate(_KEYS):
        out.write(ustr(key))
        put(7 + k)
    return out.getvalue(), np.arange(1, 12, dtype=np.uint8)


_HEX = np.frombuffer(b"0123456789abcdef", dtype=np.uint8)


def unique_id_bytes(n):
    """The same keys as unique_ids(n), as one ASCII buffer of n x 32 characters."""
    raw = np.frombuffer(os.urandom(16 * n), dtype=np.uint8)
    out = np.empty((16 * n, 2), dtype=np.uint8)
    out[:, 0] = _HEX[raw >> 4]
    out[:, 1] = _HEX[raw & 15]
    return out.tobytes()


def pickle_store_items(store, uids=None):
    """[b"}(", uint8 array, b"u"]: one pickled dict (protocol-2 opcodes) as a list of buffers."""
    lib = _lib.load()
    box, cen, off, xy, prob, typ = store.columns()
    n = len(box)
    if n == 0:
        return [b"}"]
    if uids is None:
        uids = unique_ids(n) if unique_ids is not _unique_ids_default else None
    uid_buf = "".join(uids).encode("ascii") if uids is not None else unique_id_bytes(n)
    assert len(uid_buf) == 32 * n
    _, memo = _preamble()
    fn = lib.cerb_pickle_instances

    def p(a):
        return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None

    args = (uid_buf, p(box), p(cen), p(off), p(xy), p(prob), p(typ), n, p(memo))
    need = fn(*args, None, 0)
    if need < 0:
        raise RuntimeError("cerb_pickle_instances failed (%d)" % need)
    buf = np.empty(need, dtype=np.uint8)
    got = fn(*args, p(buf), need)
    assert got == need
    return [b"}(", buf, b"u"]


def write_dat(wsi_inst_info, path):
    """`wsi_inst_info`: dict whose values are InstanceStore objects (written by the C serialiser)
    or ordinary Python objects (pickled without memo so that they embed in the stream)."""
    with open(path, "wb") as fh:
        fh.write(b"\x80\x03")
        fh.write(_preamble()[0])
        fh.write(b"}(")
        for key, val in wsi_inst_info.items():
            kb = str(key).encode()
            fh.write(b"X" + len(kb).to_bytes(4, "little") + kb)
            if isinstance(val, InstanceStore):
                for part in pickle_store_items(val):
                    fh.write(memoryview(part))
            else:
                bio = io.BytesIO()
                pk = pickle.Pickler(bio, protocol=3)
                pk.fast = True  # no memo opcodes: the sub-stream embeds anywhere
                pk.dump(val)
                raw = bio.getvalue()
                assert raw[:2] == b"\x80\x03" and raw[-1:] == b"."
                fh.write(raw[2:-1])
        fh.write(b"u.")
