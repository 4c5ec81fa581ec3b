"""Host logic of the WSI path (cerberus_b200/infer/wsi_geometry.py, wsi_reader.py): the restated
tiatoolbox / shapely placement rules against hand-derived cases (the originals cannot be imported
offline - parity unpinned)."""
import numpy as np
import pytest

from cerberus_b200.infer.wsi_geometry import (boxes_intersect, boxes_within, filter_coordinates,
                                              get_coordinates, get_tile_info, select_tile_instances)
from cerberus_b200.infer.wsi_reader import ArraySlide


def test_get_coordinates_grid_and_input_offset():
    pin, pout = get_coordinates((300, 200), [448, 448], [144, 144], [144, 144])
    # ceil(300/144)*144 = 432 -> x = 0,144,288 ; ceil(200/144)*144 = 288 -> y = 0,144 ; x fastest
    assert pout[:, :2].tolist() == [[0, 0], [144, 0], [288, 0], [0, 144], [144, 144], [288, 144]]
    assert np.array_equal(pout[:, 2:] - pout[:, :2], np.full((6, 2), 144))
    assert np.array_equal(pin[:, :2], pout[:, :2] - 152)  # (448 - 144) // 2
    assert np.array_equal(pin[:, 2:] - pin[:, :2], np.full((6, 2), 448))
    # 20000^2 slide -> 139 x 139 candidate patches
    _, po = get_coordinates((20000, 20000), [448, 448], [144, 144], [144, 144])
    assert len(po) == 139 * 139


def test_filter_coordinates_scales_boxes_into_the_mask():
    mask = np.zeros((100, 150), dtype=np.uint8)  # slide 200 x 300 (rows x cols) at scale 0.5
    mask[10:20, 100:110] = 1
    _, pout = get_coordinates((300, 200), [448, 448], [144, 144], [144, 144])
    keep = filter_coordinates(mask, pout, (200, 300))
    # mask blob = slide rows 20..40, cols 200..220 -> only the output box x 144..288, y 0..144
    assert keep.tolist() == [False, True, False, False, False, False]
    assert filter_coordinates(np.ones((200, 300), np.uint8), pout, (200, 300)).all()


def test_tile_info_single_tile_slide():
    info = get_tile_info((400, 300), [4096, 4096], [144, 144], 64)
    assert len(info) == 1
    boxes, flags = info[0]
    assert boxes.tolist() == [[0, 0, 4032, 4032]]  # floor(4096/144)*144
    assert flags.tolist() == [[0, 0, 0, 0]]


def test_tile_info_grid_strips_and_crosses():
    info = get_tile_info((1000, 700), [450, 450], [144, 144], 64)  # tile -> 432
    assert len(info) == 4
    grid, gflags = info[0]
    assert grid.tolist() == [[0, 0, 432, 432], [432, 0, 864, 432], [864, 0, 1296, 432],
                             [0, 432, 432, 864], [432, 432, 864, 864], [864, 432, 1296, 864]]
    # [top, bottom, left, right]; sides on (or beyond) the slide border are never removed
    assert gflags.tolist() == [[0, 1, 0, 1], [0, 1, 1, 1], [0, 1, 1, 0],
                               [1, 0, 0, 1], [1, 0, 1, 1], [1, 0, 1, 0]]
    vert, vflags = info[1]  # over the vertical seams x = 432 and x = 864, one per tile row
    assert vert.tolist() == [[368, 0, 496, 432], [800, 0, 928, 432], [368, 432, 496, 864], [800, 432, 928, 864]]
    assert vflags.tolist() == [[0, 1, 0, 0], [0, 1, 0, 0], [1, 0, 0, 0], [1, 0, 0, 0]]
    hor, hflags = info[2]  # over the horizontal seam y = 432
    assert hor.tolist() == [[0, 368, 432, 496], [432, 368, 864, 496], [864, 368, 1296, 496]]
    assert hflags.tolist() == [[0, 0, 0, 1], [0, 0, 1, 1], [0, 0, 1, 0]]
    cross, cflags = info[3]
    assert cross.tolist() == [[304, 304, 560, 560], [736, 304, 992, 560]]
    assert not cflags.any()


def test_box_predicates_follow_shapely_semantics():
    boxes = np.array([[0, 0, 10, 10], [10, 0, 20, 10], [30, 30, 40, 40]])
    assert boxes_intersect(boxes, (10, 5, 12, 6)).tolist() == [0, 1]      # touching counts
    assert boxes_within(boxes, (0, 0, 20, 10)).tolist() == [0, 1]         # closed containment
    assert boxes_within(boxes, (1, 0, 20, 10)).tolist() == [1]


def test_select_tile_instances_rules():
    m, tb = 64, [432, 0, 864, 432]
    inst = np.array([
        [100, 380, 120, 400],   # 0: entirely inside the bottom band (y >= 368)
        [100, 360, 120, 400],   # 1: straddles the bottom band edge
        [5, 100, 30, 120],      # 2: entirely inside the left band
        [400, 100, 431, 130],   # 3: inside the right band
        [200, 200, 220, 220],   # 4: interior
        [0, 10, 3, 20],         # 5: touches the left boundary line
    ])
    # grid tile, flags bottom + left + right: only instances fully inside a flagged band go
    sel, ref = select_tile_instances(inst, tb, [0, 1, 1, 1], 0, m)
    assert sorted(set(sel)) == [0, 2, 3, 5] and ref == []
    # vertical strip (flags top/bottom): everything touching the top/bottom bands, plus everything
    # touching the 1-pixel left/right boundary lines
    sel, _ = select_tile_instances(inst, [368, 0, 496, 432], [1, 1, 0, 0], 1, m)
    assert sorted(set(sel)) == [0, 1, 5]
    # cross tile: all four bands (containment) + removal of accumulated instances on its inner frame
    ref_boxes = np.array([[432 + 60, 100, 432 + 70, 110],     # crosses the left margin line x = 64
                          [432 + 200, 200, 432 + 210, 210]])  # interior
    sel, ref = select_tile_instances(inst, tb, [0, 0, 0, 0], 3, m, ref_boxes)
    assert sorted(set(sel)) == [0, 2, 3, 5]
    assert sorted(set(ref)) == [0]
    with pytest.raises(ValueError):
        select_tile_instances(inst, tb, [0, 0, 0, 0], 7, m)


def test_array_slide_reads_are_zero_padded(tmp_path):
    img = np.random.RandomState(0).randint(1, 256, (40, 60, 3)).astype(np.uint8)
    p = str(tmp_path / "s.npy")
    np.save(p, img)
    s = ArraySlide.open(p, 0.5)
    assert s.slide_dimensions(0.5).tolist() == [60, 40]
    r = s.read_bounds([-5, -3, 11, 9])
    assert r.shape == (12, 16, 3)
    assert not r[:3].any() and not r[:, :5].any()
    assert np.array_equal(r[3:, 5:], img[:9, :11])
    r = s.read_bounds([50, 30, 70, 50])
    assert np.array_equal(r[:10, :10], img[30:, 50:]) and not r[10:].any() and not r[:, 10:].any()
    with pytest.raises(NotImplementedError):
        s.slide_dimensions(0.25)
    with pytest.raises(NotImplementedError):
        ArraySlide.open(str(tmp_path / "x.svs"), 0.5)


def test_rank_strided_batches_cover_every_patch_once():
    n, B = 1037, 30
    for world in (1, 2, 8):
        seen = np.zeros(n, dtype=np.int32)
        for rank in range(world):
            for bi, start in enumerate(range(0, n, B)):
                if bi % world == rank:
                    seen[start:start + B] += 1
        assert (seen == 1).all()


def test_dat_writer_round_trips_like_a_plain_pickle(tmp_path):
    """dump_dat (fast ndarray reducer) -> joblib.load / pickle.load give the arrays a plain
    pickle.dump would: same values, dtypes, shapes, writable and owning their data."""
    import pickle

    import joblib
    from cerberus_b200.infer.wsi import dump_dat
    rng = np.random.RandomState(0)
    big = rng.randint(0, 1000, (50, 2)).astype(np.int64)
    obj = {"Nuclei": {"k%d" % i: {"box": np.arange(4) + i, "centroid": rng.rand(2),
                                  "contour": big[i:i + 7], "prob": 0.25 * i, "type": i}
                      for i in range(20)},
           "odd": {"strided": big[::3, 1], "fortran": np.asfortranarray(rng.rand(3, 4)),
                   "zero_d": np.array(3.5), "empty": np.zeros((0, 2), np.int32),
                   "bool": rng.rand(5) > 0.5, "f32": rng.rand(2, 3).astype(np.float32),
                   "obj": np.array([None, "a"], dtype=object),
                   "struct": np.zeros(2, dtype=[("a", "<i4"), ("b", "<f8")])},
           "proc_dimensions": np.array([700, 900]), "proc_resolution": {"resolution": 0.5, "units": "mpp"}}
    path = str(tmp_path / "x.dat")
    dump_dat(obj, path)

    def same(a, b):
        if isinstance(a, dict):
            assert list(a.keys()) == list(b.keys())
            for k in a:
                same(a[k], b[k])
        elif isinstance(a, np.ndarray):
            assert a.dtype == b.dtype and a.shape == b.shape
            assert (a == b).all() if a.dtype != object else list(a) == list(b)
            assert b.flags.writeable
        else:
            assert a == b and type(a) is type(b)

    ref = pickle.loads(pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL))
    for loaded in (joblib.load(path), pickle.load(open(path, "rb"))):
        same(obj, loaded)
        same(ref, loaded)
        loaded["Nuclei"]["k3"]["contour"] += 1  # writable, not a view of anything shared
        assert loaded["Nuclei"]["k3"]["contour"].flags.owndata


def test_row_bands_split_whole_patch_rows_balanced():
    """Multi-GPU WSI inference: contiguous bands of whole patch rows, balanced by patch count."""
    from cerberus_b200.infer.wsi import InferManager
    from cerberus_b200.infer.wsi_geometry import get_coordinates
    _, pout = get_coordinates((3000, 2500), [448, 448], [144, 144], [144, 144])
    rng = np.random.RandomState(0)
    pout = pout[np.sort(rng.choice(len(pout), int(0.6 * len(pout)), replace=False))]  # mask-filtered
    for world in (1, 2, 3, 8, 64):
        bands = InferManager._row_bands(pout, world)
        assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == len(pout)
        rows_seen = []
        for (lo, hi), nxt in zip(bands, bands[1:] + [None]):
            assert lo <= hi
            if nxt is not None:
                assert hi == nxt[0]
            rows_seen.append(set(pout[lo:hi, 1].tolist()))
        for a in range(world):
            for b in range(a + 1, world):
                assert not (rows_seen[a] & rows_seen[b])  # no patch row is shared between ranks
        if world <= 8:
            sizes = [hi - lo for lo, hi in bands]
            assert max(sizes) - min(sizes) <= 2 * 21  # within two patch rows (21 patches wide)
