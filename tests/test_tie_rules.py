"""CPU: randomized check of the two arguments that let the device watershed settle marker ties
without the whole-image heap emulation (csrc/postproc.cu):

  * ws_tie_harmless - a tie the flood can prove harmless leaves the result independent of the
    order in which the tied entries surface;
  * k_wsg_certify  - if every order of the tied entries gives the same labels, that labelling
    is the reference's.

The kernels' control flow is restated here in Python (component-parallel flood with a private
queue per mask component, tie groups, the harmless test, the enumeration of orders) with a
RANDOM order among tied marker entries, and compared with the exact restatement of
skimage.segmentation.watershed (oracle/postproc_oracle.c, one global heap) on thousands of small
tie-rich images. Every component the rules accept must come out exactly as the oracle's."""
import heapq
import itertools

import numpy as np
from scipy import ndimage

from oracle import postproc_oracle as po

NB = ((-1, 0), (0, -1), (0, 1), (1, 0))  # -W, -1, +1, +W


class Tied(Exception):
    pass


def _harmless(val, a, a_mask, e, e_mask):
    """csrc/postproc.cu::ws_tie_harmless."""
    tie = val[e]
    a_push = [(a[0] + NB[i][0], a[1] + NB[i][1]) for i in range(4) if a_mask & (1 << i)]
    e_push = [(e[0] + NB[i][0], e[1] + NB[i][1]) for i in range(4) if e_mask & (1 << i)]
    for q in a_push:
        if abs(q[0] - e[0]) + abs(q[1] - e[1]) == 1:
            return False
    for r in e_push:
        if val[r] < tie:
            return False
        if any(val[q] == val[r] for q in a_push):
            return False
    return True


def _flood(val, mask, lab, seeds, multi, rng, ranks=None, stats=None):
    """One mask component. ranks=None: the regular pass (markers age 0, random order among equal
    keys, tie groups + harmless test, raises Tied); ranks given: the enumeration pass."""
    H, W = val.shape
    n = len(seeds)
    heap = []
    for j, p in enumerate(seeds):
        if ranks is None:
            heap.append((val[p], 0, rng.random(), p))
        else:
            heap.append((val[p], ranks[j], 0.0, p))
    heapq.heapify(heap)
    age = 1 if ranks is None else n
    last = None
    g = []  # members of the current tie group: (pixel, push mask)
    clean = False
    while heap:
        v, a, _, e = heapq.heappop(heap)
        is_marker = ranks is None and a == 0
        tie_now = False
        if ranks is None:
            if is_marker:
                if last is not None and v == last and multi:
                    if not clean or len(g) >= 3:
                        raise Tied()
                    tie_now = True
                else:
                    g = []
                    clean = True
                last = v
            else:
                clean = False
        e_mask = 0
        for i, (dy, dx) in enumerate(NB):
            q = (e[0] + dy, e[1] + dx)
            if 0 <= q[0] < H and 0 <= q[1] < W and mask[q] and lab[q] == 0:
                age += 1
                lab[q] = lab[e]
                e_mask |= 1 << i
                heapq.heappush(heap, (val[q], age, 0.0, q))
        if is_marker:
            if tie_now:
                for (a_pix, a_mask) in g[:2]:
                    if not _harmless(val, a_pix, a_mask, e, e_mask):
                        raise Tied()
                stats["harmless_ties"] += 1
            if len(g) < 2:
                g.append((e, e_mask))
            else:
                g.append(None)


def _device_rules(val, markers, mask, rng, stats):
    """Labels as the device path would produce them, or None where it would fall back."""
    H, W = val.shape
    lab = (markers * mask).astype(np.int64)
    comp, ncomp = ndimage.label(mask)
    out_ok = np.zeros(ncomp + 1, bool)
    for c in range(1, ncomp + 1):
        pix = [tuple(p) for p in np.argwhere(comp == c)]
        seeds = []
        for p in pix:
            if lab[p] == 0:
                continue
            for dy, dx in NB:
                q = (p[0] + dy, p[1] + dx)
                if 0 <= q[0] < H and 0 <= q[1] < W and mask[q] and lab[q] == 0:
                    seeds.append(p)
                    break
        if not seeds:
            out_ok[c] = True
            continue
        multi = len({int(lab[p]) for p in seeds}) > 1
        saved = {p: int(lab[p]) for p in pix}
        try:
            _flood(val, mask, lab, seeds, multi, rng, stats=stats)
            out_ok[c] = True
            stats["plain"] += 1
            continue
        except Tied:
            stats["tied"] += 1
        # k_wsg_certify: every combination of permutations of the tie groups
        seeds.sort(key=lambda p: (val[p], p[0] * W + p[1]))
        groups = [list(g) for _, g in itertools.groupby(range(len(seeds)), key=lambda j: val[seeds[j]])]
        groups = [g for g in groups if len(g) > 1]
        variants = 1
        for g in groups:
            for f in range(2, len(g) + 1):
                variants *= f
        if len(groups) > 4 or variants > 24:
            for p in pix:
                lab[p] = saved[p]
            continue
        results = []
        for perms in itertools.product(*[itertools.permutations(g) for g in groups]):
            ranks = list(range(len(seeds)))
            for g, perm in zip(groups, perms):
                for j, r in zip(g, perm):
                    ranks[j] = r
            for p in pix:
                lab[p] = saved[p]
            _flood(val, mask, lab, seeds, True, rng, ranks=ranks)
            results.append([int(lab[p]) for p in pix])
        if all(r == results[0] for r in results):
            out_ok[c] = True
            stats["certified"] += 1
        else:
            stats["order_dependent"] += 1
    return lab, comp, out_ok


def _case(rng):
    H, W = rng.randint(8, 18), rng.randint(8, 20)
    f = ndimage.gaussian_filter(rng.randn(H, W), rng.uniform(0.8, 2.0))
    f = (f - f.min()) / (f.max() - f.min() + 1e-9)
    levels = rng.choice([6, 10, 16, 64])
    q = np.round(f * levels) / levels
    # most pixels get a unique value, a random subset keeps the quantised one: small tie groups
    uniq = q + rng.permutation(H * W).reshape(H, W) * 1e-6
    keep = rng.rand(H, W) < rng.choice([0.05, 0.15, 0.4])
    inner = np.where(keep, q, uniq).astype(np.float32)
    t_hi = rng.uniform(0.5, 0.75)
    mask = inner > rng.uniform(0.2, 0.45)
    markers = ndimage.label(inner > t_hi)[0]
    return (-inner).astype(np.float64), markers, mask


def test_settled_ties_reproduce_the_global_heap_order():
    rng = np.random.RandomState(12345)
    stats = {"plain": 0, "harmless_ties": 0, "tied": 0, "certified": 0, "order_dependent": 0}
    checked = 0
    for _ in range(6000):
        val, markers, mask = _case(rng)
        if not (markers * mask).any():
            continue
        ref = po.watershed(val, markers, mask=mask)
        lab, comp, ok = _device_rules(val, markers, mask, rng, stats)
        for c in np.nonzero(ok)[0]:
            if c == 0:
                continue
            sel = comp == c
            assert np.array_equal(lab[sel], ref[sel]), (stats, "component", int(c))
            checked += 1
    # the generator must actually exercise the rules
    assert checked > 6000 and stats["tied"] > 300 and stats["certified"] > 100, stats
    assert stats["harmless_ties"] > 200, stats
    assert stats["order_dependent"] > 0, stats
    print(stats, checked)
