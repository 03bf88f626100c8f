"""CPU-only checks (no GPU): the C-ABI library loads and exports every symbol include/petal_b200.h
declares, the host builder's flattened layout satisfies the invariants the kernels rely on, and
the host mirror keeps the reference's error behaviour.  No compute call is made."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pn():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "petal-neighbors_b200", "lib", "libpetal_b200.so")):
        g.build()
    import petal_neighbors_b200 as pn
    return pn


def test_header_symbols_exported(pn):
    from petal_neighbors_b200 import _ffi
    hdr = open(os.path.join(ROOT, "include", "petal_b200.h")).read()
    declared = set(re.findall(r"^(?:const char \*|int32_t |void )\s*(pn_[a-z0-9_]+)\s*\(", hdr, re.M))
    assert declared == set(_ffi.EXPORTS), declared ^ set(_ffi.EXPORTS)
    L = _ffi.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.pn_abi_version() == 1


def test_no_cpu_fallback(pn):
    pts = np.random.default_rng(0).random((100, 3)).astype(np.float32)
    t = pn.BallTree.euclidean(pts, host_only=True)
    with pytest.raises(pn.EngineError) as e:
        t.query(pts[0], 3)
    assert e.value.status == 4 and "no CPU fallback" in str(e.value)
    with pytest.raises(pn.EngineError):
        t.query_radius(pts[0], 0.1)


def test_array_errors(pn):
    with pytest.raises(pn.ArrayError) as e:
        pn.BallTree.euclidean(np.empty((0, 0)), host_only=True)   # src/ball_tree.rs:623-630
    assert e.value.kind == "Empty" and str(e.value) == "array is empty"
    arr = np.array([[1., 1.], [1., 1.1], [9., 9.]])
    with pytest.raises(pn.ArrayError) as e:
        pn.BallTree.euclidean(arr.T, host_only=True)              # src/ball_tree.rs:632-638
    assert e.value.kind == "NotContiguous" and str(e.value) == "array is not contiguous in memory"
    with pytest.raises(pn.ArrayError):
        pn.VantagePointTree.euclidean(np.empty((0, 2), np.float32), host_only=True)
    with pytest.raises(TypeError):
        pn.BallTree.new(arr, object(), host_only=True)
    assert pn.BallTree.new(arr, pn.distance.Euclidean(), host_only=True).metric == pn.BallTree.euclidean(arr, host_only=True).metric  # :640-647


def test_metric_types(pn, oracle):
    m = pn.distance.Euclidean()
    x, y = np.array([3., 4.]), np.array([0., 0.])
    assert m.distance(x, y) == 5.0 and m.rdistance(x, y) == 25.0           # src/distance.rs:130-134
    assert m.rdistance_to_distance(25.0) == 5.0 and m.distance_to_rdistance(5.0) == 25.0
    rng = np.random.default_rng(1)
    for dt in (np.float32, np.float64):
        a, b = rng.random(19).astype(dt), rng.random(19).astype(dt)
        assert m.distance(a, b) == oracle.distance(a, b)


def fold_dist(a, b):
    """sequential fold in the array's precision"""
    t = a.dtype.type
    s = np.zeros(a.shape[0], a.dtype)
    for j in range(a.shape[1]):
        diff = a[:, j] - b[j]
        s = s + diff * diff
    return np.sqrt(s)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,d,bucket", [(1, 2, 0), (9, 3, 8), (1000, 3, 8), (5000, 16, 64), (4097, 5, 256), (3000, 130, 100)])
def test_ball_layout_invariants(pn, dtype, n, d, bucket):
    pts = np.random.default_rng(n + d).random((n, d)).astype(dtype)
    t = pn.BallTree.euclidean(pts, host_only=True, bucket_size=bucket)
    lay = t.layout()
    L, nb, nn, dp = lay["n_levels"], lay["n_buckets"], lay["n_nodes"], lay["dim_padded"]
    assert nb == 1 << L and nn == (1 << (L + 1)) - 1 and dp % (16 // pts.itemsize) == 0 and dp >= d
    ids = lay["ids"]
    assert sorted(ids.tolist()) == list(range(n))                        # a permutation
    assert np.array_equal(lay["points"][:, :d], pts[ids]) and np.all(lay["points"][:, d:] == 0)
    lo, hi = lay["bucket_lo"], lay["bucket_hi"]
    assert lo[0] == 0 and hi[-1] == n and np.array_equal(lo[1:], hi[:-1])  # contiguous cover
    assert (hi - lo).max() <= max(bucket or 256, 8) and (hi - lo).max() - (hi - lo).min() <= 1
    # every node's ball contains all of its points: radius = max fold distance to the centroid
    for node in range(nn):
        first = last = node
        while first < nb - 1:
            first, last = 2 * first + 1, 2 * last + 2
        a, b = lo[first - (nb - 1)], hi[last - (nb - 1)]
        c = lay["node_center"][node]
        dist = fold_dist(lay["points"][a:b], c)
        assert dist.max() == lay["node_radius"][node]
        np.testing.assert_allclose(c[:d], pts[ids[a:b]].mean(axis=0), rtol=1e-4 if dtype == np.float32 else 1e-12)
    # median split: children sizes follow mid = (start + end) / 2 (src/ball_tree.rs:535)
    assert hi[nb // 2 - 1] == n // 2 if nb > 1 else True


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,d,bucket", [(1, 2, 0), (50, 3, 8), (3000, 4, 16), (5000, 64, 128)])
def test_vp_layout_invariants(pn, dtype, n, d, bucket):
    pts = np.random.default_rng(n * 3 + d).random((n, d)).astype(dtype)
    t = pn.VantagePointTree.euclidean(pts, host_only=True, bucket_size=bucket)
    lay = t.layout()
    L, nb, nn = lay["n_levels"], lay["n_buckets"], lay["n_nodes"]
    assert nb == 1 << L and nn == nb - 1
    ids = lay["ids"]
    assert sorted(ids.tolist()) == list(range(n))
    lo, hi = lay["bucket_lo"], lay["bucket_hi"]
    P = lay["points"]

    def span(node):  # stored rows of the subtree below `node`, excluding vantage points above it
        first = last = node
        while first < nb - 1:
            first, last = 2 * first + 1, 2 * last + 2
        return int(lo[first - (nb - 1)]), int(hi[last - (nb - 1)])

    covered = int((hi - lo).sum()) + nn
    assert covered == n                                                    # buckets + one vantage point per node
    for node in range(nn):
        vp, mu = lay["node_center"][node], lay["node_radius"][node]
        a, b = span(2 * node + 1)
        c, e = span(2 * node + 2)
        level = int(np.log2(node + 1))
        # slice order [near | far | vp]: the far child's own chain of vantage points ends its slice
        assert np.array_equal(P[e + (L - level - 1)][:d], vp[:d])
        if b > a:
            assert fold_dist(P[a:b], vp).max() <= mu                       # near: d(p, vp) <= mu
        if e > c:
            assert fold_dist(P[c:e], vp).min() >= mu                       # far:  d(p, vp) >= mu
        assert abs((b - a) - (e - c)) <= 2 ** (L + 1)


def test_shard_ranges_partition(pn):
    pts = np.random.default_rng(5).random((10001, 6)).astype(np.float32)
    seen = []
    for s in range(8):
        t = pn.BallTree.euclidean(pts, host_only=True, shard_depth=3, shard_index=s, bucket_size=64)
        inf = t.info()
        assert inf["n_points_total"] == 10001
        seen.append(t.layout()["ids"])
    sizes = [len(x) for x in seen]
    assert sum(sizes) == 10001 and max(sizes) - min(sizes) <= 1
    assert sorted(np.concatenate(seen).tolist()) == list(range(10001))
    with pytest.raises(pn.EngineError):
        pn.BallTree.euclidean(pts, host_only=True, shard_depth=3, shard_index=8)


def test_synth_is_counter_based():
    import petal_neighbors_b200  # noqa: F401
    from petal_neighbors_b200 import synth
    a = synth.uniform(100, 7, 3, np.float32)
    assert np.array_equal(a[40:60], synth.uniform(20, 7, 3, np.float32, row0=40))
    assert a.min() >= 0 and a.max() < 1
    g = synth.gaussian_mixture(64, 5, 9, n_centers=4, dtype=np.float64)
    assert np.array_equal(g[10:20], synth.gaussian_mixture(10, 5, 9, n_centers=4, dtype=np.float64, row0=10))


def test_shards_of_fewer_points_than_shards(pn):
    """n < 2^shard_depth: every point is owned by exactly one shard, the other shards are Empty (the descent
    used to stop at one-point ranges, so several shard indices returned the same point)."""
    for n, depth in ((3, 2), (5, 3), (1, 2)):
        pts = np.random.default_rng(n).random((n, 4)).astype(np.float32)
        ids, empty = [], 0
        for s in range(1 << depth):
            try:
                t = pn.BallTree.euclidean(pts, host_only=True, shard_depth=depth, shard_index=s)
                ids += t.layout()["ids"].tolist()
            except pn.ArrayError as e:
                assert e.kind == "Empty"
                empty += 1
        assert sorted(ids) == list(range(n)) and empty == (1 << depth) - n


def test_torch_generators_match_numpy():
    import torch
    import petal_neighbors_b200  # noqa: F401
    from petal_neighbors_b200 import synth
    for dt, tdt in ((np.float32, torch.float32), (np.float64, torch.float64)):
        a = synth.uniform(1000, 7, 3, dt, row0=5)
        assert np.array_equal(a, synth.uniform_torch(1000, 7, 3, tdt, row0=5, device="cpu").numpy())
        kw = dict(n_centers=37, sigma=0.05, row0=11, clip=True)
        g = synth.gaussian_mixture(3000, 9, 5, dtype=dt, **kw)
        assert np.array_equal(g, synth.gaussian_mixture_torch(3000, 9, 5, dtype=tdt, device="cpu", **kw).numpy())
