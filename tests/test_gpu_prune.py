"""Triangle-inequality pruning in front of the tensor filter (csrc/tc_prune.cuh; reference src/ball_tree.rs:211-214,
230-238): skipped (query group, point tile) pairs must never change the answer -- indices identical, distances
bit-identical to the oracle -- and on clustered data most pairs must really be skipped."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pn():
    import petal_neighbors_b200 as pn
    return pn


def bits(a):
    return a.view(np.uint32)


def check(pn, oracle, tree, pts, Q, k, vp=False):
    if vp:
        vi, vd = tree.query_nearest_batch(Q)
        idx, dist = vi[:, None], vd[:, None]
    else:
        idx, dist = tree.query_batch(Q, k)
    oi, od = oracle.brute_knn(pts, Q, k)
    bad = np.argwhere(idx != oi.astype(np.uint64))
    assert bad.size == 0, f"index mismatch at {bad[:5]}"
    assert np.array_equal(bits(dist), bits(od))
    return tree.counters()


@pytest.mark.parametrize("n,d,nq,k,centers,sigma,strong", [
    (60000, 64, 3000, 1, 64, 0.05, False), (60000, 64, 3000, 10, 64, 0.05, False), (50000, 16, 20000, 10, 64, 0.004, True),
    (40000, 32, 700, 16, 32, 0.03, False), (30000, 128, 1500, 10, 16, 0.05, False), (20000, 20, 1, 5, 8, 0.02, False),
])
def test_pruned_scan_on_clusters(pn, oracle, n, d, nq, k, centers, sigma, strong):
    """Exactness on clustered data with the pruned scan forced on and in AUTO; on tight, well separated clusters (the
    `strong` case) most (query group, tile) pairs must really be skipped.  (In d >= 64 the median-split partition mixes
    neighbouring clusters in one bucket, and tile balls prune little: see DESIGN.md.)"""
    from petal_neighbors_b200 import synth
    pts = synth.gaussian_mixture(n, d, 5, n_centers=centers, sigma=sigma)
    Q = synth.gaussian_mixture(nq, d, 6, n_centers=centers, sigma=sigma)
    for prune in (pn.PN_PRUNE_ON, pn.PN_PRUNE_AUTO):
        bt = pn.BallTree.euclidean(pts, prune=prune)
        c = check(pn, oracle, bt, pts, Q, k)
        assert 0 < c["filter_pairs"] <= n * nq and c["pairs"] <= n * nq
        if strong:
            assert c["pairs"] < 0.5 * n * nq, f"pruned only to {c['pairs'] / (n * nq):.3f} of the pairs (prune={prune})"
    off = pn.BallTree.euclidean(pts, prune=pn.PN_PRUNE_OFF)
    c = check(pn, oracle, off, pts, Q, k)
    assert c["pairs"] == n * nq
    vp = pn.VantagePointTree.euclidean(pts, prune=pn.PN_PRUNE_ON)     # tensor-path queries of a VP handle use the ball partition
    c = check(pn, oracle, vp, pts, Q, 1, vp=True)
    assert 0 < c["pairs"] <= n * nq
    if strong:
        assert c["pairs"] < 0.5 * n * nq
    vs = pn.VantagePointTree.euclidean(pts, algo=pn.PN_ALGO_SIMT)     # the VP traversal proper
    check(pn, oracle, vs, pts, Q, 1, vp=True)


@pytest.mark.parametrize("n,d,nq,k", [(30000, 16, 2000, 10), (100, 16, 50, 10), (5000, 40, 513, 1), (9000, 24, 300, 17), (700, 16, 3, 3)])
def test_forced_pruning_on_uniform_data_stays_exact(pn, oracle, n, d, nq, k):
    """Uniform data prunes (almost) nothing: forcing the pruned path must still give the exact answer; k > 16 runs the
    multi-pass dense scan whatever the option says."""
    from petal_neighbors_b200 import synth
    pts = synth.uniform(n, d, 31 + n, np.float32)
    Q = synth.uniform(nq, d, 32 + n, np.float32)
    bt = pn.BallTree.euclidean(pts, prune=pn.PN_PRUNE_ON, algo=pn.PN_ALGO_TENSOR, bucket_size=64)
    c = check(pn, oracle, bt, pts, Q, k)
    assert 0 < c["pairs"] <= n * nq * max(1, -(-k // 16))
    auto = pn.BallTree.euclidean(pts, algo=pn.PN_ALGO_TENSOR)   # AUTO: the estimate keeps pruning off here
    c = check(pn, oracle, auto, pts, Q, k)
    assert c["pairs"] == n * nq * max(1, -(-k // 16))


def test_pruning_with_ties_far_queries_and_self_query(pn, oracle):
    rng = np.random.default_rng(3)
    # clusters of exactly identical points (tile radius 0) plus lattice noise: masses of ties, also at the k-th boundary
    base = rng.integers(0, 4, size=(40, 16)).astype(np.float32) * 10
    pts = np.repeat(base, 500, axis=0) + rng.integers(0, 2, size=(20000, 16)).astype(np.float32)
    rng.shuffle(pts)
    Q = np.concatenate([pts[:600] + 0.5, rng.random((200, 16), np.float32) * 1000.0 - 300.0]).astype(np.float32)   # near and very far queries
    bt = pn.BallTree.euclidean(pts, prune=pn.PN_PRUNE_ON, algo=pn.PN_ALGO_TENSOR, bucket_size=128)
    for k in (1, 10, 16):
        check(pn, oracle, bt, pts, Q, k)
    idx, dist = bt.query_self(8)
    oi, od = oracle.brute_knn(pts, pts, 8)
    assert np.array_equal(idx, oi.astype(np.uint64)) and np.array_equal(bits(dist), bits(od))
    c = bt.counters()
    assert c["pairs"] < len(pts) ** 2            # some (point tile, query group) pairs were skipped


def test_c3_full_size_seeded(pn, oracle):
    """BASELINE config 3 at full size in AUTO.  The build-time estimates turn the SEEDED scan on for this mixture (every query
    starts from the k-th distance within its home bucket: exact reranks fall from ~290 to ~10 per query at k = 1).  With the
    reference partition tile skipping would stay off -- in d = 64 its buckets mix neighbouring clusters and only ~10 % of the
    pairs could be skipped -- so the handle builds the two-means partition of the same points, whose estimates turn the tile
    bitmaps on: well under 60 % of the pairs are scanned.  Exact on samples, VP handle and ball handle."""
    from petal_neighbors_b200 import synth
    n = nq = 1_000_000
    pts = synth.fast_gaussian_mixture(n, 64, 5, n_centers=1024, sigma=0.05, center_seed=4)
    Q = synth.fast_gaussian_mixture(nq, 64, 6, n_centers=1024, sigma=0.05, center_seed=4)
    vp = pn.VantagePointTree.euclidean(pts)
    inf = vp.info()
    assert inf["tensor_partition"] == 1 and inf["prune_seeded"] == 1 and inf["prune_tiles"] == 1, inf
    vi, vd = vp.query_nearest_batch(Q)
    c = vp.counters()
    assert 0 < c["pairs"] < 0.6 * float(n) * nq, c["pairs"] / (float(n) * nq)
    assert c["rerank_pairs"] / nq < 100, c["rerank_pairs"] / nq
    sample = np.arange(0, nq, nq // 400)[:400]
    oi, od = oracle.brute_knn(pts, Q[sample], 1)
    assert np.array_equal(vi[sample], oi[:, 0].astype(np.uint64)) and np.array_equal(bits(vd[sample]), bits(od[:, 0]))
    bt = pn.BallTree.euclidean(pts, prune=pn.PN_PRUNE_ON)              # tiles forced on: some pairs are skipped, same answer
    idx, dist = bt.query_batch(Q[:300_000], 10)
    c = bt.counters()
    assert c["pairs"] < float(n) * 300_000
    s2 = sample[sample < 300_000]
    oi, od = oracle.brute_knn(pts, Q[s2], 10)
    assert np.array_equal(idx[s2], oi.astype(np.uint64)) and np.array_equal(bits(dist[s2]), bits(od))
    auto = pn.BallTree.euclidean(pts)                                  # a ball handle in AUTO keeps the two-means partition as well
    assert auto.info()["tensor_partition"] == 1
    idx, dist = auto.query_batch(Q[:300_000], 10)
    assert auto.counters()["pairs"] < 0.95 * float(n) * 300_000    # (73 queries per bucket: a CTA's 512 queries span seven buckets)
    assert np.array_equal(idx[s2], oi.astype(np.uint64)) and np.array_equal(bits(dist[s2]), bits(od))
