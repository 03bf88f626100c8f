"""Parity of the tensor path (tcgen05 FP16 filter + exact rerank, csrc/tc_filter.cuh) against the
oracle: the filter may only ever ADD candidates, the rerank is the exact fold, so indices must be
identical (ties by index) and distances bit-identical, exactly as for the SIMT path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pn():
    import petal_neighbors_b200 as pn
    return pn


def bits(a):
    return a.view(np.uint32)


def check(pn, oracle, pts, Q, k, expect_tensor=True, **opts):
    bt = pn.BallTree.euclidean(pts, algo=pn.PN_ALGO_TENSOR, **opts)
    idx, dist = bt.query_batch(Q, k)
    oi, od = oracle.brute_knn(pts, Q, k)
    bad = np.argwhere(idx != oi.astype(np.uint64))
    assert bad.size == 0, f"index mismatch at {bad[:5]} got {idx[tuple(bad[0])]} want {oi[tuple(bad[0])]}"
    assert np.array_equal(bits(dist), bits(od)), "distances are not bit-identical"
    c = bt.counters()
    if pts.shape[1] + 6 <= 384 and expect_tensor:
        # every pair, unless the build-time estimate turned pruning on for this data (clusters): then at most every pair
        assert 0 < c["filter_pairs"] <= pts.shape[0] * Q.shape[0] * max(1, -(-k // 16))
    else:
        assert c["filter_pairs"] == 0
    return bt, c


@pytest.mark.parametrize("n,d,nq,k", [
    (1000, 16, 300, 10),       # DVR=4, MT=2, one K chunk
    (128, 16, 256, 10),        # exactly one tile
    (5000, 16, 1000, 1),       # K=1 kernel
    (4097, 20, 513, 16),       # DVR=8
    (3000, 32, 257, 10),       # d = 32 -> Kp = 64, two K chunks
    (6000, 64, 300, 10),       # generic query path, three K chunks, MT=2
    (4000, 100, 200, 10),      # Kp = 128 -> MT=1
    (3000, 128, 130, 10),      # Kp = 160, five K chunks
    (2000, 3, 100, 10),        # forced tensor at tiny d
    (3000, 16, 200, 40),       # multi-pass k > 16
    (50, 16, 10, 10),          # n < one tile
    (7, 16, 3, 10),            # k > n: padded rows
    (20000, 16, 5000, 10),
    (1500, 300, 300, 10),      # Kp = 320: ten K chunks, MT=1, two requests per tile
    (1200, 200, 260, 10),      # Kp = 224: seven chunks -> second request reads one chunk past the tensor (zero fill)
    (2500, 170, 300, 10),      # Kp = 192: six chunks, one request per tile, MT=1
    (800, 500, 200, 10),       # too wide for the resident A operand: falls back to the exact wide scan
])
def test_tensor_knn_random(pn, oracle, n, d, nq, k):
    from petal_neighbors_b200 import synth
    pts = synth.uniform(n, d, 51 + n + d, np.float32)
    Q = synth.uniform(nq, d, 52 + n + d, np.float32)
    check(pn, oracle, pts, Q, k)


def test_tensor_ties_and_duplicates(pn, oracle):
    rng = np.random.default_rng(17)
    pts = rng.integers(0, 4, size=(5000, 16)).astype(np.float32)
    Q = rng.integers(0, 4, size=(300, 16)).astype(np.float32)
    for k in (1, 10, 16):
        check(pn, oracle, pts, Q, k)
    pts = np.ones((1000, 16), np.float32)
    check(pn, oracle, pts, pts[:50] + np.float32(0.5), 10)
    check(pn, oracle, pts, pts[:50], 10)        # all distances exactly zero


def test_tensor_clustered_offset_and_far_queries(pn, oracle):
    """Data far from the origin (centring matters), tight clusters (tiny distances vs large norms),
    and queries far outside the data (large |q'|: the margin grows, the answer must not change)."""
    from petal_neighbors_b200 import synth
    pts = synth.gaussian_mixture(8000, 32, 5, n_centers=16, sigma=0.01, dtype=np.float32) + np.float32(100.0)
    Q = synth.gaussian_mixture(400, 32, 6, n_centers=16, sigma=0.01, dtype=np.float32) + np.float32(100.0)
    check(pn, oracle, pts, Q, 10)
    Qfar = Q * np.float32(3.0)
    check(pn, oracle, pts, Qfar, 5)
    pts2 = synth.uniform(4000, 16, 1, np.float32) * np.float32(1e-3) + np.float32(7.0)
    check(pn, oracle, pts2, pts2[:300], 10)     # self-queries: first neighbour at distance 0


def test_tensor_auto_selection(pn, oracle):
    from petal_neighbors_b200 import synth
    pts = synth.uniform(30000, 16, 2, np.float32)
    Q = synth.uniform(4096, 16, 3, np.float32)
    bt = pn.BallTree.euclidean(pts)             # AUTO: f32, d >= 16 -> tensor at every batch size
    idx, dist = bt.query_batch(Q, 10)
    assert bt.counters()["filter_pairs"] == 30000 * 4096
    oi, od = oracle.brute_knn(pts, Q, 10)
    assert np.array_equal(idx, oi.astype(np.uint64)) and np.array_equal(bits(dist), bits(od))
    idx, dist = bt.query_batch(Q[:100], 10)     # the point stream is split over the SMs for small batches
    assert bt.counters()["filter_pairs"] == 30000 * 100
    assert np.array_equal(idx, oi[:100].astype(np.uint64)) and np.array_equal(bits(dist), bits(od[:100]))
    idx, dist = bt.query_batch(Q[:1], 10)       # a single query too
    assert bt.counters()["filter_pairs"] == 30000
    assert np.array_equal(idx, oi[:1].astype(np.uint64)) and np.array_equal(bits(dist), bits(od[:1]))
    i1, d1 = bt.query(Q[0], 10)                 # the reference's own call shape
    assert np.array_equal(i1, oi[0].astype(i1.dtype)) and np.array_equal(bits(d1), bits(od[0]))
    bs = pn.BallTree.euclidean(pts, algo=pn.PN_ALGO_SIMT)
    idx2, dist2 = bs.query_batch(Q, 10)
    assert bs.counters()["filter_pairs"] == 0 and np.array_equal(idx2, oi.astype(np.uint64))


def test_tensor_vp_tree(pn, oracle):
    from petal_neighbors_b200 import synth
    pts = synth.gaussian_mixture(20000, 64, 5, n_centers=32, dtype=np.float32)
    Q = synth.gaussian_mixture(3000, 64, 6, n_centers=32, dtype=np.float32)
    vp = pn.VantagePointTree.euclidean(pts, algo=pn.PN_ALGO_TENSOR)
    vi, vd = vp.query_nearest_batch(Q)
    oi, od = oracle.brute_knn(pts, Q, 1)
    assert np.array_equal(vi, oi[:, 0].astype(np.uint64)) and np.array_equal(bits(vd), bits(od[:, 0]))


def test_tensor_adversarial_scales(pn, oracle):
    """Inputs chosen to stress the filter's rounding bound; the answer must stay bit-exact.
    (a) one extreme outlier: the scale factor is set by it, every other scaled coordinate lands in fp16's
        subnormal range (the absolute 2^-14 term of E_q has to carry the filter);
    (b) queries far outside the fp16-safe range (scaled norm > 200): E_q = +inf, everything is reranked;
    (c) tiny coordinates (1e-20 scale) and huge ones (1e15 scale);
    (d) near-ties at the k-th boundary: distances that differ in the last few ulps."""
    from petal_neighbors_b200 import synth
    rng = np.random.default_rng(23)
    base = synth.uniform(3000, 16, 7, np.float32)
    Q = synth.uniform(2100, 16, 8, np.float32)
    a = base.copy()
    a[1234] = np.float32(1.0e6)                                    # (a)
    check(pn, oracle, a, Q, 10)
    check(pn, oracle, base, Q * np.float32(5000.0), 5)             # (b)
    # (c) at 1e-20 the square of the power-of-two scale leaves the float range: the tree stays on the exact scan
    check(pn, oracle, base * np.float32(1e-20), Q * np.float32(1e-20), 10, expect_tensor=False)
    check(pn, oracle, base * np.float32(1e-12), Q * np.float32(1e-12), 10)
    check(pn, oracle, base * np.float32(1e15), Q * np.float32(1e15), 10)
    ring = rng.standard_normal((4000, 16)).astype(np.float32)      # (d) points on a thin shell around the origin
    ring /= np.linalg.norm(ring, axis=1, keepdims=True)
    ring *= (1.0 + 1e-6 * rng.standard_normal((4000, 1))).astype(np.float32)
    Qc = (1e-3 * rng.standard_normal((2100, 16))).astype(np.float32)  # queries near the centre: all distances ~ 1
    check(pn, oracle, ring, Qc, 10)
    check(pn, oracle, ring, Qc, 1)


@pytest.mark.parametrize("n,d,nq,k", [
    (40000, 16, 700, 10),      # 313 tiles, 2 query tiles of 512: the point stream is split 4 ways, shared k-th bounds
    (40000, 16, 700, 40),      # the same with three passes (floors) over the split stream
    (40000, 16, 3, 1),         # a tiny batch, K = 1
    (36000, 64, 600, 10),      # generic row width, four subtiles, split stream
    (33000, 128, 300, 10),     # two subtiles x two stages, split stream
    (70000, 16, 76000, 10),    # one whole wave (148 x 512 queries) + a one-tile tail launch split over the stream
])
def test_tensor_split_point_stream(pn, oracle, n, d, nq, k):
    """Batches smaller than a wave (and the last partial wave of larger ones) split the point stream over grid.y;
    the per-split lists are merged and the splits share their k-th bounds -- results must not change."""
    from petal_neighbors_b200 import synth
    pts = synth.uniform(n, d, 61 + n + d, np.float32)
    pts[n // 2] = pts[n // 3]                      # an exact duplicate across two different splits
    Q = synth.uniform(nq, d, 62 + n + d, np.float32)
    Q[0] = pts[n // 3]                             # ... queried at distance 0: tie broken by index
    check(pn, oracle, pts, Q, k)


def test_tensor_self_query_split(pn, oracle):
    """Self query on a tensor-eligible tree: the row map of the merge kernel and the tail launch together."""
    from petal_neighbors_b200 import synth
    n, d, k = 30000, 16, 10
    pts = synth.uniform(n, d, 77, np.float32)
    bt = pn.BallTree.euclidean(pts, algo=pn.PN_ALGO_TENSOR)
    idx, dist = bt.query_self(k)
    oi, od = oracle.brute_knn(pts, pts, k)
    assert np.array_equal(idx, oi.astype(np.uint64)) and np.array_equal(bits(dist), bits(od))
    assert bt.counters()["h2d_bytes"] == 0
