"""Parity tests proper: the CUDA path, called through the C ABI (via the ctypes mirror of the
reference API), against the CPU oracle on the same seeded inputs.  Bar: indices identical
(ties broken by index) and distances BIT-identical for f32 and f64 (stricter than the
north star's 1e-5 / 1e-12 relative tolerances)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pn():
    import petal_neighbors_b200 as pn
    return pn


def bits(a):
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def assert_knn_equal(idx, dist, oi, od):
    assert np.array_equal(idx, oi.astype(np.uint64)), f"index mismatch at {np.argwhere(idx != oi.astype(np.uint64))[:5]}"
    assert np.array_equal(bits(dist), bits(od)), "distances are not bit-identical"


# ------------------------------------------------------------------ reference KATs on the GPU
def test_reference_kats(pn):
    # README.md:17-20, src/ball_tree.rs:73-77, :97-100, :132-135, src/vantage_point_tree.rs:82-86
    pts = np.array([[1., 1.], [1., 2.], [9., 9.]])
    bt = pn.BallTree.euclidean(pts)
    idx, dist = bt.query(np.array([3., 3.]), 2)
    assert idx.tolist() == [1, 0] and dist[0] == np.sqrt(5.) and dist[1] == np.sqrt(8.)
    i, d = bt.query_nearest(np.array([8., 8.]))
    assert i == 2 and abs(d - np.sqrt(2.)) < 1e-8
    i, d = pn.VantagePointTree.euclidean(pts).query_nearest(np.array([8., 8.]))
    assert i == 2 and abs(d - np.sqrt(2.)) < 1e-8
    bt = pn.BallTree.euclidean(np.array([[1., 0.], [2., 0.], [9., 0.]]))
    assert bt.query_radius(np.array([3., 0.]), 1.5).tolist() == [1]


def test_ball_tree_3_and_6(pn):
    # src/ball_tree.rs:649-716
    eps = np.finfo(np.float64).eps
    bt = pn.BallTree.euclidean(np.array([[1., 1.], [1., 1.1], [9., 9.]]))
    p = np.array([0., 0.])
    i, d = bt.query_nearest(p)
    assert i == 0 and abs(d - np.sqrt(2.)) <= eps
    idx, dist = bt.query(p, 0)
    assert idx.size == 0 and dist.size == 0
    idx, dist = bt.query(p, 1)
    assert idx.tolist() == [0] and dist[0] == d
    assert bt.query_radius(p, 2.).tolist() == [0, 1]
    assert bt.query_radius(np.array([20., 20.]), 1.).size == 0
    i, d = bt.query_nearest(np.array([1.1, 1.2]))
    assert i == 1 and abs(d - np.sqrt(2 * 0.1 * 0.1)) <= eps
    i, d = bt.query_nearest(np.array([7., 7.]))
    assert i == 2 and abs(d - np.sqrt(8.)) <= eps
    pts6 = np.array([[1.0, 2.0], [1.1, 2.2], [0.9, 1.9], [1.0, 2.1], [-2.0, 3.0], [-2.2, 3.1]])
    i, d = pn.BallTree.euclidean(pts6).query_nearest(np.array([1., 2.]))
    assert i == 0 and d == 0.0
    assert pn.VantagePointTree.euclidean(pts6).query_nearest(np.array([0.95, 1.96]))[0] == 0  # vantage_point_tree.rs:220-233


def test_identical_points_and_radius_1d(pn):
    # src/ball_tree.rs:718-740 (distance only is pinned; the engine breaks the tie by index -> 0)
    bt = pn.BallTree.euclidean(np.ones((8, 2)))
    i, d = bt.query_nearest(np.array([1., 2.]))
    assert d == 1.0 and i == 0
    assert bt.query_nearest(np.array([1., 1.]))[1] == 0.0
    idx, dist = bt.query(np.array([1., 2.]), 8)
    assert idx.tolist() == list(range(8)) and np.all(dist == 1.0)
    # src/ball_tree.rs:767-782
    bt = pn.BallTree.euclidean(np.array([[0.], [2.], [3.], [4.], [6.], [8.], [10.]]))
    assert bt.query_radius(np.array([0.1]), 1.).tolist() == [0]
    assert bt.query_radius(np.array([3.2]), 1.).tolist() == [2, 3]
    assert bt.query_radius(np.array([9.]), 0.9).size == 0


def test_errors(pn):
    with pytest.raises(pn.ArrayError) as e:
        pn.BallTree.euclidean(np.empty((0, 0)))
    assert e.value.kind == "Empty"
    with pytest.raises(pn.ArrayError) as e:
        pn.BallTree.euclidean(np.array([[1., 1.], [1., 1.1], [9., 9.]]).T)
    assert e.value.kind == "NotContiguous"
    with pytest.raises(pn.ArrayError):
        pn.VantagePointTree.euclidean(np.empty((0, 3), np.float32))


# ------------------------------------------------------------------ randomized parity vs the oracle
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,d,nq,k,bucket", [
    (1, 3, 5, 1, 0), (7, 2, 3, 10, 0), (300, 1, 64, 5, 8), (5000, 3, 1000, 10, 32),
    (5000, 5, 129, 16, 64), (20000, 16, 700, 10, 0), (4000, 17, 130, 10, 128), (3000, 33, 200, 3, 0),
    (3000, 64, 150, 10, 0), (2000, 100, 100, 10, 64), (2500, 128, 131, 10, 0), (6000, 8, 256, 40, 64),
    (3000, 10, 1, 10, 16), (1500, 7, 33, 100, 32),
    (900, 300, 70, 10, 64), (500, 1000, 40, 5, 0),   # rows wider than 1024 bytes: unstaged "wide" scan
])
def test_ball_knn_random(pn, oracle, dtype, n, d, nq, k, bucket):
    from petal_neighbors_b200 import synth
    pts = synth.uniform(n, d, 11 + n + d, dtype)
    Q = synth.uniform(nq, d, 12 + n + d, dtype)
    bt = pn.BallTree.euclidean(pts, bucket_size=bucket)
    idx, dist = bt.query_batch(Q, k)
    oi, od = oracle.brute_knn(pts, Q, k)
    assert_knn_equal(idx, dist, oi, od)
    if n > 1:
        c = bt.counters()
        assert 0 < c["pairs"] <= n * nq * max(1, -(-k // 16)) and c["kernel_launches"] > 0
    ni, nd = bt.query_nearest_batch(Q)
    assert np.array_equal(ni, oi[:, 0].astype(np.uint64)) and np.array_equal(bits(nd), bits(od[:, 0]))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_ties_broken_by_index(pn, oracle, dtype):
    """Integer lattice: masses of exactly equal distances, also straddling the k-th boundary."""
    rng = np.random.default_rng(7)
    pts = rng.integers(0, 6, size=(4000, 3)).astype(dtype)
    Q = rng.integers(0, 6, size=(300, 3)).astype(dtype)
    bt = pn.BallTree.euclidean(pts, bucket_size=32)
    for k in (1, 10, 16, 33):
        idx, dist = bt.query_batch(Q, k)
        oi, od = oracle.brute_knn(pts, Q, k)
        assert_knn_equal(idx, dist, oi, od)
    vp = pn.VantagePointTree.euclidean(pts, bucket_size=32)
    vi, vd = vp.query_nearest_batch(Q)
    oi, od = oracle.brute_knn(pts, Q, 1)
    assert np.array_equal(vi, oi[:, 0].astype(np.uint64)) and np.array_equal(bits(vd), bits(od[:, 0]))


def test_k_larger_than_n_is_padded(pn, oracle):
    pts = np.random.default_rng(3).random((7, 3)).astype(np.float32)
    bt = pn.BallTree.euclidean(pts)
    idx, dist = bt.query_batch(pts[:2], 20)
    oi, od = oracle.brute_knn(pts, pts[:2], 20)
    assert_knn_equal(idx, dist, oi, od)
    assert np.all(idx[:, 7:] == np.iinfo(np.uint64).max) and np.all(np.isinf(dist[:, 7:]))
    i, d = bt.query(pts[0], 20)  # reference: returns n results (S3)
    assert i.size == 7 and d.size == 7


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,d,nq,bucket,clustered", [
    (1, 4, 3, 0, False), (5000, 3, 500, 32, False), (20000, 64, 300, 0, True), (8000, 16, 257, 64, True),
    (3000, 2, 1, 16, False), (6000, 40, 100, 128, False),
])
def test_vp_nearest_random(pn, oracle, dtype, n, d, nq, bucket, clustered):
    from petal_neighbors_b200 import synth
    if clustered:
        pts = synth.gaussian_mixture(n, d, 5, n_centers=32, dtype=dtype)
        Q = synth.gaussian_mixture(nq, d, 6, n_centers=32, dtype=dtype)
    else:
        pts = synth.uniform(n, d, 21, dtype)
        Q = synth.uniform(nq, d, 22, dtype)
    vp = pn.VantagePointTree.euclidean(pts, bucket_size=bucket)
    vi, vd = vp.query_nearest_batch(Q)
    oi, od = oracle.brute_knn(pts, Q, 1)
    assert np.array_equal(vi, oi[:, 0].astype(np.uint64))
    assert np.array_equal(bits(vd), bits(od[:, 0]))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,d,nq,bucket,quant", [
    (5000, 3, 400, 32, 0.5), (20000, 3, 1000, 0, 0.2), (3000, 1, 50, 8, 0.5), (4000, 16, 200, 64, 0.5),
    (2000, 64, 40, 0, 0.9), (7, 2, 4, 0, 0.5),
])
def test_radius_random(pn, oracle, dtype, n, d, nq, bucket, quant):
    from petal_neighbors_b200 import synth
    pts = synth.uniform(n, d, 31, dtype)
    Q = synth.uniform(nq, d, 32, dtype)
    _, od = oracle.brute_knn(pts, Q, min(n, 30))
    r = dtype(np.quantile(od[:, -1], quant))
    bt = pn.BallTree.euclidean(pts, bucket_size=bucket)
    offs, ind = bt.query_radius_batch(Q, r)
    boffs, bind = oracle.brute_radius(pts, Q, r)
    assert np.array_equal(offs, boffs.astype(np.uint64))
    assert np.array_equal(ind, bind.astype(np.uint64))
    # radius exactly equal to a realised distance: strict `<` (src/ball_tree.rs:277)
    r_exact = od[0, min(n, 30) - 1]
    offs, ind = bt.query_radius_batch(Q[:1], r_exact)
    boffs, bind = oracle.brute_radius(pts, Q[:1], r_exact)
    assert np.array_equal(ind, bind.astype(np.uint64))
    # huge radius: whole-node inclusion path returns every point
    offs, ind = bt.query_radius_batch(Q[:3], dtype(1e6))
    assert offs.tolist() == [0, n, 2 * n, 3 * n] and np.array_equal(ind[:n], np.arange(n, dtype=np.uint64))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,d,nq,k,bucket,clustered", [
    (1, 3, 4, 1, 0, False), (7, 2, 5, 10, 0, False), (5000, 3, 400, 10, 32, False), (20000, 16, 600, 10, 0, False),
    (3000, 64, 150, 16, 0, True), (4000, 5, 200, 40, 16, False), (900, 24, 64, 3, 8, True),
])
def test_vp_knn_and_radius_extensions(pn, oracle, dtype, n, d, nq, k, bucket, clustered):
    """k-NN and radius search on a vantage-point handle return what BallTree returns for the same
    points (the reference VP tree has query_nearest only, src/vantage_point_tree.rs:88-98)."""
    from petal_neighbors_b200 import synth
    pts = synth.gaussian_mixture(n, d, 61, 16, 0.05, dtype=dtype) if clustered else synth.uniform(n, d, 61 + n, dtype)
    Q = synth.gaussian_mixture(nq, d, 62, 16, 0.05, dtype=dtype) if clustered else synth.uniform(nq, d, 62 + n, dtype)
    vp = pn.VantagePointTree.euclidean(pts, bucket_size=bucket)
    idx, dist = vp.query_batch(Q, k)
    oi, od = oracle.brute_knn(pts, Q, k)
    assert_knn_equal(idx, dist, oi, od)
    i1, d1 = vp.query(Q[0], k)
    m = min(k, n)
    assert np.array_equal(i1, oi[0, :m].astype(np.uintp)) and np.array_equal(bits(d1), bits(od[0, :m]))
    r = dtype(np.quantile(od[:, min(k, n) - 1], 0.5))
    offs, ind = vp.query_radius_batch(Q, r)
    boffs, bind = oracle.brute_radius(pts, Q, r)
    assert np.array_equal(offs, boffs.astype(np.uint64)) and np.array_equal(ind, bind.astype(np.uint64))
    # strict `<` at a realised distance, and the single-point form
    r_exact = od[0, m - 1]
    assert np.array_equal(vp.query_radius(Q[0], r_exact), oracle.brute_radius(pts, Q[:1], r_exact)[1].astype(np.uintp))
    # the nearest query still answers after the ball partition was attached
    ni, nd = vp.query_nearest_batch(Q)
    assert np.array_equal(ni, oi[:, 0].astype(np.uint64)) and np.array_equal(bits(nd), bits(od[:, 0]))


def test_empty_query_batches(pn):
    pts = np.random.default_rng(0).random((100, 3)).astype(np.float32)
    bt = pn.BallTree.euclidean(pts)
    idx, dist = bt.query_batch(np.empty((0, 3), np.float32), 5)
    assert idx.shape == (0, 5)
    offs, ind = bt.query_radius_batch(np.empty((0, 3), np.float32), 0.1)
    assert offs.tolist() == [0] and ind.size == 0


def test_strided_inputs(pn, oracle):
    rng = np.random.default_rng(9)
    big = rng.random((3000, 10)).astype(np.float32)
    pts = big[:, :6]           # row stride 10, unit column stride: accepted like the reference
    Qbig = rng.random((200, 9)).astype(np.float32)
    Q = Qbig[:, :6]
    bt = pn.BallTree.euclidean(pts, bucket_size=64)
    idx, dist = bt.query_batch(Q, 10)
    oi, od = oracle.brute_knn(np.ascontiguousarray(pts), np.ascontiguousarray(Q), 10)
    assert_knn_equal(idx, dist, oi, od)


def test_reference_traversal_agrees(pn, oracle):
    """The restated reference traversals (not only brute force) agree with the GPU on random
    continuous data: same indices, bit-identical distances."""
    from petal_neighbors_b200 import synth
    pts = synth.uniform(10000, 3, 1, np.float64)  # config 1 shape: every point is a query
    bt = pn.BallTree.euclidean(pts)
    idx, dist = bt.query_batch(pts, 10)
    ref = oracle.BallTree.euclidean(pts)
    oi, od, _ = ref.query_batch(pts, 10, n_threads=0)
    assert_knn_equal(idx, dist, oi, od)
    offs, ind = bt.query_radius_batch(pts[:2000], 0.05)
    roffs, rind = ref.query_radius_batch(pts[:2000], 0.05, n_threads=0)
    assert np.array_equal(offs, roffs.astype(np.uint64))
    for q in range(0, 2000, 97):
        assert sorted(rind[roffs[q]:roffs[q + 1]].tolist()) == ind[offs[q]:offs[q + 1]].tolist()


def test_sharded_by_subtree_merge(pn, oracle):
    """Points sharded by depth-2 subtree (4 shards), per-shard top-k merged on the GPU with
    pn_merge_topk_dev == unsharded result == oracle."""
    import torch
    from petal_neighbors_b200 import synth
    pts = synth.uniform(30000, 16, 41, np.float32)
    Q = synth.uniform(500, 16, 42, np.float32)
    k = 10
    shards = [pn.BallTree.euclidean(pts, shard_depth=2, shard_index=s, bucket_size=128) for s in range(4)]
    assert sum(t.info()["n_points"] for t in shards) == 30000
    qd = torch.from_numpy(Q).cuda()
    idx_l = torch.empty((4, 500, k), dtype=torch.int64, device="cuda")
    dist_l = torch.empty((4, 500, k), dtype=torch.float32, device="cuda")
    for s, t in enumerate(shards):
        t.query_knn_dev(qd.data_ptr(), 500, 16, k, idx_l[s].data_ptr(), dist_l[s].data_ptr())
    out_i = torch.empty((500, k), dtype=torch.int64, device="cuda")
    out_d = torch.empty((500, k), dtype=torch.float32, device="cuda")
    pn.merge_topk_dev(np.float32, 0, idx_l.data_ptr(), dist_l.data_ptr(), 4, 500, k, out_i.data_ptr(), out_d.data_ptr())
    oi, od = oracle.brute_knn(pts, Q, k)
    assert np.array_equal(out_i.cpu().numpy().astype(np.uint64), oi.astype(np.uint64))
    assert np.array_equal(bits(out_d.cpu().numpy()), bits(od))


def test_config2_full_size_properties(pn, oracle):
    """BASELINE config 2 at full size (1M x 16 f32, k = 10): size-independent properties over all
    queries (sorted rows, valid unique indices, self-consistency of returned distances) plus exact
    parity against the oracle on a seeded sample."""
    from petal_neighbors_b200 import synth
    n, d, k = 1_000_000, 16, 10
    pts = synth.uniform(n, d, 2, np.float32)
    nq = 200_000
    Q = synth.uniform(nq, d, 3, np.float32)
    bt = pn.BallTree.euclidean(pts)
    idx, dist = bt.query_batch(Q, k)
    assert np.all(np.diff(dist, axis=1) >= 0) and np.all(idx < n)
    srt = np.sort(idx, axis=1)
    assert np.all(srt[:, 1:] != srt[:, :-1])
    # recompute the returned distances with the sequential f32 fold (vectorised over rows)
    rows = np.arange(0, nq, 37)
    acc = np.zeros((rows.size, k), np.float32)
    for j in range(d):
        diff = Q[rows, j][:, None] - pts[idx[rows], j]
        acc = acc + diff * diff
    assert np.array_equal(bits(np.sqrt(acc)), bits(dist[rows]))
    sample = np.arange(0, nq, 401)[:400]
    oi, od = oracle.brute_knn(pts, Q[sample], k)
    assert_knn_equal(idx[sample], dist[sample], oi, od)


@pytest.mark.parametrize("dtype,n,d,k,algo", [
    (np.float64, 10000, 3, 10, 0), (np.float32, 5000, 16, 10, 1), (np.float32, 6000, 16, 10, 2),
    (np.float32, 3000, 5, 20, 0), (np.float32, 300, 2, 1, 0), (np.float32, 4100, 64, 5, 2),
])
def test_self_query(pn, oracle, dtype, n, d, k, algo):
    """pn_balltree_query_self: every stored point is a query (benches/ball_tree.rs:53-59); row i is the
    answer for points[i] regardless of the bucket order the engine stores the points in."""
    from petal_neighbors_b200 import synth
    pts = synth.uniform(n, d, 91 + n, dtype)
    pts[n // 2] = pts[n // 3]                      # an exact duplicate: tie at distance 0 broken by index
    bt = pn.BallTree.euclidean(pts, algo=algo, bucket_size=64)
    idx, dist = bt.query_self(k)
    oi, od = oracle.brute_knn(pts, pts, k)
    assert_knn_equal(idx, dist, oi, od)
    assert np.all(dist[:, 0] == 0)
    assert bt.counters()["h2d_bytes"] == 0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,d", [(2, 2), (1, 3), (33, 1), (100, 7), (257, 40), (64, 130)])
def test_pairwise(pn, oracle, dtype, n, d):
    """distance::pairwise (src/distance.rs:58-74): KAT :130-141 and bit-exact parity with the oracle."""
    assert pn.distance.pairwise(np.array([[3., 4.], [0., 0.]])).tolist() == [[0., 5.], [5., 0.]]
    assert pn.distance.pairwise(np.array([[0.]])).tolist() == [[0.]]
    x = np.random.default_rng(n * d).random((n, d)).astype(dtype)
    got = pn.distance.pairwise(x, pn.distance.Euclidean())
    want = oracle.pairwise(x)
    assert np.array_equal(bits(got), bits(want))
    assert np.array_equal(got, got.T) and np.all(np.diag(got) == 0)


def test_pairwise_row_blocks(pn, oracle):
    """n x n above the 256 MiB device buffer: the matrix is produced in row blocks (n = 9000: three of them)."""
    x = np.random.default_rng(77).random((9000, 5)).astype(np.float32)
    got = pn.distance.pairwise(x, pn.distance.Euclidean())
    want = oracle.pairwise(x)
    assert np.array_equal(bits(got), bits(want))


def test_concurrent_queries_on_one_handle(pn, oracle):
    """Queries on one tree from several host threads are legal (reference: &self queries,
    Euclidean: Sync, src/distance.rs:19); the engine serialises them and every caller gets its answer."""
    import threading
    from petal_neighbors_b200 import synth
    pts = synth.uniform(20000, 8, 5, np.float32)
    bt = pn.BallTree.euclidean(pts, bucket_size=64)
    Qs = [synth.uniform(500 + 37 * i, 8, 100 + i, np.float32) for i in range(6)]
    out = [None] * len(Qs)

    def work(i):
        if i % 2:
            out[i] = bt.query_batch(Qs[i], 10)
        else:
            out[i] = bt.query_radius_batch(Qs[i], np.float32(0.25))

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(Qs))]
    [t.start() for t in th]
    [t.join() for t in th]
    for i, Q in enumerate(Qs):
        if i % 2:
            oi, od = oracle.brute_knn(pts, Q, 10)
            assert_knn_equal(out[i][0], out[i][1], oi, od)
        else:
            boffs, bind = oracle.brute_radius(pts, Q, np.float32(0.25))
            assert np.array_equal(out[i][0], boffs.astype(np.uint64)) and np.array_equal(out[i][1], bind.astype(np.uint64))


@pytest.mark.parametrize("kind,d", [("ball", 3), ("ball", 32), ("vp", 24)])
def test_sessions_concurrent_callers(pn, oracle, kind, d):
    """pn_tree_session: every thread queries through its own session of one tree (own stream and workspaces, the tree's
    arrays shared); results equal the oracle's; the tree may be destroyed before its sessions."""
    import threading
    from petal_neighbors_b200 import synth
    n, nq, k = 30000, 4000, 10
    pts = synth.uniform(n, d, 71, np.float32)
    tree = (pn.BallTree if kind == "ball" else pn.VantagePointTree).euclidean(pts)
    Qs = [synth.uniform(nq, d, 80 + t, np.float32) for t in range(4)]
    sessions = [tree.session() for _ in Qs]
    assert sessions[0].info()["device_bytes"] == 0 and sessions[0].num_points() == n
    tree.close()   # the sessions keep the arrays alive
    out = [None] * len(Qs)

    def work(t):
        for _ in range(3):
            if kind == "ball":
                out[t] = sessions[t].query_batch(Qs[t], k) + sessions[t].query_radius_batch(Qs[t][:200], 0.2 if d == 3 else 1.2)
            else:
                ni, nd = sessions[t].query_nearest_batch(Qs[t])
                out[t] = (ni[:, None], nd[:, None])

    th = [threading.Thread(target=work, args=(t,)) for t in range(len(Qs))]
    [x.start() for x in th]
    [x.join() for x in th]
    for t, Q in enumerate(Qs):
        kk = k if kind == "ball" else 1
        oi, od = oracle.brute_knn(pts, Q, kk)
        assert_knn_equal(out[t][0], out[t][1], oi, od)
        if kind == "ball":
            boffs, bind = oracle.brute_radius(pts, Q[:200], np.float32(0.2 if d == 3 else 1.2))
            assert np.array_equal(out[t][2], boffs.astype(np.uint64)) and np.array_equal(out[t][3], bind.astype(np.uint64))
    s2 = sessions[0].session()   # a session of a session borrows from the same owner
    for x in sessions:
        x.close()
    oi, od = oracle.brute_knn(pts, Qs[0][:64], 1)
    ni, nd = s2.query_nearest_batch(Qs[0][:64])
    assert np.array_equal(ni, oi[:, 0].astype(np.uint64))
    s2.close()


def test_randomized_shapes_all_entry_points(pn, oracle):
    """Seeded sweep over random (n, d, nq, k, bucket, dtype, engine): k-NN, 1-NN, radius, VP 1-NN and
    self-query against the oracle; every case bit-exact."""
    rng = np.random.default_rng(20261018)
    for case in range(40):
        dtype = np.float32 if rng.random() < 0.6 else np.float64
        n = int(rng.integers(1, 4000))
        d = int(rng.choice([1, 2, 3, 4, 5, 8, 13, 16, 17, 31, 32, 33, 64, 70]))
        nq = int(rng.integers(1, 400))
        k = int(rng.choice([1, 2, 5, 10, 16, 17, 31]))
        bucket = int(rng.choice([0, 8, 16, 64, 256]))
        algo = int(rng.choice([0, 1, 2])) if dtype == np.float32 else 1
        quant = bool(rng.random() < 0.3)                 # lattice data: many exact ties
        pts = rng.random((n, d))
        Q = rng.random((nq, d))
        if quant:
            pts, Q = np.round(pts * 4), np.round(Q * 4)
        pts, Q = pts.astype(dtype), Q.astype(dtype)
        tag = f"case {case}: n={n} d={d} nq={nq} k={k} bucket={bucket} {dtype.__name__} algo={algo} quant={quant}"
        bt = pn.BallTree.euclidean(pts, bucket_size=bucket, algo=algo)
        oi, od = oracle.brute_knn(pts, Q, k)
        idx, dist = bt.query_batch(Q, k)
        assert np.array_equal(idx, oi.astype(np.uint64)) and np.array_equal(bits(dist), bits(od)), tag
        ni, nd = bt.query_nearest_batch(Q)
        assert np.array_equal(ni, oi[:, 0].astype(np.uint64)) and np.array_equal(bits(nd), bits(od[:, 0])), tag
        r = dtype(np.quantile(od[:, min(k, n) - 1], 0.5))
        offs, ind = bt.query_radius_batch(Q, r)
        boffs, bind = oracle.brute_radius(pts, Q, r)
        assert np.array_equal(offs, boffs.astype(np.uint64)) and np.array_equal(ind, bind.astype(np.uint64)), tag
        vi, vd = pn.VantagePointTree.euclidean(pts, bucket_size=bucket, algo=algo).query_nearest_batch(Q)
        assert np.array_equal(vi, oi[:, 0].astype(np.uint64)) and np.array_equal(bits(vd), bits(od[:, 0])), tag
        if case % 4 == 0:
            si, sdist = bt.query_self(min(k, 16))
            soi, sod = oracle.brute_knn(pts, pts, min(k, 16))
            assert np.array_equal(si, soi.astype(np.uint64)) and np.array_equal(bits(sdist), bits(sod)), tag
