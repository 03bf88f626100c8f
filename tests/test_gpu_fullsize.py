"""The BASELINE configurations at FULL size on the GPU, through the C ABI: exact parity against the oracle on seeded
samples of the queries, plus size-independent properties over the whole output (sorted rows, valid unique indices,
distances that recompute bit-exactly from the returned indices, CSR consistency).

  C3  VantagePointTree 1M x 64 f32 Gaussian mixture, 1M queries, 1-NN       (tensor path and the pruned SIMT scan)
  C4  BallTree::query_radius 10M x 3 f32, 1M queries, r = 0.01               (+ k-NN on the same tree)
  T   BallTree 10M x 128 f32 uniform, 100k queries, k = 10                   (north-star target shape: the launch plan
      with two subtiles x two stages and a tail wave split over the 78 125 point tiles)
  C5-shaped: 1M x 128 f32 "SIFT-shaped" mixture sharded by the eight depth-3 subtrees, merged with pn_merge_topk_dev
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pn():
    import petal_neighbors_b200 as pn
    return pn


def bits(a):
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def recompute(pts, Q, rows, idx):
    """sequential f32 fold of the returned neighbours (vectorised over rows; numpy never fuses mul+add)"""
    acc = np.zeros((rows.size, idx.shape[1]), pts.dtype)
    for j in range(pts.shape[1]):
        diff = Q[rows, j][:, None] - pts[idx[rows], j]
        acc = acc + diff * diff
    return np.sqrt(acc)


def test_c3_vp_1m_x_64_mixture(pn, oracle):
    from petal_neighbors_b200 import synth
    n = nq = 1_000_000
    pts = synth.fast_gaussian_mixture(n, 64, 5, n_centers=1024, sigma=0.05, center_seed=4)
    Q = synth.fast_gaussian_mixture(nq, 64, 6, n_centers=1024, sigma=0.05, center_seed=4)
    sample = np.arange(0, nq, nq // 320)[:320]
    oi, od = oracle.brute_knn(pts, Q[sample], 1)
    for algo in (pn.PN_ALGO_AUTO, pn.PN_ALGO_SIMT):
        vp = pn.VantagePointTree.euclidean(pts, algo=algo)
        qq = Q if algo == pn.PN_ALGO_AUTO else Q[:200_000]     # the pruned scan is 30x slower: a fifth of the batch
        vi, vd = vp.query_nearest_batch(qq)
        c = vp.counters()
        assert (c["filter_pairs"] > 0) == (algo == pn.PN_ALGO_AUTO)
        assert np.all(vi < n)
        rows = np.arange(0, qq.shape[0], 53)
        assert np.array_equal(bits(recompute(pts, qq, rows, vi[:, None].astype(np.int64))[:, 0]), bits(vd[rows]))
        s = sample[sample < qq.shape[0]]
        m = s.size
        assert np.array_equal(vi[s], oi[:m, 0].astype(np.uint64)) and np.array_equal(bits(vd[s]), bits(od[:m, 0]))
        del vp


def test_c4_radius_10m_x_3(pn, oracle):
    from petal_neighbors_b200 import synth
    n, nq, r = 10_000_000, 1_000_000, np.float32(0.01)
    pts = synth.fast_uniform(n, 3, 7, np.float32)
    Q = synth.fast_uniform(nq, 3, 8, np.float32)
    bt = pn.BallTree.euclidean(pts)
    offs, ind = bt.query_radius_batch(Q, r)
    assert offs[0] == 0 and np.all(np.diff(offs.astype(np.int64)) >= 0) and offs[-1] == ind.size
    assert 35 < ind.size / nq < 48                     # ~ n * 4/3 pi r^3 = 41.9 expected hits (fewer near the faces)
    assert np.all(ind < n)
    # every reported pair is inside the radius, on the bit-exact distance (all pairs, vectorised)
    qrow = np.repeat(np.arange(nq), np.diff(offs.astype(np.int64)))
    acc = np.zeros(ind.size, np.float32)
    for j in range(3):
        diff = Q[qrow, j] - pts[ind, j]
        acc = acc + diff * diff
    assert np.all(np.sqrt(acc) < r)
    # ascending within each query
    inner = np.ones(ind.size, bool)
    inner[offs[1:-1][offs[1:-1] < ind.size]] = False
    assert np.all((np.diff(ind.astype(np.int64)) > 0) | ~inner[1:])
    sample = np.arange(0, nq, nq // 256)[:256]
    boffs, bind = oracle.brute_radius(pts, Q[sample], r)
    for t, qi in enumerate(sample):
        assert np.array_equal(ind[offs[qi]:offs[qi + 1]], bind[boffs[t]:boffs[t + 1]].astype(np.uint64)), f"query {qi}"
    # k-NN on the same tree (d = 3: the pruned SIMT scan at 10M points)
    idx, dist = bt.query_batch(Q, 10)
    assert np.all(np.diff(dist, axis=1) >= 0) and np.all(idx < n)
    oi, od = oracle.brute_knn(pts, Q[sample], 10)
    assert np.array_equal(idx[sample], oi.astype(np.uint64)) and np.array_equal(bits(dist[sample]), bits(od))


def test_t_10m_x_128(pn, oracle):
    from petal_neighbors_b200 import synth
    n, d, nq, k = 10_000_000, 128, 100_000, 10
    pts = synth.fast_uniform(n, d, 2, np.float32)
    Q = synth.fast_uniform(nq, d, 3, np.float32)
    bt = pn.BallTree.euclidean(pts)
    idx, dist = bt.query_batch(Q, k)
    c = bt.counters()
    assert c["filter_pairs"] > 0 and c["pairs"] <= n * nq
    assert np.all(np.diff(dist, axis=1) >= 0) and np.all(idx < n)
    srt = np.sort(idx, axis=1)
    assert np.all(srt[:, 1:] != srt[:, :-1])
    rows = np.arange(0, nq, 97)
    assert np.array_equal(bits(recompute(pts, Q, rows, idx.astype(np.int64))), bits(dist[rows]))
    # samples from the whole waves AND from the tail launch (the last rows) that splits the point stream
    sample = np.concatenate([np.arange(0, nq, nq // 192)[:192], np.arange(nq - 64, nq)])
    oi, od = oracle.brute_knn(pts, Q[sample], k)
    assert np.array_equal(idx[sample], oi.astype(np.uint64)) and np.array_equal(bits(dist[sample]), bits(od))


def test_c5_shaped_eight_shards(pn, oracle):
    """C5's structure at 1M x 128: mixture of 4096 centres (sigma 0.1, clipped), points sharded by the 8 depth-3 subtrees,
    every shard answers all queries, pn_merge_topk_dev merges the eight sorted lists."""
    import torch
    from petal_neighbors_b200 import synth
    n, d, nq, k = 1_000_000, 128, 40_000, 10
    pts = synth.fast_gaussian_mixture(n, d, 9, n_centers=4096, sigma=0.1, center_seed=4, clip=True)
    Q = synth.fast_gaussian_mixture(nq, d, 10, n_centers=4096, sigma=0.1, center_seed=4, clip=True)
    qd = torch.from_numpy(Q).cuda()
    idx_l = torch.empty((8, nq, k), dtype=torch.int64, device="cuda")
    dist_l = torch.empty((8, nq, k), dtype=torch.float32, device="cuda")
    total = 0
    for s in range(8):
        t = pn.BallTree.euclidean(pts, shard_depth=3, shard_index=s)
        total += t.info()["n_points"]
        t.query_knn_dev(qd.data_ptr(), nq, d, k, idx_l[s].data_ptr(), dist_l[s].data_ptr())
        del t
    assert total == n
    out_i = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    out_d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    pn.merge_topk_dev(np.float32, 0, idx_l.data_ptr(), dist_l.data_ptr(), 8, nq, k, out_i.data_ptr(), out_d.data_ptr())
    idx, dist = out_i.cpu().numpy(), out_d.cpu().numpy()
    assert np.all(np.diff(dist, axis=1) >= 0)
    sample = np.arange(0, nq, nq // 320)[:320]
    oi, od = oracle.brute_knn(pts, Q[sample], k)
    assert np.array_equal(idx[sample].astype(np.uint64), oi.astype(np.uint64)) and np.array_equal(bits(dist[sample]), bits(od))
