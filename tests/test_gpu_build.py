"""Device-side tree construction (csrc/gpu_build.cu; reference build_subtree / Node::init / max_spread_column /
halve_node_indices, src/ball_tree.rs:445-613, and create_node, src/vantage_point_tree.rs:146-197) against the host builder, which stays as the checker: the flattened layouts
must be bit-identical (ids, bucket ranges, centroids, radii, point rows), also on data with masses of equal column values
(ties broken by original index on both sides), for shards, and for points that never leave the device."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pn():
    import petal_neighbors_b200 as pn
    return pn


def bits(a):
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def same_layout(a, b):
    la, lb = a.layout(), b.layout()
    for key in ("n_points", "n_points_total", "dim", "dim_padded", "n_levels", "n_buckets", "n_nodes", "bucket_size_max"):
        assert la[key] == lb[key], key
    assert np.array_equal(la["bucket_lo"], lb["bucket_lo"]) and np.array_equal(la["bucket_hi"], lb["bucket_hi"])
    assert np.array_equal(la["ids"], lb["ids"]), f"ids differ at {np.argwhere(la['ids'] != lb['ids'])[:5].ravel()}"
    assert np.array_equal(bits(la["points"]), bits(lb["points"]))
    assert np.array_equal(bits(la["node_center"]), bits(lb["node_center"])), "centroids differ"
    assert np.array_equal(bits(la["node_radius"]), bits(lb["node_radius"])), "radii differ"


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,d,bucket,kind", [
    (1, 3, 0, "u"), (2, 2, 8, "u"), (7, 5, 8, "u"), (300, 1, 8, "u"), (5000, 3, 32, "u"), (20000, 16, 0, "u"),
    (4097, 17, 64, "u"), (33000, 64, 0, "u"), (9000, 130, 128, "u"), (3000, 300, 0, "u"),
    (6000, 3, 16, "lattice"), (40000, 8, 64, "lattice"), (5000, 4, 8, "const"), (70000, 10, 0, "mix"),
    (300000, 3, 0, "u"), (200000, 32, 0, "u"),
])
def test_device_builder_matches_host(pn, dtype, n, d, bucket, kind):
    from petal_neighbors_b200 import synth
    rng = np.random.default_rng(n + d)
    if kind == "u":
        pts = synth.uniform(n, d, 100 + n + d, dtype)
    elif kind == "lattice":      # many exactly equal column values: the (value, index) order decides the split
        pts = rng.integers(0, 5, size=(n, d)).astype(dtype)
    elif kind == "const":        # every column constant: spread 0 everywhere, column 0, pure index split
        pts = np.full((n, d), 0.25, dtype)
    else:
        pts = synth.gaussian_mixture(n, d, 5, n_centers=16, dtype=dtype)
        pts[::7] = -pts[::7]     # negative values and a signed zero
        pts[5, 0] = -0.0
    host = pn.BallTree.euclidean(pts, bucket_size=bucket, builder=pn.PN_BUILDER_HOST)
    dev = pn.BallTree.euclidean(pts, bucket_size=bucket, builder=pn.PN_BUILDER_DEVICE)
    same_layout(host, dev)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,d,bucket,kind", [
    (1, 3, 0, "u"), (7, 5, 8, "u"), (9, 2, 8, "u"), (300, 1, 8, "u"), (5000, 3, 32, "u"), (20000, 16, 0, "u"), (4097, 17, 64, "u"),
    (33000, 64, 0, "u"), (9000, 130, 128, "u"), (6000, 3, 16, "lattice"), (40000, 8, 64, "lattice"), (5000, 4, 8, "const"),
    (70000, 10, 0, "mix"), (300000, 3, 0, "u"), (200000, 32, 0, "u"),
])
def test_device_vp_builder_matches_host(pn, dtype, n, d, bucket, kind):
    """Vantage-point trees: vantage point = last element of the sorted slice = maximum of the (distance, id) key, near / far
    by a select at rank (len-1)/2, buckets sorted at the end -- the stored order, vantage points and thresholds must equal
    the host builder's sort-based construction, also where masses of distances tie (lattice, constant data)."""
    from petal_neighbors_b200 import synth
    rng = np.random.default_rng(n + d + 1)
    if kind == "u":
        pts = synth.uniform(n, d, 200 + n + d, dtype)
    elif kind == "lattice":
        pts = rng.integers(0, 5, size=(n, d)).astype(dtype)
    elif kind == "const":
        pts = np.full((n, d), 0.25, dtype)
    else:
        pts = synth.gaussian_mixture(n, d, 5, n_centers=16, dtype=dtype)
    host = pn.VantagePointTree.euclidean(pts, bucket_size=bucket, builder=pn.PN_BUILDER_HOST)
    dev = pn.VantagePointTree.euclidean(pts, bucket_size=bucket, builder=pn.PN_BUILDER_DEVICE)
    same_layout(host, dev)


def test_device_vp_builder_queries(pn, oracle):
    """the pruned SIMT traversal on a device-built vantage-point tree (the tensor path of a VP handle uses the ball partition)"""
    from petal_neighbors_b200 import synth
    pts = synth.gaussian_mixture(50000, 24, 5, n_centers=64, dtype=np.float32)
    Q = synth.gaussian_mixture(3000, 24, 6, n_centers=64, dtype=np.float32)
    oi, od = oracle.brute_knn(pts, Q, 1)
    for algo in (pn.PN_ALGO_SIMT, pn.PN_ALGO_AUTO):
        vp = pn.VantagePointTree.euclidean(pts, builder=pn.PN_BUILDER_DEVICE, algo=algo)
        vi, vd = vp.query_nearest_batch(Q)
        assert np.array_equal(vi, oi[:, 0].astype(np.uint64)) and np.array_equal(bits(vd), bits(od[:, 0]))
    p64 = pts.astype(np.float64)
    vp = pn.VantagePointTree.euclidean(p64, builder=pn.PN_BUILDER_DEVICE)
    oi, od = oracle.brute_knn(p64, Q.astype(np.float64), 1)
    vi, vd = vp.query_nearest_batch(Q.astype(np.float64))
    assert np.array_equal(vi, oi[:, 0].astype(np.uint64)) and np.array_equal(bits(vd), bits(od[:, 0]))


@pytest.mark.parametrize("n,depth", [(10001, 3), (50000, 2), (5, 3)])
def test_device_builder_shards(pn, n, depth):
    pts = np.random.default_rng(5).random((n, 6)).astype(np.float32)
    seen = []
    for s in range(1 << depth):
        try:
            host = pn.BallTree.euclidean(pts, shard_depth=depth, shard_index=s, bucket_size=64, builder=pn.PN_BUILDER_HOST)
        except pn.ArrayError:
            with pytest.raises(pn.ArrayError):
                pn.BallTree.euclidean(pts, shard_depth=depth, shard_index=s, bucket_size=64, builder=pn.PN_BUILDER_DEVICE)
            continue
        dev = pn.BallTree.euclidean(pts, shard_depth=depth, shard_index=s, bucket_size=64, builder=pn.PN_BUILDER_DEVICE)
        same_layout(host, dev)
        seen.append(dev.layout()["ids"])
    assert sorted(np.concatenate(seen).tolist()) == list(range(n))


def test_device_resident_points(pn, oracle):
    """pn_balltree_create_dev_*: the points are a CUDA tensor (with a row stride), nothing crosses PCIe; same layout as the
    host build of the same rows, and queries on it are exact."""
    import torch
    from petal_neighbors_b200 import synth
    big = synth.uniform_torch(60000, 24, 3, torch.float32)      # row stride 24, 20 columns used
    view = big[:, :20]
    dev = pn.BallTree.euclidean(view)
    pts = view.cpu().numpy().copy()
    host = pn.BallTree.euclidean(pts, builder=pn.PN_BUILDER_HOST)
    same_layout(host, dev)
    Q = synth.uniform(700, 20, 4, np.float32)
    idx, dist = dev.query_batch(Q, 10)
    oi, od = oracle.brute_knn(pts, Q, 10)
    assert np.array_equal(idx, oi.astype(np.uint64)) and np.array_equal(bits(dist), bits(od))
    offs, ind = dev.query_radius_batch(Q[:100], np.float32(0.9))
    boffs, bind = oracle.brute_radius(pts, Q[:100], np.float32(0.9))
    assert np.array_equal(offs, boffs.astype(np.uint64)) and np.array_equal(ind, bind.astype(np.uint64))
    big64 = synth.uniform_torch(40000, 3, 5, torch.float64)
    dev64 = pn.BallTree.euclidean(big64, bucket_size=32)
    same_layout(pn.BallTree.euclidean(big64.cpu().numpy(), bucket_size=32, builder=pn.PN_BUILDER_HOST), dev64)


def test_device_builder_at_scale(pn, oracle):
    """10M x 128 built on the device in well under the host builder's time; exact k-NN on it."""
    import time
    import torch
    from petal_neighbors_b200 import synth
    n, d = 10_000_000, 128
    pts_dev = synth.uniform_torch(n, d, 2, torch.float32)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bt = pn.BallTree.euclidean(pts_dev)
    build_s = time.perf_counter() - t0
    inf = bt.info()
    assert inf["n_points"] == n and inf["n_levels"] == 16
    assert build_s < 3.0, f"device build took {build_s:.2f} s"
    Q = synth.fast_uniform(4096, d, 3, np.float32)
    idx, dist = bt.query_batch(Q, 10)
    sample = np.arange(0, 4096, 64)
    oi, od = oracle.brute_knn(pts_dev.cpu().numpy(), Q[sample], 10)
    assert np.array_equal(idx[sample], oi.astype(np.uint64)) and np.array_equal(bits(dist[sample]), bits(od))
