"""Pins the CPU oracle against every known-answer vector the reference holds for the hot path
(SURVEY.md 8c items 1-11).  All citations are into /root/reference (petal-neighbors v0.18.0).
Distances whose expected value is a closed form use the reference's own tolerance
(approx::abs_diff_eq! default epsilon = machine epsilon)."""
import numpy as np
import pytest

EPS = np.finfo(np.float64).eps


def test_readme_query_k2(oracle):
    # README.md:17-20 == src/ball_tree.rs:97-100
    t = oracle.BallTree.euclidean(np.array([[1., 1.], [1., 2.], [9., 9.]]))
    idx, dist = t.query(np.array([3., 3.]), 2)
    assert idx.tolist() == [1, 0]
    assert abs(dist[0] - np.sqrt(5.)) <= EPS * 4 and abs(dist[1] - np.sqrt(8.)) <= EPS * 4


def test_doctest_query_nearest(oracle):
    # src/ball_tree.rs:73-77 and src/vantage_point_tree.rs:82-86
    pts = np.array([[1., 1.], [1., 2.], [9., 9.]])
    for cls in (oracle.BallTree, oracle.VantagePointTree):
        i, d = cls.euclidean(pts).query_nearest(np.array([8., 8.]))
        assert i == 2 and abs(np.sqrt(2.) - d) < 1e-8


def test_doctest_query_radius(oracle):
    # src/ball_tree.rs:132-135
    t = oracle.BallTree.euclidean(np.array([[1., 0.], [2., 0.], [9., 0.]]))
    assert t.query_radius(np.array([3., 0.]), 1.5).tolist() == [1]


def test_ball_tree_empty_and_column_base(oracle):
    # src/ball_tree.rs:623-638
    with pytest.raises(oracle.ArrayError, match="empty"):
        oracle.BallTree.euclidean(np.empty((0, 0)))
    arr = np.array([[1., 1.], [1., 1.1], [9., 9.]])
    with pytest.raises(oracle.ArrayError, match="contiguous"):
        oracle.BallTree.euclidean(arr.T)  # reversed_axes(): Fortran order
    with pytest.raises(oracle.ArrayError, match="empty"):
        oracle.VantagePointTree.euclidean(np.empty((0, 2)))


def test_ball_tree_3(oracle):
    # src/ball_tree.rs:649-698
    t = oracle.BallTree.euclidean(np.array([[1., 1.], [1., 1.1], [9., 9.]]))
    p = np.array([0., 0.])
    i, d = t.query_nearest(p)
    assert i == 0 and abs(d - np.sqrt(2.)) <= EPS
    idx, dist = t.query(p, 0)
    assert idx.size == 0 and dist.size == 0
    idx, dist = t.query(p, 1)
    assert idx.tolist() == [0] and abs(dist[0] - d) <= EPS
    assert sorted(t.query_radius(p, 2.).tolist()) == [0, 1]
    assert t.nearest_neighbor_in_subtree(np.array([20., 20.]), 0, 1.) is None
    assert t.query_radius(np.array([20., 20.]), 1.).tolist() == []
    p = np.array([1.1, 1.2])
    i, d = t.query_nearest(p)
    assert i == 1 and abs(d - np.sqrt(2. * 0.1 * 0.1)) <= EPS
    idx, dist = t.query(p, 1)
    assert idx.tolist() == [1] and abs(dist[0] - d) <= EPS
    p = np.array([7., 7.])
    i, d = t.query_nearest(p)
    assert i == 2 and abs(d - np.sqrt(8.)) <= EPS
    idx, dist = t.query(p, 1)
    assert idx.tolist() == [2] and abs(dist[0] - d) <= EPS


def test_ball_tree_6(oracle):
    # src/ball_tree.rs:700-716
    pts = np.array([[1.0, 2.0], [1.1, 2.2], [0.9, 1.9], [1.0, 2.1], [-2.0, 3.0], [-2.2, 3.1]])
    i, d = oracle.BallTree.euclidean(pts).query_nearest(np.array([1., 2.]))
    assert i == 0 and d == 0.0


def test_ball_tree_identical_points(oracle):
    # src/ball_tree.rs:718-740 (the reference asserts the distance only).  The idx permutation equals the survey model's
    # (SURVEY.md 7 step 1).  The returned index is 7, not the 5 SURVEY.md quotes.  Hand trace of :149-196 for q = [1, 2]:
    #   n = 8 -> height 4, 15 nodes, leaves 7..14 hold idx[0..8] = [7, 2, 1, 0, 3, 4, 6, 5], one point each; every centroid
    #   is [1, 1] with radius 0, so every node's lower bound is exactly 1.
    #   * `lb1 < lb2` is false on equal bounds (:183-187), so at every internal node child2 is searched FIRST:
    #     root -> node 2 -> node 6 -> leaf 14 = idx[7] = 5 : Some((5, 1)).
    #   * the sibling is then searched with radius = 1 (:189-191).  `lower_bound > radius` is 1 > 1 = false (:156), the leaf
    #     returns Some because `min_dist <= radius` is 1 <= 1 (:174), and `.map_or(Some(neighbor), Some)` (:191) maps a Some
    #     result through `Some`, i.e. the SECOND child's answer replaces the first on equality:
    #     leaf 13 = idx[6] = 6 replaces 5; node 5 -> (leaf 12 = 4, then leaf 11 = 3) -> 3 replaces 6 at node 2;
    #     node 1 -> node 4 -> (leaf 10 = 0, then leaf 9 = 1) -> 1; node 3 -> (leaf 8 = 2, then leaf 7 = 7) -> 7 replaces 1;
    #     at the root 7 replaces 3.
    #   The last leaf visited wins, and that is leaf 7 = idx[0] = 7.  (5 is what comes out if the first answer is kept on
    #   equality -- the survey's throwaway model evidently read map_or the other way round.)  The engine itself breaks ties
    #   by index (0 here), which the north star asks for; the reference pins neither.
    t = oracle.BallTree.euclidean(np.ones((8, 2)))
    assert t.idx.tolist() == [7, 2, 1, 0, 3, 4, 6, 5]
    i, d = t.query_nearest(np.array([1., 2.]))
    assert d == 1.0 and i == 7
    _, d = t.query_nearest(np.array([1., 1.]))
    assert d == 0.0


def test_ball_tree_query_radius_1d(oracle):
    # src/ball_tree.rs:767-782
    t = oracle.BallTree.euclidean(np.array([[0.], [2.], [3.], [4.], [6.], [8.], [10.]]))
    assert t.query_radius(np.array([0.1]), 1.).tolist() == [0]
    assert sorted(t.query_radius(np.array([3.2]), 1.).tolist()) == [2, 3]
    assert t.query_radius(np.array([9.]), 0.9).size == 0


def test_ball_tree_query_property(oracle):
    # src/ball_tree.rs:742-765: tree distances == brute-force distances (40x3, k=5)
    rng = np.random.default_rng(1234)
    for trial in range(20):
        pts = rng.random((40, 3))
        t = oracle.BallTree.euclidean(pts)
        for _ in range(10):
            q = rng.random(3)
            _, dist = t.query(q, 5)
            _, bd = oracle.brute_knn(pts, q[None, :], 5)
            assert np.array_equal(dist, bd[0])


def test_vp_euclidian(oracle):
    # src/vantage_point_tree.rs:220-233
    pts = np.array([[1.0, 2.0], [1.1, 2.2], [0.9, 1.9], [1.0, 2.1], [-2.0, 3.0], [-2.2, 3.1]])
    assert oracle.VantagePointTree.euclidean(pts).query_nearest(np.array([0.95, 1.96]))[0] == 0


def test_node_init(oracle):
    # src/ball_tree.rs:784-798
    arr = np.array([[0., 1.], [0., 9.], [0., 2.]])
    c, r = oracle.node_init(arr, [0, 1, 2])
    assert c.tolist() == [0., 4.] and abs(r - 5.) <= EPS
    c, r = oracle.node_init(arr, [0, 2])
    assert c.tolist() == [0., 1.5] and abs(r - 0.5) <= EPS


def test_halve_node_indices(oracle):
    # src/ball_tree.rs:800-835
    with pytest.raises(OverflowError):
        oracle.halve_node_indices(np.array([], dtype=np.uintp), np.array([], dtype=np.float64))
    assert oracle.halve_node_indices([0], np.array([1.])).tolist() == [0]
    idx = oracle.halve_node_indices([0, 1, 4, 3, 2], np.array([1., 2., 3., 4., 5.]))
    assert idx[0] < idx[2] and idx[1] < idx[2] and idx[2] <= idx[3] and idx[2] <= idx[4]
    idx = oracle.halve_node_indices([3, 2, 1, 0], np.array([1., 2., 3., 4.]))
    assert idx[0] < idx[2] and idx[1] < idx[2] and idx[2] <= idx[3]


def test_max_spread_column(oracle):
    # src/ball_tree.rs:837-866
    data = np.array([[0., 1.], [0., 9.], [0., 2.]])
    assert oracle.max_spread_column(data, [0, 1, 2]) == 1
    with pytest.raises(RuntimeError, match="empty matrix"):
        oracle.max_spread_column(data, [])
    with pytest.raises(RuntimeError, match="index out of bounds"):
        oracle.max_spread_column(data, [0, 4, 2])
    with pytest.raises(RuntimeError, match="empty matrix"):
        oracle.max_spread_column(np.empty((0, 0)), [0, 1, 2])


def test_pairwise(oracle):
    # src/distance.rs:130-141
    assert oracle.pairwise(np.array([[3., 4.], [0., 0.]])).tolist() == [[0., 5.], [5., 0.]]
    assert oracle.pairwise(np.array([[0.]])).tolist() == [[0.]]


def test_tree_shape(oracle):
    # src/ball_tree.rs:51-52: height = bit_length(n), size = 2^height - 1
    for n, size in ((1, 1), (3, 3), (4, 7), (7, 7), (8, 15), (1000, 1023)):
        t = oracle.BallTree.euclidean(np.random.default_rng(n).random((n, 2)))
        assert t.num_nodes() == size and t.num_points() == n


def test_k_larger_than_n_and_f32(oracle):
    # S3: k > n returns n results; f32 trees are supported (CHANGELOG.md:123)
    pts = np.random.default_rng(5).random((7, 3)).astype(np.float32)
    t = oracle.BallTree.euclidean(pts)
    idx, dist = t.query(pts[0], 20)
    assert idx.size == 7 and dist.dtype == np.float32 and np.all(np.diff(dist) >= 0)
    bi, bd = oracle.brute_knn(pts, pts[:1], 20)
    assert np.array_equal(bd[0, :7], dist) and np.all(bi[0, 7:] == np.iinfo(np.uintp).max)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("d", [1, 3, 10, 16])
def test_tree_traversals_equal_brute_force(oracle, dtype, d):
    """Both reference traversals (restated) return the brute-force answer on continuous data;
    distances bit-identical, indices identical (no ties in random data)."""
    rng = np.random.default_rng(100 + d)
    pts = rng.random((500, d)).astype(dtype)
    Q = rng.random((64, d)).astype(dtype)
    bt = oracle.BallTree.euclidean(pts)
    vp = oracle.VantagePointTree.euclidean(pts)
    bi, bd = oracle.brute_knn(pts, Q, 10)
    oi, od, _ = bt.query_batch(Q, 10, n_threads=2)
    assert np.array_equal(od, bd) and np.array_equal(oi, bi)
    vi, vd, _ = vp.query_nearest_batch(Q, n_threads=2)
    assert np.array_equal(vi, bi[:, 0]) and np.array_equal(vd, bd[:, 0])
    for qi in range(8):
        i, dd = bt.query_nearest(Q[qi])
        assert i == bi[qi, 0] and dd == bd[qi, 0]
    r = dtype(np.median(bd[:, -1]))
    offs, ind = bt.query_radius_batch(Q, r, n_threads=2)
    boffs, bind = oracle.brute_radius(pts, Q, r)
    assert np.array_equal(offs, boffs)
    for qi in range(Q.shape[0]):
        assert sorted(ind[offs[qi]:offs[qi + 1]].tolist()) == bind[boffs[qi]:boffs[qi + 1]].tolist()
