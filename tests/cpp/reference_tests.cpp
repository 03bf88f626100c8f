// The reference's own unit tests and doctests for the hot path, restated against the C++ host
// mirror (include/petal_neighbors.hpp) so they read like the originals.  Run on a GPU box by
// tests/test_gpu_cpp_mirror.py.  Citations are into the reference (petal-neighbors v0.18.0).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>

#include "petal_neighbors.hpp"

using namespace petal_neighbors;
static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)
static bool abs_diff_eq(double a, double b) { return std::fabs(a - b) <= std::numeric_limits<double>::epsilon(); }

static void ball_tree_empty() {  // src/ball_tree.rs:623-630
    bool got = false;
    try { BallTree<double>::euclidean(View2<double>(nullptr, 0, 0)); } catch (const ArrayError& e) { got = e.kind == ArrayError::Empty; }
    CHECK(got);
}
static void ball_tree_column_base() {  // src/ball_tree.rs:632-638
    const double a[] = {1., 1., 1., 1.1, 9., 9.};
    bool got = false;
    try { BallTree<double>::euclidean(View2<double>(a, 3, 2).reversed_axes()); } catch (const ArrayError& e) { got = e.kind == ArrayError::NotContiguous; }
    CHECK(got);
}
static void ball_tree_metric() {  // src/ball_tree.rs:640-647
    const double a[] = {1., 1., 1., 1.1, 9., 9.};
    auto t = BallTree<double>::make(View2<double>(a, 3, 2), distance::Euclidean{});
    auto t1 = BallTree<double>::euclidean(View2<double>(a, 3, 2));
    CHECK(t.metric == t1.metric);
}
static void ball_tree_3() {  // src/ball_tree.rs:649-698
    const double a[] = {1., 1., 1., 1.1, 9., 9.};
    auto tree = BallTree<double>::euclidean(View2<double>(a, 3, 2));
    std::vector<double> point{0., 0.};
    auto neighbor = tree.query_nearest(point);
    CHECK(neighbor.first == 0);
    CHECK(abs_diff_eq(neighbor.second, std::sqrt(2.)));
    auto r0 = tree.query(point, 0);
    CHECK(r0.first.empty() && r0.second.empty());
    auto r1 = tree.query(point, 1);
    CHECK(r1.first.size() == 1 && r1.second.size() == 1);
    CHECK(r1.first[0] == neighbor.first);
    CHECK(abs_diff_eq(r1.second[0], neighbor.second));
    auto neighbors = tree.query_radius(point, 2.);
    std::sort(neighbors.begin(), neighbors.end());
    CHECK((neighbors == std::vector<size_t>{0, 1}));
    CHECK(tree.query_radius({20., 20.}, 1.).empty());
    point = {1.1, 1.2};
    neighbor = tree.query_nearest(point);
    CHECK(neighbor.first == 1);
    CHECK(abs_diff_eq(neighbor.second, std::sqrt(2. * 0.1 * 0.1)));
    point = {7., 7.};
    neighbor = tree.query_nearest(point);
    CHECK(neighbor.first == 2);
    CHECK(abs_diff_eq(neighbor.second, std::sqrt(8.)));
    r1 = tree.query(point, 1);
    CHECK(r1.first[0] == neighbor.first && abs_diff_eq(r1.second[0], neighbor.second));
}
static void ball_tree_6() {  // src/ball_tree.rs:700-716
    const double a[] = {1.0, 2.0, 1.1, 2.2, 0.9, 1.9, 1.0, 2.1, -2.0, 3.0, -2.2, 3.1};
    auto tree = BallTree<double>::euclidean(View2<double>(a, 6, 2));
    auto neighbor = tree.query_nearest({1., 2.});
    CHECK(neighbor.first == 0 && abs_diff_eq(neighbor.second, 0.));
}
static void ball_tree_identical_points() {  // src/ball_tree.rs:718-740
    std::vector<double> a(16, 1.0);
    auto tree = BallTree<double>::euclidean(View2<double>(a.data(), 8, 2));
    CHECK(abs_diff_eq(tree.query_nearest({1., 2.}).second, 1.));
    CHECK(abs_diff_eq(tree.query_nearest({1., 1.}).second, 0.));
}
static void ball_tree_query_radius() {  // src/ball_tree.rs:767-782
    const double a[] = {0., 2., 3., 4., 6., 8., 10.};
    auto bt = BallTree<double>::euclidean(View2<double>(a, 7, 1));
    CHECK((bt.query_radius({0.1}, 1.) == std::vector<size_t>{0}));
    auto n = bt.query_radius({3.2}, 1.);
    std::sort(n.begin(), n.end());
    CHECK((n == std::vector<size_t>{2, 3}));
    CHECK(bt.query_radius({9.}, 0.9).empty());
}
static void doctests() {  // src/ball_tree.rs:69-78, 93-101, 128-136; src/vantage_point_tree.rs:78-87; README.md:13-21
    const double a[] = {1., 1., 1., 2., 9., 9.};
    auto tree = BallTree<double>::euclidean(View2<double>(a, 3, 2));
    auto n = tree.query_nearest({8., 8.});
    CHECK(n.first == 2 && std::fabs(std::sqrt(2.) - n.second) < 1e-8);
    auto q = tree.query({3., 3.}, 2);
    CHECK((q.first == std::vector<size_t>{1, 0}));
    auto vp = VantagePointTree<double>::euclidean(View2<double>(a, 3, 2));
    auto v = vp.query_nearest({8., 8.});
    CHECK(v.first == 2 && std::fabs(std::sqrt(2.) - v.second) < 1e-8);
    const double b[] = {1., 0., 2., 0., 9., 0.};
    auto t2 = BallTree<double>::euclidean(View2<double>(b, 3, 2));
    CHECK((t2.query_radius({3., 0.}, 1.5) == std::vector<size_t>{1}));
}
static void vp_euclidian() {  // src/vantage_point_tree.rs:220-233
    const double a[] = {1.0, 2.0, 1.1, 2.2, 0.9, 1.9, 1.0, 2.1, -2.0, 3.0, -2.2, 3.1};
    auto vp = VantagePointTree<double>::euclidean(View2<double>(a, 6, 2));
    CHECK(vp.query_nearest({0.95, 1.96}).first == 0);
    // extensions (not in the reference): the answers BallTree gives for the same points
    auto bt = BallTree<double>::euclidean(View2<double>(a, 6, 2));
    auto kq = vp.query({0.95, 1.96}, 4), kb = bt.query({0.95, 1.96}, 4);
    CHECK(kq.first == kb.first && kq.second == kb.second);
    CHECK(vp.query({0.95, 1.96}, 9).first.size() == 6);
    CHECK(vp.query_radius({0.95, 1.96}, 0.3) == bt.query_radius({0.95, 1.96}, 0.3));
    auto s2 = bt.session();   // a second handle onto the same tree
    CHECK(s2.query({0.95, 1.96}, 4).first == kb.first && s2.num_points() == 6);
}
static void ball_tree_query_property() {  // src/ball_tree.rs:742-765: tree distances == naive distances
    const size_t N = 40, D = 3;
    std::vector<double> pts(N * D);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&] { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return double(s >> 11) * (1.0 / 9007199254740992.0); };
    for (auto& v : pts) v = rnd();
    auto bt = BallTree<double>::euclidean(View2<double>(pts.data(), N, D));
    distance::EuclideanT<double> euclid;
    for (int t = 0; t < 10; ++t) {
        std::vector<double> q{rnd(), rnd(), rnd()};
        auto r = bt.query(q, 5);
        std::vector<double> naive(N);
        for (size_t i = 0; i < N; ++i) naive[i] = euclid.distance(&pts[i * D], q.data(), D);
        std::sort(naive.begin(), naive.end());
        for (int i = 0; i < 5; ++i) CHECK(abs_diff_eq(r.second[i], naive[i]));
    }
}

int main() {
    ball_tree_empty(); ball_tree_column_base(); ball_tree_metric(); ball_tree_3(); ball_tree_6();
    ball_tree_identical_points(); ball_tree_query_radius(); doctests(); vp_euclidian(); ball_tree_query_property();
    std::printf(failures ? "%d FAILURES\n" : "all reference tests passed (%d failures)\n", failures);
    return failures ? 1 : 0;
}
