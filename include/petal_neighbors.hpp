// petal_neighbors.hpp -- C++ host-side mirror of the petal-neighbors public API over the C ABI
// (include/petal_b200.h).  The reference is a Rust crate and no Rust toolchain exists in the
// build image, so this header is the compiled-language host binding that is actually built and
// tested here; it keeps the reference's names, argument meaning and error behaviour:
//
//   petal_neighbors::BallTree<A>::euclidean / query / query_nearest / query_radius / num_points
//                                                   (reference src/ball_tree.rs:38-142, 351-373)
//   petal_neighbors::VantagePointTree<A>::euclidean / query_nearest
//                                                   (reference src/vantage_point_tree.rs:31-98)
//     + query / query_radius on the vantage-point tree: extensions, the reference has none
//   petal_neighbors::ArrayError {Empty, NotContiguous}          (reference src/lib.rs:9-16)
//   petal_neighbors::distance::{Metric, Euclidean}              (reference src/distance.rs:9-55)
//
// Construction errors are thrown as ArrayError (Rust: Result<_, ArrayError>); every other
// non-zero status throws std::runtime_error (Rust shim: panic).  Single-point methods are
// batches of one; *_batch methods are the additions a GPU engine needs.
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "petal_b200.h"

namespace petal_neighbors {

struct ArrayError : std::runtime_error {
    enum Kind { Empty, NotContiguous } kind;
    explicit ArrayError(Kind k)
        : std::runtime_error(k == Empty ? "array is empty" : "array is not contiguous in memory"), kind(k) {}
};

namespace distance {
template <typename A> struct Metric {
    virtual ~Metric() = default;
    virtual A distance(const A* x1, const A* x2, size_t d) const = 0;
    virtual A rdistance(const A* x1, const A* x2, size_t d) const = 0;
    virtual A rdistance_to_distance(A d) const = 0;
    virtual A distance_to_rdistance(A d) const = 0;
};
// Euclidean: sequential fold, separate multiply and add (compile with -ffp-contract=off)
template <typename A> struct EuclideanT final : Metric<A> {
    A rdistance(const A* x1, const A* x2, size_t d) const override {
        A sum = A(0);
        for (size_t j = 0; j < d; ++j) { A diff = x1[j] - x2[j]; sum += diff * diff; }
        return sum;
    }
    A distance(const A* x1, const A* x2, size_t d) const override { return std::sqrt(rdistance(x1, x2, d)); }
    A rdistance_to_distance(A d) const override { return std::sqrt(d); }
    A distance_to_rdistance(A d) const override { return d * d; }
    bool operator==(const EuclideanT&) const { return true; }
};
struct Euclidean {
    bool operator==(const Euclidean&) const { return true; }
};
}  // namespace distance

namespace detail {
inline void check_create(int32_t st) {
    if (st == PN_OK) return;
    if (st == PN_EMPTY) throw ArrayError(ArrayError::Empty);
    if (st == PN_NOT_CONTIGUOUS) throw ArrayError(ArrayError::NotContiguous);
    throw std::runtime_error(std::string("petal_b200: ") + pn_last_error_message());
}
inline void check(int32_t st) {
    if (st != PN_OK) throw std::runtime_error(std::string("petal_b200: ") + pn_last_error_message());
}
// The reference's distance fold zips the two rows and silently truncates to the shorter one
// (src/distance.rs:26-35); the C ABI reads exactly `d` elements per query row, so a wrong-sized query would be
// an out-of-bounds read.  Every wrapper method checks the dimension first.
inline void check_dim(size_t got, size_t want, const char* what) {
    if (got != want)
        throw std::invalid_argument(std::string("petal_neighbors: ") + what + " has dimension " + std::to_string(got) +
                                    ", the tree has " + std::to_string(want));
}
template <typename A> struct Abi;
template <> struct Abi<float> {
    static constexpr auto ball_create = pn_balltree_create_f32;
    static constexpr auto vp_create = pn_vptree_create_f32;
    static constexpr auto query = pn_balltree_query_f32;
    static constexpr auto nearest = pn_balltree_query_nearest_f32;
    static constexpr auto radius = pn_balltree_query_radius_f32;
    static constexpr auto vp_nearest = pn_vptree_query_nearest_f32;
    static constexpr auto vp_query = pn_vptree_query_f32;
    static constexpr auto vp_radius = pn_vptree_query_radius_f32;
    static constexpr auto self = pn_balltree_query_self_f32;
};
template <> struct Abi<double> {
    static constexpr auto ball_create = pn_balltree_create_f64;
    static constexpr auto vp_create = pn_vptree_create_f64;
    static constexpr auto query = pn_balltree_query_f64;
    static constexpr auto nearest = pn_balltree_query_nearest_f64;
    static constexpr auto radius = pn_balltree_query_radius_f64;
    static constexpr auto vp_nearest = pn_vptree_query_nearest_f64;
    static constexpr auto vp_query = pn_vptree_query_f64;
    static constexpr auto vp_radius = pn_vptree_query_radius_f64;
    static constexpr auto self = pn_balltree_query_self_f64;
};
}  // namespace detail

// A borrowed row-major 2-D view: the ndarray ArrayView2 of this binding.
template <typename A> struct View2 {
    const A* data; size_t rows, cols, row_stride, col_stride;
    View2(const A* p, size_t r, size_t c) : data(p), rows(r), cols(c), row_stride(c), col_stride(1) {}
    View2(const A* p, size_t r, size_t c, size_t rs, size_t cs) : data(p), rows(r), cols(c), row_stride(rs), col_stride(cs) {}
    View2 reversed_axes() const { return View2(data, cols, rows, col_stride, row_stride); }
};

template <typename A> class BallTree {
  public:
    distance::Euclidean metric;
    static BallTree euclidean(const View2<A>& points, const pn_build_opts* opts = nullptr) { return BallTree(points, opts); }
    static BallTree make(const View2<A>& points, distance::Euclidean, const pn_build_opts* opts = nullptr) { return BallTree(points, opts); }  // BallTree::new
    BallTree(BallTree&& o) noexcept : h_(o.h_), n_(o.n_), d_(o.d_) { o.h_ = nullptr; }
    BallTree(const BallTree&) = delete;
    ~BallTree() { if (h_) pn_tree_destroy(h_); }

    std::pair<size_t, A> query_nearest(const std::vector<A>& point) const {
        detail::check_dim(point.size(), d_, "point");
        uint64_t i = 0; A dist = 0;
        detail::check(detail::Abi<A>::nearest(h_, point.data(), 1, d_, &i, &dist));
        return {size_t(i), dist};
    }
    std::pair<std::vector<size_t>, std::vector<A>> query(const std::vector<A>& point, size_t k) const {
        detail::check_dim(point.size(), d_, "point");
        if (k == 0) return {};
        std::vector<uint64_t> idx(k); std::vector<A> dist(k);
        detail::check(detail::Abi<A>::query(h_, point.data(), 1, d_, k, idx.data(), dist.data()));
        size_t m = k < n_ ? k : n_;
        return {std::vector<size_t>(idx.begin(), idx.begin() + m), std::vector<A>(dist.begin(), dist.begin() + m)};
    }
    std::vector<size_t> query_radius(const std::vector<A>& point, A distance) const {
        detail::check_dim(point.size(), d_, "point");
        auto r = query_radius_batch(View2<A>(point.data(), 1, d_), distance);
        return r.second;
    }
    // batched additions
    void query_batch(const View2<A>& q, size_t k, uint64_t* idx_out, A* dist_out) const {
        check_view(q);
        detail::check(detail::Abi<A>::query(h_, q.data, q.rows, q.row_stride, k, idx_out, dist_out));
    }
    std::pair<std::vector<size_t>, std::vector<size_t>> query_radius_batch(const View2<A>& q, A distance) const {
        check_view(q);
        uint64_t *po = nullptr, *pi = nullptr;
        detail::check(detail::Abi<A>::radius(h_, q.data, q.rows, q.row_stride, distance, &po, &pi));
        std::vector<size_t> offs(po, po + q.rows + 1), ind(pi, pi + po[q.rows]);
        pn_free(po); pn_free(pi);
        return {offs, ind};
    }
    // every stored point as a query (benches/ball_tree.rs:53-59): row-major n x k outputs
    void query_self(size_t k, uint64_t* idx_out, A* dist_out) const { detail::check(detail::Abi<A>::self(h_, k, idx_out, dist_out)); }
    size_t num_points() const { return n_; }
    pn_tree* handle() const { return h_; }
    // a second handle onto the same device-resident tree (own stream and workspaces) for a concurrent caller
    BallTree session() const {
        BallTree s(n_, d_);
        detail::check(pn_tree_session(h_, &s.h_));
        return s;
    }

  private:
    void check_view(const View2<A>& q) const {
        detail::check_dim(q.cols, d_, "query batch");
        if (q.cols > 1 && q.col_stride != 1) throw std::invalid_argument("petal_neighbors: query rows must be contiguous");
        if (q.rows > 1 && q.row_stride < q.cols) throw std::invalid_argument("petal_neighbors: query row stride < dimension");
    }
    BallTree(const View2<A>& p, const pn_build_opts* opts) : n_(p.rows), d_(p.cols) {
        detail::check_create(detail::Abi<A>::ball_create(p.data, p.rows, p.cols, p.row_stride, p.col_stride, opts, &h_));
    }
    BallTree(size_t n, size_t d) : n_(n), d_(d) {}   // session(): the handle is filled in by pn_tree_session
    pn_tree* h_ = nullptr;
    size_t n_, d_;
};

template <typename A> class VantagePointTree {
  public:
    static VantagePointTree euclidean(const View2<A>& points, const pn_build_opts* opts = nullptr) { return VantagePointTree(points, opts); }
    VantagePointTree(VantagePointTree&& o) noexcept : h_(o.h_), d_(o.d_) { o.h_ = nullptr; }
    VantagePointTree(const VantagePointTree&) = delete;
    ~VantagePointTree() { if (h_) pn_tree_destroy(h_); }
    std::pair<size_t, A> query_nearest(const std::vector<A>& needle) const {
        detail::check_dim(needle.size(), d_, "needle");
        uint64_t i = 0; A dist = 0;
        detail::check(detail::Abi<A>::vp_nearest(h_, needle.data(), 1, d_, &i, &dist));
        return {size_t(i), dist};
    }
    void query_nearest_batch(const View2<A>& q, uint64_t* idx_out, A* dist_out) const {
        detail::check_dim(q.cols, d_, "query batch");
        if ((q.cols > 1 && q.col_stride != 1) || (q.rows > 1 && q.row_stride < q.cols))
            throw std::invalid_argument("petal_neighbors: query rows must be contiguous with row stride >= dimension");
        detail::check(detail::Abi<A>::vp_nearest(h_, q.data, q.rows, q.row_stride, idx_out, dist_out));
    }
    // Extensions (not in the reference): what BallTree::query / query_radius return for the same points.
    std::pair<std::vector<size_t>, std::vector<A>> query(const std::vector<A>& point, size_t k) const {
        detail::check_dim(point.size(), d_, "point");
        if (k == 0) return {};
        std::vector<uint64_t> idx(k); std::vector<A> dist(k);
        detail::check(detail::Abi<A>::vp_query(h_, point.data(), 1, d_, k, idx.data(), dist.data()));
        size_t m = 0;
        while (m < k && idx[m] != ~uint64_t(0)) ++m;   // rows are padded when k > n
        return {std::vector<size_t>(idx.begin(), idx.begin() + m), std::vector<A>(dist.begin(), dist.begin() + m)};
    }
    std::vector<size_t> query_radius(const std::vector<A>& point, A distance) const {
        detail::check_dim(point.size(), d_, "point");
        uint64_t *po = nullptr, *pi = nullptr;
        detail::check(detail::Abi<A>::vp_radius(h_, point.data(), 1, d_, distance, &po, &pi));
        std::vector<size_t> ind(pi, pi + po[1]);
        pn_free(po); pn_free(pi);
        return ind;
    }

  private:
    VantagePointTree(const View2<A>& p, const pn_build_opts* opts) : d_(p.cols) {
        detail::check_create(detail::Abi<A>::vp_create(p.data, p.rows, p.cols, p.row_stride, p.col_stride, opts, &h_));
    }
    pn_tree* h_ = nullptr;
    size_t d_;
};

}  // namespace petal_neighbors
