// flat_tree.hpp -- host-side tree construction and flattening.
//
// "Tree construction stays on the host" (BASELINE.json north_star).  The builders below apply
// the reference's partition rules but stop at cut level L, where a node still holds a bucket of
// ~bucket_size points (the reference recurses down to 1-2 point leaves, src/ball_tree.rs:51-52,
// which on a GPU would turn every query into pointer chasing).  Exact k-NN / radius results do
// not depend on where the partition is cut.
//
//   ball tree  : src/ball_tree.rs:445-461 (Node::init: centroid = mean in idx order, radius =
//                max distance), :577-613 (max-spread column, first strictly-greatest wins),
//                :545-569 + :535-537 (median split at mid = (start+end)/2)
//   vp tree    : src/vantage_point_tree.rs:146-197 (vantage point = last element of the slice,
//                rest sorted by distance, near = first half, mu = far[0].distance)
//
// Both trees are stored as an implicit complete binary tree (children of i are 2i+1, 2i+2,
// src/ball_tree.rs:180-181): internal nodes 0 .. 2^L-2, buckets are the 2^L nodes of level L.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace petal {

template <typename A> struct Vec { static constexpr int N = 16 / sizeof(A); };  // float4 / double2

template <typename A>
struct FlatTree {
    int kind = 0;                 // 0 ball, 1 vp
    uint64_t n = 0, n_total = 0;  // points in this handle / in the caller's array
    uint32_t d = 0, dpad = 0;
    uint32_t L = 0;               // cut level
    uint32_t n_internal = 0;      // 2^L - 1
    uint32_t n_buckets = 0;       // 2^L
    uint32_t n_nodes = 0;         // ball: 2^(L+1)-1 (internal + bucket-level); vp: n_internal
    uint32_t bucket_max = 0;
    std::vector<A> pts;           // n x dpad, bucket order (zero padded columns)
    std::vector<uint32_t> ids;    // n: original row index of the i-th stored point
    std::vector<uint32_t> bucket_lo, bucket_hi;  // n_buckets
    std::vector<A> centers;       // n_nodes x dpad: ball centroid / vp vantage point
    std::vector<A> radii;         // n_nodes: ball radius (-1 = empty node) / vp threshold mu
    std::vector<uint32_t> vp_ids; // vp: original index of each internal node's vantage point
    std::vector<uint32_t> vp_pos; // vp: position (row in pts) of each internal node's vantage point
};

// Euclidean::distance, src/distance.rs:26-35 (host copy used by the builders only).
template <typename A>
static inline A fold_distance(const A* x1, const A* x2, size_t d) {
    A sum = A(0);
    for (size_t j = 0; j < d; ++j) {
        A diff = x1[j] - x2[j];
        sum += diff * diff;
    }
    return std::sqrt(sum);
}

static inline uint32_t choose_levels(uint64_t n, uint32_t bucket_size) {
    uint32_t L = 0;
    while (((n + ((uint64_t(1) << L) - 1)) >> L) > bucket_size) ++L;  // ceil(n / 2^L) <= bucket
    return L;
}

template <typename A>
class BallBuilder {
  public:
    BallBuilder(const A* pts, size_t n, size_t d, size_t stride, uint32_t threads)
        : pts_(pts), n_(n), d_(d), stride_(stride), threads_(threads ? threads : 1) {}

    // Applies `depth` levels of the reference split to idx[lo,hi) and returns the range of
    // subtree `index` at that depth (point sharding by subtree).
    void shard_range(std::vector<uint32_t>& idx, uint32_t depth, uint32_t index, size_t& lo, size_t& hi) {
        lo = 0; hi = n_;
        for (uint32_t lev = 0; lev < depth; ++lev) {
            if (hi - lo < 2) break;
            split(idx, lo, hi);
            size_t mid = (lo + hi) / 2;
            bool right = (index >> (depth - 1 - lev)) & 1u;
            if (right) lo = mid; else hi = mid;
        }
    }

    void build(std::vector<uint32_t>& idx, size_t lo, size_t hi, uint32_t bucket_size, FlatTree<A>& t) {
        const size_t n = hi - lo;
        t.kind = 0;
        t.n = n; t.n_total = n_;
        t.d = (uint32_t)d_;
        t.dpad = (uint32_t)((d_ + Vec<A>::N - 1) / Vec<A>::N * Vec<A>::N);
        t.L = choose_levels(n, bucket_size);
        t.n_internal = (1u << t.L) - 1;
        t.n_buckets = 1u << t.L;
        t.n_nodes = (1u << (t.L + 1)) - 1;
        t.bucket_lo.assign(t.n_buckets, 0);
        t.bucket_hi.assign(t.n_buckets, 0);
        t.centers.assign((size_t)t.n_nodes * t.dpad, A(0));
        t.radii.assign(t.n_nodes, A(-1));
        base_ = lo;
        recurse(idx, t, 0, lo, hi, 0);
        // flatten: points in bucket (idx) order, zero-padded to dpad
        t.pts.assign(n * (size_t)t.dpad, A(0));
        t.ids.resize(n);
        parallel_for(n, [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) {
                uint32_t src = idx[lo + i];
                t.ids[i] = src;
                std::memcpy(&t.pts[i * t.dpad], pts_ + (size_t)src * stride_, d_ * sizeof(A));
            }
        });
        t.bucket_max = 0;
        for (uint32_t b = 0; b < t.n_buckets; ++b)
            t.bucket_max = std::max(t.bucket_max, t.bucket_hi[b] - t.bucket_lo[b]);
    }

  private:
    template <typename F> void parallel_for(size_t n, F f) {
        uint32_t nt = (uint32_t)std::min<size_t>(threads_, (n + 65535) / 65536);
        if (nt <= 1) { f(0, n); return; }
        std::vector<std::thread> th;
        size_t chunk = (n + nt - 1) / nt;
        for (uint32_t i = 0; i < nt; ++i) {
            size_t b = i * chunk, e = std::min(n, b + chunk);
            if (b >= e) break;
            th.emplace_back([=] { f(b, e); });
        }
        for (auto& x : th) x.join();
    }

    // max_spread_column (src/ball_tree.rs:577-613) + halve_node_indices (:545-569).
    // The split *set* equals the reference's whenever the column values are distinct; the
    // selection algorithm is std::nth_element on (value, index) instead of the reference's
    // last-element-pivot quick-select, which is quadratic on sorted or constant columns.
    void split(std::vector<uint32_t>& idx, size_t lo, size_t hi) {
        const size_t len = hi - lo;
        std::vector<A> mn(d_), mx(d_);
        const A* r0 = pts_ + (size_t)idx[lo] * stride_;
        for (size_t j = 0; j < d_; ++j) mn[j] = mx[j] = r0[j];
        auto minmax_range = [&](size_t b, size_t e, A* pmn, A* pmx) {  // row-wise pass: same result as the per-column pass
            for (size_t t = b; t < e; ++t) {
                const A* r = pts_ + (size_t)idx[t] * stride_;
                for (size_t j = 0; j < d_; ++j) {
                    A v = r[j];
                    if (v < pmn[j]) pmn[j] = v;
                    if (v > pmx[j]) pmx[j] = v;
                }
            }
        };
        const uint32_t nt = workers_for(len);
        if (nt <= 1) {
            minmax_range(lo + 1, hi, mn.data(), mx.data());
        } else {  // min / max are order-independent: chunked over threads, identical result
            std::vector<std::vector<A>> pmn(nt, mn), pmx(nt, mx);
            std::vector<std::thread> th;
            const size_t chunk = (len + nt - 1) / nt;
            for (uint32_t w = 0; w < nt; ++w) {
                const size_t b = lo + w * chunk, e = std::min(hi, b + chunk);
                if (b >= e) break;
                th.emplace_back([&, b, e, w] { minmax_range(b, e, pmn[w].data(), pmx[w].data()); });
            }
            for (auto& x : th) x.join();
            for (uint32_t w = 0; w < nt; ++w)
                for (size_t j = 0; j < d_; ++j) { if (pmn[w][j] < mn[j]) mn[j] = pmn[w][j]; if (pmx[w][j] > mx[j]) mx[j] = pmx[w][j]; }
        }
        size_t col = 0;
        A best = mx[0] - mn[0];
        for (size_t j = 1; j < d_; ++j) {
            A s = mx[j] - mn[j];
            if (s > best) { best = s; col = j; }
        }
        const A* c = pts_ + col;
        const size_t stride = stride_;
        std::nth_element(idx.begin() + lo, idx.begin() + lo + len / 2, idx.begin() + hi,
                         [c, stride](uint32_t a, uint32_t b) {
                             A va = c[(size_t)a * stride], vb = c[(size_t)b * stride];
                             return va < vb || (va == vb && a < b);
                         });
    }

    // threads worth using for a pass over `len` rows at the current point of the recursion
    uint32_t workers_for(size_t len) const {
        if (len * d_ < (size_t)1 << 22) return 1;
        return (uint32_t)std::min<size_t>(threads_, len / 65536 + 1);
    }

    // Node::init, src/ball_tree.rs:445-461.  For large nodes the sum and the max are accumulated per chunk
    // and combined (the centroid then differs from the strictly sequential sum in the last bits; any centre
    // with radius = max exact fold distance to it is a valid ball, so results are unaffected).
    void node_init(const std::vector<uint32_t>& idx, size_t lo, size_t hi, A* center, A& radius) {
        const size_t len = hi - lo;
        if (len == 0) { radius = A(-1); return; }
        const uint32_t nt = workers_for(len);
        const size_t chunk = (len + nt - 1) / nt;
        auto sum_range = [&](size_t b, size_t e, A* acc) {
            for (size_t j = 0; j < d_; ++j) acc[j] = A(0);
            for (size_t t = b; t < e; ++t) {
                const A* r = pts_ + (size_t)idx[t] * stride_;
                for (size_t j = 0; j < d_; ++j) acc[j] += r[j];
            }
        };
        auto max_range = [&](size_t b, size_t e) {
            A m = A(0);
            for (size_t t = b; t < e; ++t) {
                A v = fold_distance(center, pts_ + (size_t)idx[t] * stride_, d_);
                if (v > m) m = v;
            }
            return m;
        };
        if (nt <= 1) {
            sum_range(lo, hi, center);
        } else {
            std::vector<std::vector<A>> part(nt, std::vector<A>(d_));
            std::vector<std::thread> th;
            for (uint32_t w = 0; w < nt; ++w) {
                const size_t b = lo + w * chunk, e = std::min(hi, b + chunk);
                if (b >= e) { for (size_t j = 0; j < d_; ++j) part[w][j] = A(0); continue; }
                th.emplace_back([&, b, e, w] { sum_range(b, e, part[w].data()); });
            }
            for (auto& x : th) x.join();
            for (size_t j = 0; j < d_; ++j) { A sacc = A(0); for (uint32_t w = 0; w < nt; ++w) sacc += part[w][j]; center[j] = sacc; }
        }
        A flen = (A)len;
        for (size_t j = 0; j < d_; ++j) center[j] /= flen;
        if (nt <= 1) {
            radius = max_range(lo, hi);
        } else {
            std::vector<A> pm(nt, A(0));
            std::vector<std::thread> th;
            for (uint32_t w = 0; w < nt; ++w) {
                const size_t b = lo + w * chunk, e = std::min(hi, b + chunk);
                if (b >= e) continue;
                th.emplace_back([&, b, e, w] { pm[w] = max_range(b, e); });
            }
            for (auto& x : th) x.join();
            A m = A(0);
            for (uint32_t w = 0; w < nt; ++w) if (pm[w] > m) m = pm[w];
            radius = m;
        }
    }

    void recurse(std::vector<uint32_t>& idx, FlatTree<A>& t, uint32_t node, size_t lo, size_t hi, uint32_t level) {
        node_init(idx, lo, hi, &t.centers[(size_t)node * t.dpad], t.radii[node]);
        if (level == t.L) {
            uint32_t b = node - t.n_internal;
            t.bucket_lo[b] = (uint32_t)(lo - base_);
            t.bucket_hi[b] = (uint32_t)(hi - base_);
            return;
        }
        if (hi - lo >= 2) split(idx, lo, hi);
        size_t mid = (lo + hi) / 2;
        // fork the left subtree onto its own thread near the top of the tree
        if ((1u << level) < threads_ && hi - lo > 32768) {
            std::thread th([&, node, lo, mid, level] { recurse(idx, t, 2 * node + 1, lo, mid, level + 1); });
            recurse(idx, t, 2 * node + 2, mid, hi, level + 1);
            th.join();
        } else {
            recurse(idx, t, 2 * node + 1, lo, mid, level + 1);
            recurse(idx, t, 2 * node + 2, mid, hi, level + 1);
        }
    }

    const A* pts_;
    size_t n_, d_, stride_;
    uint32_t threads_;
    size_t base_ = 0;
};

template <typename A>
class VpBuilder {
    struct DI { A dist; uint32_t id; };  // DistanceIndex, src/vantage_point_tree.rs:209-212

  public:
    VpBuilder(const A* pts, size_t n, size_t d, size_t stride, uint32_t threads)
        : pts_(pts), n_(n), d_(d), stride_(stride), threads_(threads ? threads : 1) {}

    void build(uint32_t bucket_size, FlatTree<A>& t) {
        t.kind = 1;
        t.n = n_; t.n_total = n_;
        t.d = (uint32_t)d_;
        t.dpad = (uint32_t)((d_ + Vec<A>::N - 1) / Vec<A>::N * Vec<A>::N);
        t.L = choose_levels(n_, bucket_size);
        t.n_internal = (1u << t.L) - 1;
        t.n_buckets = 1u << t.L;
        t.n_nodes = t.n_internal;
        t.bucket_lo.assign(t.n_buckets, 0);
        t.bucket_hi.assign(t.n_buckets, 0);
        t.centers.assign((size_t)std::max<uint32_t>(t.n_nodes, 1) * t.dpad, A(0));
        t.radii.assign(std::max<uint32_t>(t.n_nodes, 1), A(0));
        t.vp_ids.assign(std::max<uint32_t>(t.n_nodes, 1), 0);
        t.vp_pos.assign(std::max<uint32_t>(t.n_nodes, 1), 0);
        std::vector<DI> ix(n_);  // create_root, :132-144
        for (size_t i = 0; i < n_; ++i) { ix[i].dist = std::numeric_limits<A>::max(); ix[i].id = (uint32_t)i; }
        recurse(ix, t, 0, 0, n_, 0);
        t.pts.assign(n_ * (size_t)t.dpad, A(0));
        t.ids.resize(n_);
        for (size_t i = 0; i < n_; ++i) {
            t.ids[i] = ix[i].id;
            std::memcpy(&t.pts[i * t.dpad], pts_ + (size_t)ix[i].id * stride_, d_ * sizeof(A));
        }
        t.bucket_max = 0;
        for (uint32_t b = 0; b < t.n_buckets; ++b)
            t.bucket_max = std::max(t.bucket_max, t.bucket_hi[b] - t.bucket_lo[b]);
    }

  private:
    // create_node, src/vantage_point_tree.rs:146-197, stopped at level L.  The slice keeps the
    // reference's order [near | far | vantage point].
    void recurse(std::vector<DI>& ix, FlatTree<A>& t, uint32_t node, size_t lo, size_t hi, uint32_t level) {
        if (level == t.L) {
            uint32_t b = node - t.n_internal;
            t.bucket_lo[b] = (uint32_t)lo;
            t.bucket_hi[b] = (uint32_t)hi;
            return;
        }
        const size_t len = hi - lo;  // >= 2 by choice of L (bucket_size >= 8)
        const size_t vp_pos = hi - 1;
        const uint32_t vantage = ix[vp_pos].id;
        const A* vrow = pts_ + (size_t)vantage * stride_;
        auto dist_range = [&](size_t b, size_t e) {
            for (size_t r = b; r < e; ++r) ix[r].dist = fold_distance(pts_ + (size_t)ix[r].id * stride_, vrow, d_);
        };
        const size_t rest = len - 1;
        uint32_t nt = (uint32_t)std::min<size_t>(threads_ >> std::min(level, 31u), rest / 16384);
        if (nt > 1) {
            std::vector<std::thread> th;
            size_t chunk = (rest + nt - 1) / nt;
            for (uint32_t i = 0; i < nt; ++i) {
                size_t b = lo + i * chunk, e = std::min(lo + rest, b + chunk);
                if (b >= e) break;
                th.emplace_back([=] { dist_range(b, e); });
            }
            for (auto& x : th) x.join();
        } else {
            dist_range(lo, lo + rest);
        }
        // sort_unstable_by_key(|a| a.distance), :178; ties ordered by id (a legal outcome)
        std::sort(ix.begin() + lo, ix.begin() + lo + rest,
                  [](const DI& a, const DI& b) { return a.dist < b.dist || (a.dist == b.dist && a.id < b.id); });
        const size_t half = rest / 2;
        t.radii[node] = ix[lo + half].dist;  // far[0].distance, :182
        t.vp_ids[node] = vantage;
        t.vp_pos[node] = (uint32_t)vp_pos;
        std::memcpy(&t.centers[(size_t)node * t.dpad], vrow, d_ * sizeof(A));
        if ((1u << level) < threads_ && len > 32768) {
            std::thread th([&, node, lo, half, level] { recurse(ix, t, 2 * node + 1, lo, lo + half, level + 1); });
            recurse(ix, t, 2 * node + 2, lo + half, vp_pos, level + 1);
            th.join();
        } else {
            recurse(ix, t, 2 * node + 1, lo, lo + half, level + 1);
            recurse(ix, t, 2 * node + 2, lo + half, vp_pos, level + 1);
        }
    }

    const A* pts_;
    size_t n_, d_, stride_;
    uint32_t threads_;
};

}  // namespace petal
