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
            // ranges of 0 or 1 points keep descending by the same mid rule, so that exactly one shard owns each point
            // (the others come out empty, hi == lo) even when n < 2^depth
            if (hi - lo >= 2) split(idx, lo, hi);
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
        // 1. partition, top-down: one min/max pass + one selection per level (nothing else touches the rows here)
        recurse(idx, t, 0, lo, hi, 0);
        // 2. flatten: points in bucket (idx) order, zero-padded to dpad
        t.pts.assign(n * (size_t)t.dpad, A(0));
        t.ids.resize(n);
        parallel_for(n, [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) {
                uint32_t src = idx[lo + i];
                t.ids[i] = src;
                std::memcpy(&t.pts[i * t.dpad], pts_ + (size_t)src * stride_, d_ * sizeof(A));
            }
        });
        // 3. + 4. centroids bottom-up from bucket sums, radii in one pass over the flattened rows
        node_balls(t);
        t.bucket_max = 0;
        for (uint32_t b = 0; b < t.n_buckets; ++b)
            t.bucket_max = std::max(t.bucket_max, t.bucket_hi[b] - t.bucket_lo[b]);
    }

  private:
    template <typename F> void parallel_for(size_t n, F f, size_t grain = 65536) {
        uint32_t nt = (uint32_t)std::min<size_t>(threads_, (n + grain - 1) / grain);
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

    // Node::init, src/ball_tree.rs:445-461, for every node of the flattened tree at once.
    //   centroid: a bucket's sum is the sequential sum of its rows in idx order (exactly the reference's Node::init
    //     for a single-bucket tree); an internal node's sum is left child + right child, so all centroids cost ONE pass
    //     over the points instead of one per level.  (The centroid of a large node then differs from the strictly
    //     sequential sum in the last bits -- it is in fact more accurate; any centre with radius = max exact fold
    //     distance to it is a valid ball, so results are unaffected.)
    //   radius: max over the node's points of Euclidean::distance(centroid, point), src/distance.rs:26-35.  One pass
    //     over the rows: each row is folded against the centroids of its L+1 ancestors at once -- the fold over the
    //     dimensions stays sequential per ancestor (bit-exact), the loop over the ancestors vectorises.
    void node_balls(FlatTree<A>& t) {
        const size_t dp = t.dpad, nb = t.n_buckets, L = t.L;
        std::vector<A> sum((size_t)t.n_nodes * dp, A(0));
        std::vector<uint64_t> cnt(t.n_nodes, 0);
        parallel_for(nb, [&](size_t b0, size_t b1) {
            for (size_t b = b0; b < b1; ++b) {
                A* acc = &sum[(t.n_internal + b) * dp];
                for (size_t i = t.bucket_lo[b]; i < t.bucket_hi[b]; ++i) {
                    const A* r = &t.pts[i * dp];
                    for (size_t j = 0; j < d_; ++j) acc[j] += r[j];
                }
                cnt[t.n_internal + b] = t.bucket_hi[b] - t.bucket_lo[b];
            }
        }, 1);
        for (size_t node = t.n_internal; node-- > 0;) {  // children before parents
            const A* l = &sum[(2 * node + 1) * dp];
            const A* r = &sum[(2 * node + 2) * dp];
            A* o = &sum[node * dp];
            for (size_t j = 0; j < d_; ++j) o[j] = l[j] + r[j];
            cnt[node] = cnt[2 * node + 1] + cnt[2 * node + 2];
        }
        for (size_t node = 0; node < t.n_nodes; ++node) {
            A* c = &t.centers[node * dp];
            if (cnt[node] == 0) { t.radii[node] = A(-1); continue; }
            const A flen = (A)cnt[node];
            for (size_t j = 0; j < d_; ++j) c[j] = sum[node * dp + j] / flen;
            t.radii[node] = A(0);
        }
        // radii: per worker a private max per node, merged afterwards (max is order independent)
        const size_t LV = L + 1;
        const uint32_t nw = std::max<uint32_t>(1, (uint32_t)std::min<size_t>(threads_, (t.n * d_ >> 20) + 1));
        std::vector<std::vector<A>> wmax(nw, std::vector<A>(t.n_nodes, A(0)));
        auto work = [&](uint32_t w) {
            std::vector<A> ct(d_ * LV), acc(LV);  // the ancestors' centroids, transposed: ct[j][level]
            std::vector<uint32_t> anc(LV);
            A* mx = wmax[w].data();
            const size_t b0 = nb * w / nw, b1 = nb * (w + 1) / nw;
            for (size_t b = b0; b < b1; ++b) {
                if (t.bucket_hi[b] == t.bucket_lo[b]) continue;
                size_t h = t.n_internal + b;  // heap index of the bucket node; its ancestor `up` levels above is ((h+1) >> up) - 1
                for (size_t up = 0; up < LV; ++up) {
                    anc[up] = (uint32_t)(((h + 1) >> up) - 1);
                    const A* c = &t.centers[(size_t)anc[up] * dp];
                    for (size_t j = 0; j < d_; ++j) ct[j * LV + up] = c[j];
                }
                for (size_t i = t.bucket_lo[b]; i < t.bucket_hi[b]; ++i) {
                    const A* r = &t.pts[i * dp];
                    for (size_t up = 0; up < LV; ++up) acc[up] = A(0);
                    for (size_t j = 0; j < d_; ++j) {
                        const A v = r[j];
                        const A* cj = &ct[j * LV];
                        for (size_t up = 0; up < LV; ++up) {
                            const A diff = cj[up] - v;   // x1 = centroid, x2 = point, as in Node::init
                            acc[up] += diff * diff;
                        }
                    }
                    for (size_t up = 0; up < LV; ++up) {
                        const A dist = std::sqrt(acc[up]);
                        if (dist > mx[anc[up]]) mx[anc[up]] = dist;
                    }
                }
            }
        };
        if (nw == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (uint32_t w = 0; w < nw; ++w) th.emplace_back(work, w);
            for (auto& x : th) x.join();
        }
        for (size_t node = 0; node < t.n_nodes; ++node) {
            if (cnt[node] == 0) continue;
            A m = A(0);
            for (uint32_t w = 0; w < nw; ++w) if (wmax[w][node] > m) m = wmax[w][node];
            t.radii[node] = m;
        }
    }

    void recurse(std::vector<uint32_t>& idx, FlatTree<A>& t, uint32_t node, size_t lo, size_t hi, uint32_t level) {
        if (level == t.L) {
            uint32_t b = node - t.n_internal;
            t.bucket_lo[b] = (uint32_t)(lo - base_);
            t.bucket_hi[b] = (uint32_t)(hi - base_);
            // canonical storage order of a bucket: ascending original index (the selection leaves an arbitrary order);
            // the device builder (gpu_build.cu) produces the same, so the two flattened layouts are bit-identical
            std::sort(idx.begin() + lo, idx.begin() + hi);
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
