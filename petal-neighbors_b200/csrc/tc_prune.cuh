// tc_prune.cuh -- triangle-inequality pruning in front of the tensor filter (north-star stages 1 -> 3).
//
// The reference prunes a node when its lower bound exceeds the query's current k-th distance and visits the nearer child
// first (src/ball_tree.rs:211-214, 230-238).  The tensor filter streams 128-point tiles of the bucket-ordered point image
// against CTAs of 256 / 512 queries, so the unit that can be skipped is (query group, point tile):
//   * every 128-row tile of the stored order gets its own ball (centre = mean of its rows, radius = largest exact fold
//     distance to it): rows are stored bucket by bucket, so a tile is a spatially compact half bucket;
//   * the queries are sorted by home bucket (home_bucket_kernel), so a CTA's queries are neighbours, and every warp of 32
//     consecutive sorted queries gets a ball of its own;
//   * each query gets a SEED bound before the scan: its exact k-th distance among the points of its home bucket -- the
//     "nearer child first" of the reference turned into a first pass;
//   * a tile is scanned by a CTA iff for some 32-query warp w   |c_w - c_T| - R_w - R_T - slack > max seed of w  fails,
//     i.e. iff some query of the CTA may still find a neighbour in it.  One bit per (CTA, tile).
// Exactness: the seed is the k-th smallest of a SUBSET of the exact distances, hence >= the final k-th distance; the bound
// uses the same conservative slack as the SIMT traversal ((2d+8) u on the sum of the three computed lengths, two triangle
// steps), so a skipped tile holds no point that the brute-force (distance, index) order would keep.
#pragma once
#include "kernels.cuh"

namespace petal {
namespace tc {

constexpr int PR_TILE = 128;  // = BN of tc_filter.cuh

// one block per tile: centre = mean of the tile's rows, radius = max Euclidean::distance(centre, row)
__global__ void __launch_bounds__(128) tile_balls_kernel(const float* __restrict__ pts, uint32_t n, uint32_t d, uint32_t dpad,
                                                         float* __restrict__ centers, float* __restrict__ radii) {
    const uint32_t t = blockIdx.x;
    const uint32_t lo = t * PR_TILE, hi = min(n, lo + PR_TILE);
    extern __shared__ __align__(16) float sc[];  // dpad centre coordinates
    __shared__ float wmax[4];
    for (uint32_t j = threadIdx.x; j < dpad; j += blockDim.x) {
        float acc = 0.f;
        if (j < d) {
            for (uint32_t i = lo; i < hi; ++i) acc += pts[(size_t)i * dpad + j];
            acc = acc / (float)(hi - lo);
        }
        sc[j] = acc;
        centers[(size_t)t * dpad + j] = acc;
    }
    __syncthreads();
    float r = 0.f;
    const uint32_t i = lo + threadIdx.x;
    if (i < hi) {
        float acc = 0.f;
        for (uint32_t j = 0; j < dpad; j += 4) {
            const float4 c = *reinterpret_cast<const float4*>(sc + j);
            const float4 p = *reinterpret_cast<const float4*>(pts + (size_t)i * dpad + j);
            acc = fold(acc, c, p);
        }
        r = xsqrt(acc);
    }
    for (int o = 16; o; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = r;
    __syncthreads();
    if (threadIdx.x == 0) radii[t] = fmaxf(fmaxf(wmax[0], wmax[1]), fmaxf(wmax[2], wmax[3]));
}

// Build-time estimate of what pruning can do on this point set: for a sample of tiles taken as stand-ins for a query
// group (a query inside tile s has k neighbours within ~2 R_s of itself), the fraction of tiles T whose ball is out of
// reach, |c_s - c_T| - R_s - R_T > 2 R_s.  Uniform data in d >= 16 gives ~0, well-separated clusters ~1.
__global__ void prune_estimate_kernel(const float* __restrict__ centers, const float* __restrict__ radii, uint32_t n_tiles, uint32_t dpad,
                                      uint32_t n_samples, unsigned long long* __restrict__ out /* [0] prunable, [1] pairs */) {
    const uint32_t s = (uint32_t)(((unsigned long long)blockIdx.x * n_tiles) / n_samples);
    const float4* cs = reinterpret_cast<const float4*>(centers + (size_t)s * dpad);
    const float rs = radii[s];
    unsigned long long cnt = 0, tot = 0;
    for (uint32_t t = threadIdx.x; t < n_tiles; t += blockDim.x) {
        const float4* ct = reinterpret_cast<const float4*>(centers + (size_t)t * dpad);
        float acc = 0.f;
        for (uint32_t j = 0; j < dpad / 4; ++j) acc = fold(acc, __ldg(cs + j), __ldg(ct + j));
        cnt += (xsqrt(acc) - rs - radii[t] > 2.f * rs) ? 1 : 0;
        ++tot;
    }
    atomicAdd(&out[0], cnt);
    atomicAdd(&out[1], tot);
}

// Build-time estimate of what SEEDING can do: sample stored points stand in for queries; the seed is the distance of the
// 11-th nearest point of the own bucket (the sample itself is one of them); counted is the share of a fixed sample of
// all points that lies within the seed.  share x n = candidates a query would still hand to the exact rerank when it
// starts from its seed: a few hundred on clustered data, tens of thousands on uniform data in d >= 16 (where the running
// threshold of the stream is just as good and the set-up passes are not worth it).
// The same sample also measures what TILE skipping can do for a single query: the share of tile balls that lie beyond the
// seed, |q - c_T| - R_T > seed (out[4] of out[5]).  A CTA needs the union over its queries, so this is an upper bound of
// what the bitmaps will skip.
__global__ void __launch_bounds__(256) seed_estimate_kernel(const DevTree<float> t, const uint32_t* __restrict__ sample_row,
                                                            const uint32_t* __restrict__ sample_bucket, uint32_t m_points,
                                                            const float* __restrict__ tcen, const float* __restrict__ trad, uint32_t n_tiles,
                                                            unsigned long long* __restrict__ out /* [2] within, [3] examined, [4] tiles out of reach, [5] tiles */) {
    __shared__ float sd[1024];
    __shared__ float s_seed;
    __shared__ unsigned int s_cnt;
    const uint32_t row = sample_row[blockIdx.x], b = sample_bucket[blockIdx.x];
    const uint32_t lo = t.bucket_lo[b], m = min(t.bucket_hi[b] - lo, 1024u);
    const float4* qr = t.pts + (size_t)row * t.dv;
    if (threadIdx.x == 0) { s_seed = pos_inf<float>(); s_cnt = 0; }
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        const float4* pr = t.pts + (size_t)(lo + i) * t.dv;
        float acc = 0.f;
        for (uint32_t j = 0; j < t.dv; ++j) acc = fold(acc, __ldg(qr + j), __ldg(pr + j));
        sd[i] = acc;
    }
    __syncthreads();
    // the squared distance of rank 10 (0-based) among the bucket's points: the smallest value with at least 10 smaller ones
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        uint32_t rank = 0;
        for (uint32_t j = 0; j < m; ++j) rank += sd[j] < sd[i] ? 1u : 0u;
        if (rank >= 10) atomicMin(reinterpret_cast<unsigned int*>(&s_seed), __float_as_uint(sd[i]));
    }
    __syncthreads();
    const float seed = s_seed;  // +inf when the bucket has fewer than 11 points
    unsigned int c = 0;
    for (uint32_t i = threadIdx.x; i < m_points; i += blockDim.x) {
        const uint32_t p = (uint32_t)(((unsigned long long)i * t.n) / m_points);
        const float4* pr = t.pts + (size_t)p * t.dv;
        float acc = 0.f;
        for (uint32_t j = 0; j < t.dv; ++j) acc = fold(acc, __ldg(qr + j), __ldg(pr + j));
        c += acc <= seed ? 1u : 0u;
    }
    atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) { atomicAdd(&out[2], (unsigned long long)s_cnt); atomicAdd(&out[3], (unsigned long long)m_points); s_cnt = 0; }
    __syncthreads();
    const float seed_d = xmul(xsqrt(seed), 1.0000002f);
    c = 0;
    for (uint32_t tt = threadIdx.x; tt < n_tiles; tt += blockDim.x) {
        const float4* ct = reinterpret_cast<const float4*>(tcen) + (size_t)tt * t.dv;
        float acc = 0.f;
        for (uint32_t j = 0; j < t.dv; ++j) acc = fold(acc, __ldg(qr + j), __ldg(ct + j));
        const float cd = xsqrt(acc), rt = trad[tt];
        c += (xsub(xsub(cd, rt), xmul(t.slack, xadd(cd, rt))) > seed_d) ? 1u : 0u;
    }
    atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) { atomicAdd(&out[4], (unsigned long long)s_cnt); atomicAdd(&out[5], (unsigned long long)n_tiles); }
}

// every tile for every group (seeding without tile pruning)
__global__ void fill_bitmap_kernel(uint32_t n_groups, uint32_t n_tiles, uint32_t words, uint32_t* __restrict__ bits, uint32_t* __restrict__ cnt) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n_groups * words) return;
    const uint32_t w = (uint32_t)(e % words);
    const uint32_t rest = n_tiles - w * 32;
    bits[e] = rest >= 32 ? 0xffffffffu : ((1u << rest) - 1u);
    if (w == 0) cnt[e / words] = n_tiles;
}

// sorted query rows: out[i] = q[order[i]] (padded rows), so that CTA x of the filter serves sorted slots [x QT, (x+1) QT)
__global__ void gather_queries_kernel(const float4* __restrict__ q, const uint32_t* __restrict__ order, uint32_t nq, uint32_t dv,
                                      float4* __restrict__ out) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)nq * dv) return;
    const uint32_t i = (uint32_t)(e / dv), j = (uint32_t)(e % dv);
    out[e] = q[(size_t)order[i] * dv + j];
}

// Seed bound: exact k-th distance of every (sorted) query among the points of its home bucket, as thresh2(kth) in the
// squared domain (+inf when the bucket holds fewer than k points).  One WARP per query: the lanes fold 32 points of the
// bucket at a time and the k smallest squared sums live, sorted, across the lanes (lane i = i-th smallest; one ballot
// finds a candidate's position, one shfl_up shifts the tail).  sqrt_rn is monotone, so the k-th smallest sqrt'd distance
// is the sqrt of the k-th smallest sum: the same value a sequential pass over the bucket finds.  (A thread per query
// -- a 244-point x dv dependent chain per thread -- took 17.8 ms per million queries at d = 64; the prepass was 10 % of
// config 3.)
__global__ void __launch_bounds__(256) seed_bound_kernel(const DevTree<float> t, const float4* __restrict__ qs, const uint32_t* __restrict__ order,
                                                         const uint32_t* __restrict__ home, uint32_t nq, uint32_t k, float* __restrict__ seed_t2) {
    const int lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= nq) return;
    const unsigned full = 0xffffffffu;
    const uint32_t b = home[order ? order[i] : i];
    const uint32_t lo = t.bucket_lo[b], hi = t.bucket_hi[b];
    const float4* qr = qs + (size_t)i * t.dv;
    float ks = pos_inf<float>();    // this lane's entry of the sorted list
    float kth = pos_inf<float>();   // entry k-1 (warp-uniform)
    // NP points per lane and round: NP independent fold chains keep NP L2 round trips in flight (one chain per lane left the
    // kernel latency-bound: 16 % of the issue slots at 92 % occupancy, profiles/r02_ncu_full_pruned_scan_c3p.md)
    constexpr int NP = 4;
    for (uint32_t base = lo; base < hi; base += 32 * NP) {
        float acc[NP];
        const float4* pr[NP];
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            const uint32_t p = base + u * 32 + lane;
            pr[u] = t.pts + (size_t)min(p, hi - 1) * t.dv;
            acc[u] = 0.f;
        }
        for (uint32_t j = 0; j < t.dv; ++j) {
            const float4 qv = __ldg(qr + j);
#pragma unroll
            for (int u = 0; u < NP; ++u) acc[u] = fold(acc[u], qv, __ldg(pr[u] + j));
        }
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            const float cand = base + u * 32 + lane < hi ? acc[u] : pos_inf<float>();
            unsigned mask = __ballot_sync(full, cand < kth);
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                const float c = __shfl_sync(full, cand, src);
                const uint32_t pos = __popc(__ballot_sync(full, ks <= c));   // entries not larger than the candidate: a prefix
                if (pos >= k) continue;
                const float up = __shfl_up_sync(full, ks, 1);
                if ((uint32_t)lane > pos) ks = up;
                else if ((uint32_t)lane == pos) ks = c;
                kth = __shfl_sync(full, ks, (int)k - 1);
            }
        }
    }
    if (lane == 0) seed_t2[i] = kth < pos_inf<float>() ? thresh2(xsqrt(kth)) : pos_inf<float>();
}

// ---- tile bitmaps: one bit per (query group of a filter CTA, 128-point tile) -----------------------------------------------
// A group is QT = 32 n_sub sorted queries.  The unit of the test is a BALL of queries (centre c, radius R, and
// mu = max over its queries of seed_q + |q - c|): every point p of tile T has |q - p| >= |c - c_T| - |q - c| - R_T, so no
// query of the ball needs T when  |c - c_T| - R_T - slack (|c - c_T| + R + R_T) > mu.
//   1. warp_balls_kernel   : TWO balls for every 32 consecutive sorted queries (the warp split around two far-apart queries)
//   2. ball_tile_kernel    : every (ball, tile) pair -- 32 balls in shared memory per block, a thread per tile with 32
//                            accumulators, one ballot per ball gives 32 tile bits at a time -> a bitmap and a count per ball
//   3. wide_select_kernel  : a warp whose count is more than twice the smallest count of its group (queries of three
//                            clusters in one warp ...) is "wide": at most a quarter of the warps
//   4. ball_tile_kernel<PER_QUERY>: the 32 queries of a wide warp as 32 balls of radius 0, each against its own seed; the
//                            OR of the 32 answers replaces the warp's two bitmaps
//   5. group_bits_kernel   : OR over the balls of a group, tiles to scan per group, pairs the scan will see
// (r2, first version: one kernel per group that refined ball verdicts query by query, re-reading the 32 query rows for
// every candidate tile -- 180 ms per million queries on the two-means partition of BASELINE config 3, where the refinement
// does reject tiles and therefore never gave up; the filter it fed took 50 ms.)
constexpr uint32_t PR_NONE = 0xffffffffu;

// Two balls per warp of 32 sorted queries (balls 2w and 2w + 1).  Queries are sorted by home bucket only, and a bucket that
// holds fragments of two clusters sends queries of both into one warp: one ball around them all would reach every tile.
// The warp is therefore split around two far-apart queries (a: the query farthest from the warp's mean, b: the query
// farthest from a; every query joins the nearer one), and each half gets its own centre, radius and bound.  A half
// without queries has mu = -inf and needs nothing.
__global__ void __launch_bounds__(256) warp_balls_kernel(const float4* __restrict__ qs, const float* __restrict__ seed_t2, uint32_t nq, uint32_t n_warps,
                                                         uint32_t dv, float4* __restrict__ bcen, float* __restrict__ brad, float* __restrict__ bmu,
                                                         float* __restrict__ qtheta) {
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31, w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_warps) return;
    const uint32_t q0 = w * 32, qi = q0 + lane;
    const bool act = qi < nq;
    const uint32_t n_act = q0 < nq ? min(32u, nq - q0) : 0u;
    const float4* qr = qs + (size_t)(act ? qi : 0) * dv;
    // the seed as a DISTANCE bound: sqrt of the squared-domain threshold, rounded up
    const float th = act ? xmul(xsqrt(seed_t2[qi]), 1.0000002f) : -pos_inf<float>();
    qtheta[qi] = th;   // (the array is padded to whole warps)
    float4* c0 = bcen + (size_t)(2 * w) * dv;
    float4* c1 = c0 + dv;
    auto sum4 = [&](float4 v) {
        for (int o = 16; o; o >>= 1) {
            v.x += __shfl_xor_sync(full, v.x, o); v.y += __shfl_xor_sync(full, v.y, o);
            v.z += __shfl_xor_sync(full, v.z, o); v.w += __shfl_xor_sync(full, v.w, o);
        }
        return v;
    };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // (1) the warp's mean, parked in ball 0's centre
    const float inv = n_act ? 1.f / (float)n_act : 0.f;
    for (uint32_t j = 0; j < dv; ++j) {
        const float4 v = sum4(act ? __ldg(qr + j) : zero4);
        if (lane == 0) c0[j] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
    }
    __syncwarp();
    float acc = 0.f;
    if (act) for (uint32_t j = 0; j < dv; ++j) acc = fold(acc, __ldg(qr + j), c0[j]);
    // (2) a = the query farthest from the mean, b = the query farthest from a (lowest lane on ties)
    float far = act ? acc : -1.f, best = far;
    for (int o = 16; o; o >>= 1) best = fmaxf(best, __shfl_xor_sync(full, best, o));
    const int la = __ffs(__ballot_sync(full, far == best)) - 1;
    float da = 0.f;
    for (uint32_t j = 0; j < dv; ++j) {
        const float4 mine = act ? __ldg(qr + j) : zero4;
        float4 a;
        a.x = __shfl_sync(full, mine.x, la); a.y = __shfl_sync(full, mine.y, la); a.z = __shfl_sync(full, mine.z, la); a.w = __shfl_sync(full, mine.w, la);
        da = fold(da, mine, a);
    }
    far = act ? da : -1.f; best = far;
    for (int o = 16; o; o >>= 1) best = fmaxf(best, __shfl_xor_sync(full, best, o));
    const int lb = __ffs(__ballot_sync(full, far == best)) - 1;
    float db = 0.f;
    for (uint32_t j = 0; j < dv; ++j) {
        const float4 mine = act ? __ldg(qr + j) : zero4;
        float4 b;
        b.x = __shfl_sync(full, mine.x, lb); b.y = __shfl_sync(full, mine.y, lb); b.z = __shfl_sync(full, mine.z, lb); b.w = __shfl_sync(full, mine.w, lb);
        db = fold(db, mine, b);
    }
    const bool g1 = act && db < da;          // nearer b: second half; everything else that is live: first half
    const bool g0 = act && !g1;
    const uint32_t n1 = __popc(__ballot_sync(full, g1)), n0 = n_act - n1;
    const float inv0 = n0 ? 1.f / (float)n0 : 0.f, inv1 = n1 ? 1.f / (float)n1 : 0.f;
    __syncwarp();
    // (3) the centres of the two halves
    for (uint32_t j = 0; j < dv; ++j) {
        const float4 mine = act ? __ldg(qr + j) : zero4;
        const float4 s0 = sum4(g0 ? mine : zero4), s1 = sum4(g1 ? mine : zero4);
        if (lane == 0) {
            c0[j] = make_float4(s0.x * inv0, s0.y * inv0, s0.z * inv0, s0.w * inv0);
            c1[j] = make_float4(s1.x * inv1, s1.y * inv1, s1.z * inv1, s1.w * inv1);
        }
    }
    __syncwarp();
    // (4) radius and bound of each half: |q - p| >= |c - c_T| - |q - c| - R_T for every point p of tile T
    const float4* cm = g1 ? c1 : c0;
    acc = 0.f;
    if (act) for (uint32_t j = 0; j < dv; ++j) acc = fold(acc, __ldg(qr + j), cm[j]);
    const float r = act ? xsqrt(acc) : 0.f;
    const float mu = act ? xmul(xadd(th, r), 1.0000002f) : -pos_inf<float>();
    float r0 = g0 ? r : 0.f, r1 = g1 ? r : 0.f, m0 = g0 ? mu : -pos_inf<float>(), m1 = g1 ? mu : -pos_inf<float>();
    for (int o = 16; o; o >>= 1) {
        r0 = fmaxf(r0, __shfl_xor_sync(full, r0, o)); r1 = fmaxf(r1, __shfl_xor_sync(full, r1, o));
        m0 = fmaxf(m0, __shfl_xor_sync(full, m0, o)); m1 = fmaxf(m1, __shfl_xor_sync(full, m1, o));
    }
    if (lane == 0) { brad[2 * w] = r0; brad[2 * w + 1] = r1; bmu[2 * w] = m0; bmu[2 * w + 1] = m1; }
}

// PER_QUERY = false: block b tests balls [32 b, 32 b + 32) (centres bcen, radii brad, bounds bmu); out_bits[ball][word],
//                    out_cnt[ball] = tiles the ball needs.
// PER_QUERY = true : block b tests the 32 queries of wide ball wlist[b] (centres = the query rows, radius 0, bound =
//                    qtheta); out_bits[b][word] = OR over the 32 queries.
template <bool PER_QUERY>
__global__ void __launch_bounds__(256) ball_tile_kernel(const float4* __restrict__ cen, const float* __restrict__ brad, const float* __restrict__ bmu,
                                                        uint32_t n_balls, const uint32_t* __restrict__ wlist, const uint32_t* __restrict__ n_wide,
                                                        uint32_t wide_cap, const float* __restrict__ tcen, const float* __restrict__ trad,
                                                        uint32_t n_tiles, uint32_t dv, float slack, uint32_t words, uint32_t* __restrict__ out_bits,
                                                        uint32_t* __restrict__ out_cnt) {
    extern __shared__ float4 sm4[];   // [32][dv] ball centres
    __shared__ float s_r[32], s_mu[32];
    __shared__ uint32_t s_cnt[32];
    uint32_t ball0;
    if (PER_QUERY) {
        if (blockIdx.x >= min(*n_wide, wide_cap)) return;
        ball0 = wlist[blockIdx.x] * 32;   // first query of the wide ball: centres are query rows
    } else {
        ball0 = blockIdx.x * 32;
    }
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (uint32_t e = threadIdx.x; e < 32 * dv; e += blockDim.x) {
        const uint32_t b = e / dv;
        sm4[e] = (PER_QUERY || ball0 + b < n_balls) ? __ldg(cen + (size_t)(ball0 + b) * dv + (e % dv)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (threadIdx.x < 32) {
        const uint32_t b = ball0 + threadIdx.x;
        s_r[threadIdx.x] = PER_QUERY ? 0.f : (b < n_balls ? brad[b] : 0.f);
        s_mu[threadIdx.x] = (PER_QUERY || b < n_balls) ? bmu[b] : -pos_inf<float>();
        s_cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    uint32_t mine = 0;   // tiles ball `lane` needs, over this warp's share of the tiles
    for (uint32_t t0 = warp * 32; t0 < words * 32; t0 += n_warps * 32) {
        const uint32_t t = t0 + lane;
        const bool live = t < n_tiles;
        float acc[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) acc[b] = 0.f;
        if (live) {
            const float4* ct = reinterpret_cast<const float4*>(tcen) + (size_t)t * dv;
            for (uint32_t j = 0; j < dv; ++j) {
                const float4 c = __ldg(ct + j);
#pragma unroll
                for (int b = 0; b < 32; ++b) acc[b] = fold(acc[b], sm4[b * dv + j], c);
            }
        }
        const float rt = live ? trad[t] : 0.f;
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const float cd = xsqrt(acc[b]);
            const float sum = xadd(xadd(cd, s_r[b]), rt);
            const float lb = xsub(xsub(cd, rt), xmul(slack, sum));
            const unsigned need = __ballot_sync(0xffffffffu, live && !(lb > s_mu[b]));
            if (PER_QUERY) word |= need;
            else if ((int)lane == b) word = need;
        }
        if (PER_QUERY) {
            if (lane == 0) out_bits[(size_t)blockIdx.x * words + (t0 >> 5)] = word;
        } else {
            if (ball0 + lane < n_balls) out_bits[(size_t)(ball0 + lane) * words + (t0 >> 5)] = word;
            mine += __popc(word);
        }
    }
    if (!PER_QUERY) {
        if (mine) atomicAdd(&s_cnt[lane], mine);
        __syncthreads();
        if (threadIdx.x < 32 && ball0 + threadIdx.x < n_balls) out_cnt[ball0 + threadIdx.x] = s_cnt[threadIdx.x];
    }
}

// one thread per group: its wide warps get a slot in the refinement list (wslot[warp], PR_NONE otherwise).  A warp counts
// with the larger of its two balls' tile counts.
__global__ void wide_select_kernel(const uint32_t* __restrict__ ball_cnt, const float* __restrict__ bmu, uint32_t n_groups, uint32_t n_sub,
                                   uint32_t wide_cap, uint32_t* __restrict__ wslot, uint32_t* __restrict__ wlist, uint32_t* __restrict__ n_wide) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    auto live = [&](uint32_t w) { return bmu[2 * w] > -pos_inf<float>() || bmu[2 * w + 1] > -pos_inf<float>(); };
    auto need = [&](uint32_t w) { return max(ball_cnt[2 * w], ball_cnt[2 * w + 1]); };
    uint32_t lo = PR_NONE;
    for (uint32_t i = 0; i < n_sub; ++i)
        if (live(g * n_sub + i)) lo = min(lo, need(g * n_sub + i));
    for (uint32_t i = 0; i < n_sub; ++i) {
        const uint32_t w = g * n_sub + i;
        uint32_t slot = PR_NONE;
        if (lo != PR_NONE && live(w) && need(w) > 2u * lo + 16u) {
            slot = atomicAdd(n_wide, 1u);
            if (slot < wide_cap) wlist[slot] = w; else slot = PR_NONE;
        }
        wslot[w] = slot;
    }
}

// one block per group: OR over its warps (both balls of a warp, or the refined bitmap of a wide warp), cnt[g] = tiles to
// scan, total[0] += pairs the group will see
__global__ void __launch_bounds__(256) group_bits_kernel(const uint32_t* __restrict__ ball_bits, const uint32_t* __restrict__ wide_bits,
                                                         const uint32_t* __restrict__ wslot, uint32_t nq, uint32_t qt, uint32_t n_sub, uint32_t words,
                                                         uint32_t* __restrict__ bits, uint32_t* __restrict__ cnt, unsigned long long* __restrict__ total) {
    __shared__ uint32_t s_cnt;
    __shared__ uint32_t s_slot[16];
    const uint32_t g = blockIdx.x;
    if (threadIdx.x == 0) s_cnt = 0;
    if (threadIdx.x < n_sub) s_slot[threadIdx.x] = wslot[g * n_sub + threadIdx.x];
    __syncthreads();
    uint32_t mine = 0;
    for (uint32_t wd = threadIdx.x; wd < words; wd += blockDim.x) {
        uint32_t v = 0;
        for (uint32_t i = 0; i < n_sub; ++i) {
            const uint32_t slot = s_slot[i];
            const size_t w = (size_t)g * n_sub + i;
            v |= slot != PR_NONE ? wide_bits[(size_t)slot * words + wd] : (ball_bits[(2 * w) * words + wd] | ball_bits[(2 * w + 1) * words + wd]);
        }
        bits[(size_t)g * words + wd] = v;
        mine += __popc(v);
    }
    if (mine) atomicAdd(&s_cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0) {
        cnt[g] = s_cnt;
        const uint32_t live = min(qt, nq - g * qt);
        atomicAdd(total, (unsigned long long)s_cnt * PR_TILE * live);
    }
}

// the tile sequence of a group, read from its bitmap; every role of the filter CTA walks it on its own
struct TileIter {
    const uint32_t* w;
    uint32_t wi, cur;
    __device__ __forceinline__ void init(const uint32_t* words) { w = words; wi = 0; cur = __ldg(w); }
    __device__ __forceinline__ uint32_t next() {  // callers take exactly cnt[g] tiles, so a set bit always exists
        while (!cur) cur = __ldg(w + ++wi);
        const uint32_t b = __ffs(cur) - 1;
        cur &= cur - 1;
        return wi * 32 + b;
    }
};

}  // namespace tc
}  // namespace petal
