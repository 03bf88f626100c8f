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
    for (uint32_t base = lo; base < hi; base += 32) {
        const uint32_t p = base + lane;
        float acc = pos_inf<float>();
        if (p < hi) {
            const float4* pr = t.pts + (size_t)p * t.dv;
            acc = 0.f;
            for (uint32_t j = 0; j < t.dv; ++j) acc = fold(acc, __ldg(qr + j), __ldg(pr + j));
        }
        unsigned mask = __ballot_sync(full, acc < kth);
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const float c = __shfl_sync(full, acc, src);
            const uint32_t pos = __popc(__ballot_sync(full, ks <= c));   // entries not larger than the candidate: a prefix
            if (pos >= k) continue;
            const float up = __shfl_up_sync(full, ks, 1);
            if ((uint32_t)lane > pos) ks = up;
            else if ((uint32_t)lane == pos) ks = c;
            kth = __shfl_sync(full, ks, (int)k - 1);
        }
    }
    if (lane == 0) seed_t2[i] = kth < pos_inf<float>() ? thresh2(xsqrt(kth)) : pos_inf<float>();
}

// One block per query group (QT = 32 * n_sub sorted queries): one bit per point tile.  Two levels: (1) the ball of each
// 32-query warp of the group against the tile ball -- one distance per (warp, tile); (2) only where that ball test cannot
// exclude the tile, the warp's 32 queries one by one (lane = query), each against its own seed; a tile confirmed by one
// query is not examined further.  With dense queries (1) decides almost everything; with sparse queries, whose warps span
// several clusters, (2) keeps the lists short.  bits[g * words + w] bit b <-> tile 32 w + b; cnt[g] = tiles to scan;
// total[0] += pairs the group will see.
__global__ void __launch_bounds__(256) tile_bitmap_kernel(const float4* __restrict__ qs, const float* __restrict__ seed_t2, uint32_t nq, uint32_t qt,
                                                          const float* __restrict__ tcen, const float* __restrict__ trad, uint32_t n_tiles,
                                                          uint32_t dv, float slack, uint32_t words, uint32_t* __restrict__ bits,
                                                          uint32_t* __restrict__ cnt, unsigned long long* __restrict__ total) {
    extern __shared__ float4 sm4[];            // [n_sub][dv] warp centres
    __shared__ float s_r[16], s_mu[16];        // warp radius; max over the warp's queries of (seed + distance to the warp centre)
    __shared__ float s_qtheta[512];            // every query's own seed as a distance (-inf: no such query)
    __shared__ uint32_t s_cnt;
    const uint32_t g = blockIdx.x, n_sub = qt / 32;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    if (threadIdx.x == 0) s_cnt = 0;
    for (uint32_t w = warp; w < n_sub; w += n_warps) {
        const uint32_t qi = g * qt + w * 32 + lane;
        const bool act = qi < nq;
        const uint32_t n_act = min(32u, nq > g * qt + w * 32 ? nq - (g * qt + w * 32) : 0u);
        const float4* qr = qs + (size_t)qi * dv;
        for (uint32_t j = 0; j < dv; ++j) {
            float4 v = act ? __ldg(qr + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int o = 16; o; o >>= 1) {
                v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
                v.z += __shfl_xor_sync(0xffffffffu, v.z, o); v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
            }
            const float inv = n_act ? 1.f / (float)n_act : 0.f;
            if (lane == 0) sm4[w * dv + j] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
        }
        __syncwarp();
        float acc = 0.f;
        if (act) for (uint32_t j = 0; j < dv; ++j) acc = fold(acc, __ldg(qr + j), sm4[w * dv + j]);
        float r = act ? xsqrt(acc) : 0.f;
        // the seed as a DISTANCE bound: sqrt of the squared-domain threshold, rounded up
        const float th = act ? xmul(xsqrt(seed_t2[qi]), 1.0000002f) : -pos_inf<float>();
        s_qtheta[w * 32 + lane] = th;
        // |q - p| >= |c_w - c_T| - |q - c_w| - R_T for every point p of tile T, so query q can skip T when
        // |c_w - c_T| - R_T - slack > seed_q + |q - c_w|; the warp needs T unless that holds for its largest seed_q + |q - c_w|
        float mu = act ? xmul(xadd(th, r), 1.0000002f) : -pos_inf<float>();
        for (int o = 16; o; o >>= 1) { r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o)); mu = fmaxf(mu, __shfl_xor_sync(0xffffffffu, mu, o)); }
        if (lane == 0) { s_r[w] = r; s_mu[w] = mu; }   // a warp without a live query has mu = -inf: it needs nothing
    }
    __syncthreads();
    uint32_t mine = 0, tried = 0, rejected = 0;
    for (uint32_t t0 = warp * 32; t0 < words * 32; t0 += n_warps * 32) {
        const uint32_t t = t0 + lane;
        uint32_t wmask = 0;  // warps of the group whose ball cannot exclude tile t
        if (t < n_tiles) {
            const float4* ct = reinterpret_cast<const float4*>(tcen) + (size_t)t * dv;
            const float rt = trad[t];
            float acc[16];
#pragma unroll
            for (int w = 0; w < 16; ++w) acc[w] = 0.f;
            for (uint32_t j = 0; j < dv; ++j) {
                const float4 c = __ldg(ct + j);
#pragma unroll
                for (int w = 0; w < 16; ++w)
                    if (w < (int)n_sub) acc[w] = fold(acc[w], sm4[w * dv + j], c);
            }
#pragma unroll
            for (int w = 0; w < 16; ++w) {
                if (w < (int)n_sub) {
                    const float cd = xsqrt(acc[w]);
                    const float sum = xadd(xadd(cd, s_r[w]), rt);
                    const float lb = xsub(xsub(cd, rt), xmul(slack, sum));
                    wmask |= (lb > s_mu[w]) ? 0u : 1u << w;
                }
            }
        }
        // level 2: lane = query of warp w, one candidate tile at a time (warp-uniform loop).  It only pays where it rejects
        // tiles: once 64 refinements of this warp have rejected fewer than a quarter, level 1's verdicts are taken as they are.
        unsigned needed = 0;  // lanes (tiles) confirmed
        for (uint32_t w = 0; w < n_sub; ++w) {
            unsigned cand = __ballot_sync(0xffffffffu, (wmask >> w) & 1u) & ~needed;
            if (!cand) continue;
            if (tried >= 64 && rejected * 4 < tried) { needed |= cand; continue; }
            const uint32_t qi = g * qt + w * 32 + lane;
            const float4* qr = qs + (size_t)min(qi, nq - 1) * dv;
            const float qth = s_qtheta[w * 32 + lane];
            while (cand) {
                const int src = __ffs(cand) - 1;
                cand &= cand - 1;
                const uint32_t tt = t0 + src;
                const float4* ct = reinterpret_cast<const float4*>(tcen) + (size_t)tt * dv;
                const float rt = trad[tt];
                float acc = 0.f;
                for (uint32_t j = 0; j < dv; ++j) acc = fold(acc, __ldg(qr + j), __ldg(ct + j));
                const float cd = xsqrt(acc);
                const float lb = xsub(xsub(cd, rt), xmul(slack, xadd(cd, rt)));
                ++tried;
                if (__ballot_sync(0xffffffffu, !(lb > qth))) needed |= 1u << src; else ++rejected;
            }
        }
        if (lane == 0) { bits[(size_t)g * words + (t0 >> 5)] = needed; mine += __popc(needed); }
    }
    if (lane == 0 && mine) atomicAdd(&s_cnt, mine);
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
