// tc_filter.cuh -- tensor-core k-NN scan: tcgen05 TF32 filter + exact difference-form rerank.
//
// Where the query x bucket tile really is a dense contraction (f32, d >= 16: the ball bounds
// prune nothing, pairs/(N*Q) ~ 1), the distance matrix is evaluated on the 5th-gen tensor cores as
//        D~^2(q,p) = |q'|^2 + |p'|^2 - 2 q'.p'          (q' = q - c, p' = p - c, c = data mean)
// by ONE augmented TF32 contraction: A row = [q'_0 .. q'_{d-1}, 0.., hi(|q'|^2), lo(|q'|^2), 1, 1],
// B row = [-2p'_0 .. -2p'_{d-1}, 0.., 1, 1, hi(|p'|^2), lo(|p'|^2)], so the TMEM accumulator holds
// D~^2 itself and the epilogue is a bare threshold test.  The filter is only a filter: every
// element with D~^2 <= Theta_q is re-evaluated with the exact sequential non-FMA fold of
// kernels.cuh (Euclidean::distance, reference src/distance.rs:26-35) and selected on the
// (sqrt'd distance, index) key, so results are bit-identical to the SIMT path and the oracle.
//
// Theta_q = thresh2(kth_q) * (1 + (d+4) 2^-23) + E_q with the rigorous TF32 bound
//   E_q = 1.01 * 2^-8 |q'| Pmax + (Kp + 8) 2^-21 (|q'| + Pmax)^2,   Pmax = max_p |p'|
// (input truncation to TF32: 2^-10 relative per factor on the -2q'.p' terms; hi/lo split norms:
// 2^-20; fp32 accumulation, norm evaluation and centring: the quadratic term).
//
// Structure (one CTA = MT x 128 queries, persistent over all point tiles; 1 CTA / SM):
//   warp 4MT   : TMA producer, cp.async.bulk.tensor 2-D boxes [128 rows x 32 tf32] (128B swizzle)
//   warp 4MT+1 : tcgen05.mma issuer (one elected lane), kind::tf32, M=128 N=128 K=8 per instruction,
//                accumulators double-buffered in TMEM (2 x MT x 128 columns)
//   warps 0..4MT-1 : epilogue, one query row per thread: tcgen05.ld 32x32b.x32 -> min tree ->
//                    threshold -> (rare) exact rerank + sorted insertion into the thread's top-k
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace petal {
namespace tc {

constexpr int BM = 128;            // queries per accumulator tile (TMEM lanes)
constexpr int BN = 128;            // points per B tile (TMEM columns per accumulator)
constexpr int KC = 32;             // tf32 elements per K chunk = one 128-byte swizzle row
constexpr int CHUNK_BYTES = BN * KC * 4;  // 16 KB
constexpr int NUM_ACC = 2;         // accumulator stages in TMEM

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32, both operands K-major
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 = 1 [16,30) | SBO>>4 = 1024/16 [32,46) | version 1 [46,48) | SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32 instruction descriptor: D=F32 [4,6)=1, A=B=TF32 [7,10)=[10,13)=2, K-major both, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct FilterArgs {
    DevTree<float> t;
    const float4* q;       // nq x dpad exact zero-padded queries (rerank)
    const float* q_margin; // nq: E_q
    uint32_t nq, k;
    uint32_t n_tiles;      // ceil(n / BN)
    uint32_t nkc;          // K chunks (Kp / 32)
    uint32_t stages;       // B ring depth
    float t2_scale;        // 1 + (d+4) 2^-23
    float* part_d;         // [nq][k]
    uint32_t* part_i;
    const float* floor_d;
    const uint32_t* floor_i;
    unsigned long long* counters;  // [2] filter hits (elements passed to the exact rerank)
};

// hi/lo split of a non-negative fp32 value into two TF32-exact pieces (hi has 10 mantissa bits)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

// B operand: one row per stored point (bucket order).  pmax_bits receives max |p'| (as float bits).
__global__ void build_baug_kernel(const float* __restrict__ pts, const float* __restrict__ center, uint32_t n, uint32_t d,
                                  uint32_t dpad, uint32_t kp, float* __restrict__ baug, unsigned int* pmax_bits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = pts + (size_t)i * dpad;
    float* o = baug + (size_t)i * kp;
    float nrm = 0.f;
    for (uint32_t j = 0; j < d; ++j) {
        const float v = p[j] - center[j];
        nrm = nrm + v * v;
        o[j] = -2.0f * v;
    }
    for (uint32_t j = d; j < kp - 4; ++j) o[j] = 0.f;
    float hi, lo;
    split_tf32(nrm, hi, lo);
    o[kp - 4] = 1.f; o[kp - 3] = 1.f; o[kp - 2] = hi; o[kp - 1] = lo;
    atomicMax(pmax_bits, __float_as_uint(sqrtf(nrm) * 1.000001f));
}

// A operand + per-query margin E_q
__global__ void build_aaug_kernel(const float* __restrict__ q, const float* __restrict__ center, uint32_t nq, uint32_t d,
                                  uint32_t dpad, uint32_t kp, float pmax, float* __restrict__ aaug, float* __restrict__ q_margin) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const float* p = q + (size_t)i * dpad;
    float* o = aaug + (size_t)i * kp;
    float nrm = 0.f;
    for (uint32_t j = 0; j < d; ++j) {
        const float v = p[j] - center[j];
        nrm = nrm + v * v;
        o[j] = v;
    }
    for (uint32_t j = d; j < kp - 4; ++j) o[j] = 0.f;
    float hi, lo;
    split_tf32(nrm, hi, lo);
    o[kp - 4] = hi; o[kp - 3] = lo; o[kp - 2] = 1.f; o[kp - 1] = 1.f;
    const float qn = sqrtf(nrm) * 1.000001f;
    const float s = qn + pmax;
    q_margin[i] = 1.01f * 0.00390625f * qn * pmax + (float)(kp + 8) * 4.76837158203125e-07f * s * s;
}

template <int DVR, int K, int MT>
__global__ void __launch_bounds__((4 * MT + 2) * 32, 1)
knn_filter_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const FilterArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve: [A: MT*nkc chunks][B ring: stages chunks][barriers]
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* smem_a = smem;
    unsigned char* smem_b = smem + (size_t)MT * a.nkc * CHUNK_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)a.stages * CHUNK_BYTES);
    uint64_t* full_bar = bars;                       // [stages]
    uint64_t* empty_bar = bars + a.stages;           // [stages]
    uint64_t* tfull_bar = bars + 2 * a.stages;       // [NUM_ACC]
    uint64_t* tempty_bar = tfull_bar + NUM_ACC;      // [NUM_ACC]
    uint64_t* a_bar = tempty_bar + NUM_ACC;          // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_bar + 1);
    uint32_t* qbuf = tmem_slot + 4;  // per epilogue warp: 64 x u32 point rows, then 64 x u8 owner lanes

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int EPI_WARPS = 4 * MT;
    constexpr int TMEM_COLS = NUM_ACC * MT * BN;  // 256 or 512 (power of two)

    if (warp == EPI_WARPS && lane == 0) {
        for (uint32_t s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < NUM_ACC; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], EPI_WARPS); }
        mbar_init(a_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == EPI_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t row_base = blockIdx.x * (MT * BM);

    if (warp == EPI_WARPS) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_expect_tx(a_bar, (uint32_t)(MT * a.nkc * CHUNK_BYTES));
            for (int mt = 0; mt < MT; ++mt)
                for (uint32_t c = 0; c < a.nkc; ++c)
                    tma_load_2d(&map_a, a_bar, smem_a + (size_t)(mt * a.nkc + c) * CHUNK_BYTES, (int)(c * KC), (int)(row_base + mt * BM));
            uint32_t it = 0;
            for (uint32_t j = 0; j < a.n_tiles; ++j) {
                for (uint32_t c = 0; c < a.nkc; ++c, ++it) {
                    const uint32_t s = it % a.stages, ph = (it / a.stages) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    mbar_expect_tx(&full_bar[s], CHUNK_BYTES);
                    tma_load_2d(&map_b, &full_bar[s], smem_b + (size_t)s * CHUNK_BYTES, (int)(c * KC), (int)(j * BN));
                }
            }
        }
    } else if (warp == EPI_WARPS + 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
            mbar_wait(a_bar, 0);
            uint32_t it = 0;
            for (uint32_t j = 0; j < a.n_tiles; ++j) {
                const uint32_t as = j % NUM_ACC, aph = (j / NUM_ACC) & 1u;
                mbar_wait(&tempty_bar[as], aph ^ 1u);
                tc_fence_after();
                for (uint32_t c = 0; c < a.nkc; ++c, ++it) {
                    const uint32_t s = it % a.stages, ph = (it / a.stages) & 1u;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(smem_b + (size_t)s * CHUNK_BYTES);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint32_t a_addr = smem_u32(smem_a + (size_t)(mt * a.nkc + c) * CHUNK_BYTES);
                        const uint32_t d_tmem = tmem_base + (as * MT + mt) * BN;
#pragma unroll
                        for (int ks = 0; ks < KC / 8; ++ks)
                            tc_mma_tf32(d_tmem, make_desc_sw128(a_addr + ks * 32), make_desc_sw128(b_addr + ks * 32), idesc,
                                        (c > 0 || ks > 0) ? 1u : 0u);
                    }
                    tc_commit(&empty_bar[s]);  // smem slot free once these MMAs have read it
                }
                tc_commit(&tfull_bar[as]);     // accumulator tile(s) complete
            }
        }
    } else {
        // ================= epilogue: one query row per thread =================
        // Filter hits are not evaluated in place (a divergent chain of dependent L2 round trips
        // that would hold the accumulator stage hostage for the whole CTA): they are pushed to a
        // per-warp queue with ballot compaction and drained 32 at a time, the 32 lanes evaluating
        // 32 exact distances in parallel and handing each result to the lane that owns the query.
        const int mt = warp >> 2, quad = warp & 3;
        const uint32_t wrow0 = row_base + mt * BM + quad * 32;  // first query row of this warp
        const uint32_t qrow = wrow0 + lane;
        const bool active = qrow < a.nq;
        const DevTree<float>& t = a.t;
        const int DV = DVR > 0 ? DVR : (int)t.dv;
        uint32_t* q_prow = qbuf + warp * 64;
        unsigned char* q_owner = reinterpret_cast<unsigned char*>(qbuf + EPI_WARPS * 64) + warp * 64;
        TopK<float, K> topk;
        topk.init(active, a.k);
        if (a.floor_d && active) topk.set_floor(a.floor_d[qrow], a.floor_i[qrow]);
        const float margin = active ? a.q_margin[qrow] : 0.f;
        const float t2s = a.t2_scale;
        float theta = active ? pos_inf<float>() : -pos_inf<float>();
        const unsigned full = 0xffffffffu;
        const unsigned lt_mask = (1u << lane) - 1u;
        unsigned long long hits = 0;
        int qn = 0;  // queue fill, warp-uniform

        // exact evaluation of queue entries [0, cnt), cnt <= 32 (warp-uniform)
        auto drain = [&](int cnt) {
            float s = pos_inf<float>();
            uint32_t id = NO_ID;
            int o = 0;
            if (lane < cnt) {
                o = q_owner[lane];
                const uint32_t prow = q_prow[lane];
                const float4* qr = a.q + (size_t)(wrow0 + o) * DV;
                const float4* pr = t.pts + (size_t)prow * DV;
                float acc = 0.f;
                if (DVR > 0) {
#pragma unroll
                    for (int jc = 0; jc < (DVR > 0 ? DVR : 1); ++jc) acc = fold(acc, __ldg(qr + jc), __ldg(pr + jc));
                } else {
                    for (int jc = 0; jc < DV; ++jc) acc = fold(acc, __ldg(qr + jc), __ldg(pr + jc));
                }
                s = acc;
                id = __ldg(t.ids + prow);
            }
            for (int e = 0; e < cnt; ++e) {
                const float se = __shfl_sync(full, s, e);
                const uint32_t ide = __shfl_sync(full, id, e);
                const int oe = __shfl_sync(full, o, e);
                if (lane == oe && se <= topk.t2) {
                    topk.offer_sq(se, ide);
                    theta = xadd(xmul(topk.t2, t2s), margin);
                }
            }
        };

        for (uint32_t j = 0; j < a.n_tiles; ++j) {
            const uint32_t as = j % NUM_ACC, aph = (j / NUM_ACC) & 1u;
            mbar_wait(&tfull_bar[as], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (as * MT + mt) * BN;
#pragma unroll 1
            for (int cc = 0; cc < BN / 32; ++cc) {
                float v[32];
                tmem_ld32(taddr + cc * 32, v);
                float m0 = fminf(v[0], v[1]), m1 = fminf(v[2], v[3]), m2 = fminf(v[4], v[5]), m3 = fminf(v[6], v[7]);
#pragma unroll
                for (int i = 8; i < 32; i += 8) {
                    m0 = fminf(m0, fminf(v[i], v[i + 1])); m1 = fminf(m1, fminf(v[i + 2], v[i + 3]));
                    m2 = fminf(m2, fminf(v[i + 4], v[i + 5])); m3 = fminf(m3, fminf(v[i + 6], v[i + 7]));
                }
                const float m = fminf(fminf(m0, m1), fminf(m2, m3));
                if (__any_sync(full, m <= theta)) {
                    uint32_t bits = 0;
                    if (m <= theta) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) bits |= (v[i] <= theta ? 1u : 0u) << i;
                        const int valid = (int)t.n - (int)(j * BN + cc * 32);  // rows past the last point are zero-filled
                        if (valid < 32) bits &= valid > 0 ? ((1u << valid) - 1u) : 0u;
                    }
                    for (;;) {
                        const bool has = bits != 0;
                        const unsigned mask = __ballot_sync(full, has);
                        if (!mask) break;
                        if (has) {
                            const int i = __ffs(bits) - 1;
                            bits &= bits - 1;
                            const int slot = qn + __popc(mask & lt_mask);
                            q_prow[slot] = j * BN + cc * 32 + i;
                            q_owner[slot] = (unsigned char)lane;
                        }
                        qn += __popc(mask);
                        hits += __popc(mask);
                        __syncwarp();
                        if (qn >= 32) {
                            drain(32);
                            const int rest = qn - 32;  // < 32
                            uint32_t tp = 0; unsigned char to = 0;
                            if (lane < rest) { tp = q_prow[32 + lane]; to = q_owner[32 + lane]; }
                            __syncwarp();
                            if (lane < rest) { q_prow[lane] = tp; q_owner[lane] = to; }
                            __syncwarp();
                            qn = rest;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
            if (qn >= 16) { drain(qn); qn = 0; __syncwarp(); }
        }
        if (qn > 0) { drain(qn); qn = 0; }
        if (active) {
            const size_t base = (size_t)qrow * a.k;
            topk.store(a.part_d + base, a.part_i + base, a.k);
        }
        if (a.counters && lane == 0) atomicAdd(&a.counters[2], hits);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace tc
}  // namespace petal
