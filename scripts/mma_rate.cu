// mma_rate.cu -- micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128) on one SM as a function of
// operand source (SS / TS), swizzle mode, N, number of issuing threads and commit cadence.  Operands are
// whatever is in shared memory (zeros); only the issue/execute rate is measured.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate scripts/mma_rate.cu && /tmp/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad),
                 "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem),
                 "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
// swz: 0 none(interleave), 2 = 128B, 4 = 64B, 6 = 32B ; sbo in bytes
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t swz, uint32_t sbo, uint32_t lbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)swz << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int dfmt) {
    return ((uint32_t)dfmt << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Cfg {
    int ts;        // 1: A from TMEM
    int swz;       // 2 / 4 / 6 / 0
    int n;         // MMA N
    int issuers;   // 1 or 2 issuing threads (different warps)
    int commit_every;  // commit to a dummy barrier every c MMAs (0 = only at the end)
    int reps;      // MMAs per issuer
    int dfmt;      // 1 = f32 accumulate, 0 = f16 accumulate
    int kadv;      // 1: walk K inside swizzle rows and over 8 KB slots like the real kernel; 0: same operands
    int uni;       // 1: the whole warp runs the issue loop on warp-uniform values, elect.sync picks the issuing lane
};

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Cfg c, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t done_bar[2], dummy_bar[2];
    __shared__ uint32_t tmem_slot;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&done_bar[i], 1); mbar_init(&dummy_bar[i], (1 << 20) - 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);
    if (warp < c.issuers && (c.uni || lane == 0)) {
        const uint32_t idesc = make_idesc_f16(128, c.n, c.dfmt);
        const uint32_t row_bytes = c.swz == 2 ? 128 : (c.swz == 4 ? 64 : 32);   // no swizzle: 2 core matrices of 16 B per row
        const uint32_t sbo = c.swz ? 8 * row_bytes : 128;          // 8-row group stride
        const uint32_t lbo = c.swz ? 16 : 256 * 16;                // no-swizzle: next K core matrix
        const uint32_t ksteps = row_bytes / 32;                    // MMAs (K=16 fp16 = 32 B) per swizzle row
        const uint32_t a_bytes = 128 * row_bytes, b_bytes = c.n * row_bytes;
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 48 * 1024;   // A region 48 KB, B ring 96 KB
        // SS: each issuer owns 512/issuers columns (two accumulators when they fit); TS: the last 64 columns
        // hold the A operand (8 columns per K=16 step), the accumulators share the first 448
        const uint32_t region = (c.ts ? 448u : 512u) / c.issuers;
        const uint32_t nacc = region / c.n >= 2 ? 2 : 1;
        const uint32_t a_tmem = tmem_base + 448;
        // no divisions in the timed loop: 4 A slots, 4 B slots, power-of-two K steps
        const uint32_t kshift = ksteps == 4 ? 2 : (ksteps == 2 ? 1 : 0);
        const uint64_t ad0 = make_desc(a_base, c.swz, sbo, lbo), bd0 = make_desc(b_base, c.swz, sbo, lbo);
        const uint32_t a_step = a_bytes >> 4, b_step = b_bytes >> 4;
        const uint32_t d0 = tmem_base + warp * region, d1 = d0 + (nacc - 1) * c.n;
        const int ce = c.commit_every;
        long long t0 = clock64();
        int since = 0;
        if (c.uni == 2) {
            // 8 MMAs per iteration, all descriptors loop-invariant and live at once (distinct uniform registers)
            uint64_t adu[8], bdu[8];
#pragma unroll
            for (uint32_t u = 0; u < 8; ++u) {
                const uint32_t slot = c.kadv ? (u >> kshift) & 3u : 0u, ks = c.kadv ? u & (ksteps - 1) : 0u;
                adu[u] = ad0 + slot * a_step + ks * 2; bdu[u] = bd0 + slot * b_step + ks * 2;
            }
            for (int r = 0; r < c.reps; r += 8) {
                const uint32_t d = (r & 8) ? d1 : d0;
                if (elect_one()) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (c.ts) mma_ts(d, a_tmem + u * 8, bdu[u], idesc, u ? 1u : 0u);
                        else mma_ss(d, adu[u], bdu[u], idesc, u ? 1u : 0u);
                        if (ce && (u % ce) == ce - 1) tc_commit(&dummy_bar[warp]);
                    }
                }
                __syncwarp();
            }
        } else
        for (int r = 0; r < c.reps; ++r) {
            const uint32_t slot = c.kadv ? ((uint32_t)r >> kshift) & 3u : 0u, ks = c.kadv ? (uint32_t)r & (ksteps - 1) : 0u;
            const uint64_t ad = ad0 + slot * a_step + ks * 2, bd = bd0 + slot * b_step + ks * 2;
            const uint32_t d = (r & 8) ? d1 : d0;
            if (c.uni) {
                if (elect_one()) {
                    if (c.ts) mma_ts(d, a_tmem + (r & 7) * 8, bd, idesc, (r & 7) ? 1u : 0u);
                    else mma_ss(d, ad, bd, idesc, (r & 7) ? 1u : 0u);
                    if (ce && since + 1 == ce) tc_commit(&dummy_bar[warp]);
                }
                __syncwarp();
                if (ce && ++since == ce) since = 0;
            } else {
                if (c.ts) mma_ts(d, a_tmem + (r & 7) * 8, bd, idesc, (r & 7) ? 1u : 0u);
                else mma_ss(d, ad, bd, idesc, (r & 7) ? 1u : 0u);
                if (ce && ++since == ce) { since = 0; tc_commit(&dummy_bar[warp]); }
            }
        }
        long long t1 = clock64();
        if (lane == 0) tc_commit(&done_bar[warp]);
        mbar_wait(&done_bar[warp], 0);
        long long t2 = clock64();
        if (lane == 0) {
            out[(blockIdx.x * 2 + warp) * 2 + 0] = t1 - t0;
            out[(blockIdx.x * 2 + warp) * 2 + 1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * 4 * sizeof(long long));
    const size_t smem = 161 * 1024 + 1024;
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const Cfg cfgs[] = {
        // ts swz  n  iss commit reps dfmt kadv uni
        {0, 4, 128, 1, 0, 4000, 1, 1, 0},   // current kernel's MMA: SS, 64B swizzle, N=128, divergent single-lane issue
        {0, 4, 128, 1, 0, 4000, 1, 1, 1},   // warp-uniform issue
        {0, 4, 128, 1, 0, 4000, 1, 1, 2},   // unrolled x8, distinct descriptors
        {0, 4, 128, 1, 0, 4000, 1, 0, 2},   // same operands every time
        {0, 4, 128, 2, 0, 4000, 1, 1, 2},   // two issuers
        {0, 4, 128, 1, 1, 4000, 1, 1, 2},   // commit every MMA
        {0, 4, 128, 1, 2, 4000, 1, 1, 2},   // commit every 2
        {0, 4, 128, 1, 4, 4000, 1, 1, 2},
        {0, 4, 128, 2, 2, 4000, 1, 1, 2},   // two issuers, commit every 2
        {0, 2, 128, 1, 0, 4000, 1, 1, 2},   // 128B swizzle
        {0, 6, 128, 1, 0, 4000, 1, 1, 2},   // 32B swizzle
        {0, 0, 128, 1, 0, 4000, 1, 1, 2},   // no swizzle (core-matrix interleave)
        {0, 4, 256, 1, 0, 4000, 1, 1, 2},   // N=256
        {0, 4, 256, 1, 2, 4000, 1, 1, 2},
        {0, 4, 64, 1, 0, 4000, 1, 1, 2},    // N=64
        {0, 4, 32, 1, 0, 4000, 1, 1, 2},    // N=32
        {1, 4, 128, 1, 0, 4000, 1, 1, 2},   // TS: A in TMEM
        {1, 4, 128, 2, 0, 4000, 1, 1, 2},
        {1, 4, 128, 1, 2, 4000, 1, 1, 2},
        {0, 4, 128, 1, 0, 4000, 0, 1, 2},   // f16 accumulate
        {1, 4, 64, 1, 0, 4000, 1, 1, 2},    // TS N=64
        {1, 4, 64, 2, 0, 4000, 1, 1, 2},    // TS N=64, two issuers
        {0, 4, 64, 2, 0, 4000, 1, 1, 2},    // SS N=64, two issuers
        {0, 4, 256, 2, 0, 4000, 1, 1, 2},   // SS N=256, two issuers
        {1, 4, 256, 1, 0, 4000, 1, 1, 2},   // TS N=256
        {1, 4, 192, 2, 0, 4000, 1, 1, 2},   // TS N=192, two issuers
    };
    for (int grid : {1}) {
        for (const Cfg& c : cfgs) {
            cudaMemset(out, 0, 148 * 4 * sizeof(long long));
            mma_rate_kernel<<<grid, 128, smem>>>(c, out);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[4];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            printf("grid %3d  uni %d %s swz %d N %3d issuers %d commit/%d dfmt %d kadv %d : issue %.1f cyc/MMA, complete %.1f cyc/MMA (per issuer; x%d in parallel) %s\n",
                   grid, c.uni, c.ts ? "TS" : "SS", c.swz, c.n, c.issuers, c.commit_every, c.dfmt, c.kadv, (double)h[0] / c.reps, (double)h[1] / c.reps,
                   c.issuers, e == cudaSuccess ? "" : cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
        }
    }
    return 0;
}
