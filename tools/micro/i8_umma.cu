// Micro-benchmark: tcgen05.mma kind::i8 (UTCIMMA) on sm_100a -- (1) one 128 x N x 32 MMA against a CPU product
// (checks the shared-memory / instruction descriptor encodings), (2) issue-bound throughput of the 28-product
// slice schedule the int8-sliced FP64 contraction would use (7 accumulators in TMEM, operands resident in smem).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8_umma tools/micro/i8_umma.cu && ./i8_umma
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *b, uint32_t phase) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(phase) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t phase) {
    for (long long it = 0; !mbar_try_wait(b, phase); it++)
        if (it > 200000000ll) { printf("mbarrier timeout\n"); __trap(); }
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// canonical K-major no-swizzle operand: [row group of 8][k chunk of 16 B][row in group][16 B]
__device__ __host__ inline int canon_off(int row, int kbyte) { return (row >> 3) * 256 + (kbyte >> 4) * 128 + (row & 7) * 16 + (kbyte & 15); }

template <int N>
__global__ void __launch_bounds__(128) check_kernel(const int8_t *A, const int8_t *B, int32_t *C, uint32_t lbo, uint32_t sbo) {
    __shared__ __align__(128) int8_t sA[128 * 32];
    __shared__ __align__(128) int8_t sB[N * 32];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 32; i += 128) sA[canon_off(i / 32, i % 32)] = A[i];
    for (int i = tid; i < N * 32; i += 128) sB[canon_off(i / 32, i % 32)] = B[i];
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"((uint32_t)(N < 32 ? 32 : N)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        umma_i8(tb, smem_desc(smem_u32(sA), lbo, sbo), smem_desc(smem_u32(sB), lbo, sbo), make_idesc(128, N), 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; j++) C[tid * N + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"((uint32_t)(N < 32 ? 32 : N)) : "memory");
}

// A operand from TMEM: tcgen05.cp (smem -> TMEM, 128 lanes x 256 bit = the 32 int8 of one row) then the .ts form of the MMA
template <int N>
__global__ void __launch_bounds__(128) check_ts_kernel(const int8_t *A, const int8_t *B, int32_t *C, int mode) {
    __shared__ __align__(128) int8_t sA[128 * 32];
    __shared__ __align__(128) int8_t sB[N * 32];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 32; i += 128) sA[canon_off(i / 32, i % 32)] = A[i];
    for (int i = tid; i < N * 32; i += 128) sB[canon_off(i / 32, i % 32)] = B[i];
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base, ta = tb + 128;  // D at columns 0.., A at columns 128..135
    if (tid == 0) {
        if (mode == 0) {  // one 128x256b copy: both 16-byte k-chunks of a row through the descriptor's LBO
            asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(ta), "l"(smem_desc(smem_u32(sA), 128, 256)) : "memory");
        } else {          // two 128x128b copies: k-chunk 0 -> columns 0..3, k-chunk 1 -> columns 4..7
            asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(ta), "l"(smem_desc(smem_u32(sA), 128, 256)) : "memory");
            asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(ta + 4), "l"(smem_desc(smem_u32(sA) + 128, 128, 256)) : "memory");
        }
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}"
                     ::"r"(tb), "r"(ta), "l"(smem_desc(smem_u32(sB), 128, 256)), "r"(make_idesc(128, N)), "r"(0u) : "memory");
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; j++) C[tid * N + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u) : "memory");
}

// throughput of the A-from-TMEM schedule: per round 7 x tcgen05.cp (A planes -> TMEM columns 448..503) + 28 MMAs (N = 64)
__global__ void __launch_bounds__(128) rate_ts_kernel(int rounds, long long *cycles, int32_t *sink, int with_cp) {
    extern __shared__ __align__(128) int8_t sm[];
    int8_t *sA = sm;                  // 7 planes x 4 KB
    int8_t *sB = sm + 7 * 4096;       // 7 planes x 64*32
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 7 * 4096 + 7 * 64 * 32; i += 128) sm[i] = (int8_t)((i * 37 + 11) % 127 - 63);
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    long long t0 = clock64();
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, 64);
        const uint64_t da0 = smem_desc(smem_u32(sA), 128, 256), db0 = smem_desc(smem_u32(sB), 128, 256);
        for (int r = 0; r < rounds; r++) {
            if (with_cp || r == 0) {
#pragma unroll
                for (int i = 0; i < 7; i++)
                    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tb + 448 + 8 * i), "l"(da0 + (uint64_t)(i * 4096 >> 4)) : "memory");
            }
#pragma unroll
            for (int i = 0; i < 7; i++)
#pragma unroll
                for (int j = 0; j < 7; j++)
                    if (i + j < 7)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}"
                                     ::"r"(tb + (i + j) * 64), "r"(tb + 448 + 8 * i), "l"(db0 + (uint64_t)(j * 64 * 32 >> 4)), "r"(idesc), "r"(1u) : "memory");
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(tb + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (v[0] == 0x12345678u) sink[0] = (int32_t)v[1];
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

static void run_rate_ts(int with_cp) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long *dcy; int32_t *dsink;
    CK(cudaMalloc(&dcy, sms * 8)); CK(cudaMalloc(&dsink, 64));
    const int smem = 7 * 4096 + 7 * 64 * 32;
    CK(cudaFuncSetAttribute(rate_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int rounds = 4000;
    rate_ts_kernel<<<sms, 128, smem>>>(10, dcy, dsink, with_cp);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    rate_ts_kernel<<<sms, 128, smem>>>(rounds, dcy, dsink, with_cp);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> cy(sms);
    CK(cudaMemcpy(cy.data(), dcy, sms * 8, cudaMemcpyDeviceToHost));
    double ops = 2.0 * 128 * 64 * 32 * 28.0 * rounds * sms;
    printf("A from TMEM, N=64, 7 accumulators, %s: %.3f ms, %.2f POPS, %.1f clk per round of 28 MMAs (SM0), FP64-equivalent %.1f TF\n",
           with_cp ? "7 x tcgen05.cp per round" : "no copies (A resident)", ms, ops / (ms * 1e-3) / 1e15, (double)cy[0] / rounds, ops / 28.0 / (ms * 1e-3) / 1e12);
    cudaFree(dcy); cudaFree(dsink);
}

// Do FP64 instructions and the int8 tensor pipe overlap?  warp 0 issues the 28-product schedule for `rounds` rounds,
// warps 1..16 run `fp_iters` x 8 independent DFMA (or FFMA) chains each; report the kernel time for MMA only, FP only, both.
template <int FP32>
__global__ void __launch_bounds__(544) overlap_kernel(int rounds, int fp_iters, double *sink, int avoid_smsp0) {
    extern __shared__ __align__(128) int8_t sm[];
    int8_t *sA = sm, *sB = sm + 7 * 4096;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 7 * 4096 + 7 * 64 * 32; i += 544) sm[i] = (int8_t)((i * 37 + 11) % 127 - 63);
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    if (warp == 0) {
        if (tid == 0) {
            const uint32_t idesc = make_idesc(128, 64);
            const uint64_t da0 = smem_desc(smem_u32(sA), 128, 256), db0 = smem_desc(smem_u32(sB), 128, 256);
            for (int r = 0; r < rounds; r++) {
#pragma unroll
                for (int i = 0; i < 7; i++)
#pragma unroll
                    for (int j = 0; j < 7; j++)
                        if (i + j < 7) umma_i8(tb + (i + j) * 64, da0 + (uint64_t)(i * 4096 >> 4), db0 + (uint64_t)(j * 64 * 32 >> 4), idesc, 1u);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
        }
    } else if (!(avoid_smsp0 && (warp & 3) == 0)) {
        if (FP32) {
            float a[8], m = 1.0000001f, c = 1e-9f;
            for (int k = 0; k < 8; k++) a[k] = (float)(tid + k);
            for (int it = 0; it < fp_iters; it++)
#pragma unroll
                for (int k = 0; k < 8; k++) a[k] = fmaf(a[k], m, c);
            float t = 0; for (int k = 0; k < 8; k++) t += a[k];
            if (t == 12345.678f) sink[0] = t;
        } else {
            double a[8], m = 1.0000000001, c = 1e-12;
            for (int k = 0; k < 8; k++) a[k] = (double)(tid + k);
            for (int it = 0; it < fp_iters; it++)
#pragma unroll
                for (int k = 0; k < 8; k++) a[k] = fma(a[k], m, c);
            double t = 0; for (int k = 0; k < 8; k++) t += a[k];
            if (t == 12345.678) sink[0] = t;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

template <int FP32>
static void run_overlap() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *dsink;
    CK(cudaMalloc(&dsink, 64));
    const int smem = 7 * 4096 + 7 * 64 * 32;
    CK(cudaFuncSetAttribute(overlap_kernel<FP32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    auto timeit = [&](int rounds, int iters, int avoid = 0) {
        overlap_kernel<FP32><<<sms, 544, smem>>>(rounds, iters, dsink, avoid);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        overlap_kernel<FP32><<<sms, 544, smem>>>(rounds, iters, dsink, avoid);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        return ms;
    };
    const int rounds = 2000, iters = FP32 ? 40000 : 10000;
    float tm = timeit(rounds, 0), tf = timeit(0, iters), tb2 = timeit(rounds, iters);
    printf("overlap of int8 UMMA with %s FMA (16 warps x 8 chains): MMA only %.3f ms, FMA only %.3f ms, both %.3f ms  (sum %.3f, max %.3f)\n",
           FP32 ? "FP32" : "FP64", tm, tf, tb2, tm + tf, tm > tf ? tm : tf);
    float tf3 = timeit(0, iters, 1), tb3 = timeit(rounds, iters, 1);
    printf("   same with no FMA warp on the MMA issuer's scheduler (12 warps): FMA only %.3f ms, both %.3f ms\n", tf3, tb3);
    cudaFree(dsink);
}

// cta_group::2: a pair of CTAs (cluster of 2) shares one MMA of M = 256 (128 rows per CTA), N = 64: each CTA stages its own
// 128 rows of A and HALF of B (32 rows); D (128 lanes x 64 columns per accumulator) lands in each CTA's own TMEM.
// Rate of the 28-product schedule, issued by the leader CTA only.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) rate2_kernel(int rounds, long long *cycles, int32_t *sink, const int8_t *A,
                                                                             const int8_t *B, int32_t *C) {
    extern __shared__ __align__(128) int8_t sm[];
    int8_t *sA = sm;                  // 7 planes x 4 KB (my 128 rows)
    int8_t *sB = sm + 7 * 4096;       // 7 planes x 32 rows x 32 B = 1 KB (my half of B)
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (A) {  // correctness mode: plane 0 only, A (256 x 32), B (64 x 32) row-major int8
        for (int i = tid; i < 128 * 32; i += 128) sA[canon_off(i / 32, i % 32)] = A[(rank * 128 + i / 32) * 32 + i % 32];
        for (int i = tid; i < 32 * 32; i += 128) sB[canon_off(i / 32, i % 32)] = B[(rank * 32 + i / 32) * 32 + i % 32];
    } else {
        for (int i = tid; i < 7 * 4096 + 7 * 1024; i += 128) sm[i] = (int8_t)((i * 37 + 11) % 127 - 63);
    }
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    long long t0 = clock64();
    if (rank == 0 && tid == 0) {
        const uint32_t idesc = make_idesc(256, 64);
        const uint64_t da0 = smem_desc(smem_u32(sA), 128, 256), db0 = smem_desc(smem_u32(sB), 128, 256);
        if (A) {
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}"
                         ::"r"(tb), "l"(da0), "l"(db0), "r"(idesc), "r"(0u) : "memory");
        } else {
            for (int r = 0; r < rounds; r++) {
#pragma unroll
                for (int i = 0; i < 7; i++)
#pragma unroll
                    for (int j = 0; j < 7; j++)
                        if (i + j < 7)
                            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}"
                                         ::"r"(tb + (i + j) * 64), "l"(da0 + (uint64_t)(i * 4096 >> 4)), "l"(db0 + (uint64_t)(j * 1024 >> 4)), "r"(idesc), "r"(1u) : "memory");
            }
        }
        // arrive on the barrier of BOTH CTAs (same shared-memory offset in each)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((unsigned short)3) : "memory");
    }
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (A) {
        for (int c0 = 0; c0 < 64; c0 += 8) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(tb + ((uint32_t)(warp * 32) << 16) + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; j++) C[(rank * 128 + tid) * 64 + c0 + j] = (int32_t)v[j];
        }
    } else {
        uint32_t v[4];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(tb + ((uint32_t)(warp * 32) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (v[0] == 0x12345678u) sink[0] = (int32_t)v[1];
    }
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

static void run_rate2() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    sms &= ~1;
    long long *dcy; int32_t *dsink;
    CK(cudaMalloc(&dcy, sms * 8)); CK(cudaMalloc(&dsink, 64));
    const int smem = 7 * 4096 + 7 * 1024;
    CK(cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // correctness: one MMA, M = 256, N = 64, K = 32
    {
        std::vector<int8_t> A(256 * 32), B(64 * 32);
        srand(3);
        for (auto &v : A) v = (int8_t)(rand() % 255 - 127);
        for (auto &v : B) v = (int8_t)(rand() % 255 - 127);
        int8_t *dA, *dB; int32_t *dC;
        CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dC, 256 * 64 * 4));
        CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
        CK(cudaMemset(dC, 0xff, 256 * 64 * 4));
        rate2_kernel<<<2, 128, smem>>>(1, dcy, dsink, dA, dB, dC);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("cta_group::2 check: CUDA error %s\n", cudaGetErrorString(e)); return; }
        std::vector<int32_t> Cc(256 * 64);
        CK(cudaMemcpy(Cc.data(), dC, Cc.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int i = 0; i < 256; i++)
            for (int j = 0; j < 64; j++) {
                int32_t s = 0;
                for (int k = 0; k < 32; k++) s += (int32_t)A[i * 32 + k] * (int32_t)B[j * 32 + k];
                if (s != Cc[i * 64 + j]) { if (bad < 3) printf("  mismatch (%d,%d): got %d want %d\n", i, j, Cc[i * 64 + j], s); bad++; }
            }
        printf("check cta_group::2 256x64x32 int8 UMMA: %s (%d mismatches)\n", bad ? "FAIL" : "OK", bad);
        cudaFree(dA); cudaFree(dB); cudaFree(dC);
        if (bad) return;
    }
    const int rounds = 4000;
    rate2_kernel<<<sms, 128, smem>>>(10, dcy, dsink, nullptr, nullptr, nullptr);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    rate2_kernel<<<sms, 128, smem>>>(rounds, dcy, dsink, nullptr, nullptr, nullptr);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> cy(sms);
    CK(cudaMemcpy(cy.data(), dcy, sms * 8, cudaMemcpyDeviceToHost));
    double ops = 2.0 * 256 * 64 * 32 * 28.0 * rounds * (sms / 2);
    printf("cta_group::2, M=256 (2 x 128), N=64, 7 accumulators: %.3f ms, %.2f POPS, %.1f clk per MMA (pair 0), FP64-equivalent %.1f TF\n", ms,
           ops / (ms * 1e-3) / 1e15, (double)cy[0] / (28.0 * rounds), ops / 28.0 / (ms * 1e-3) / 1e12);
    cudaFree(dcy); cudaFree(dsink);
}

template <int N>
static int run_check_ts(int mode) {
    std::vector<int8_t> A(128 * 32), B(N * 32);
    srand(11);
    for (auto &v : A) v = (int8_t)(rand() % 255 - 127);
    for (auto &v : B) v = (int8_t)(rand() % 255 - 127);
    int8_t *dA, *dB;
    int32_t *dC;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dC, 128 * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xff, 128 * N * 4));
    check_ts_kernel<N><<<1, 128>>>(dA, dB, dC, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("check_ts mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return -1; }
    std::vector<int32_t> C(128 * N);
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < 128; i++)
        for (int j = 0; j < N; j++) {
            int32_t s = 0;
            for (int k = 0; k < 32; k++) s += (int32_t)A[i * 32 + k] * (int32_t)B[j * 32 + k];
            if (s != C[i * N + j]) { if (bad < 3) printf("  mismatch (%d,%d): got %d want %d\n", i, j, C[i * N + j], s); bad++; }
        }
    printf("check A-from-TMEM 128x%dx32 (cp mode %d: %s): %s (%d mismatches)\n", N, mode, mode ? "2 x 128x128b" : "1 x 128x256b", bad ? "FAIL" : "OK", bad);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return bad;
}

// throughput: NACC accumulators of N columns, NPROD products per round (slice pair schedule), operands resident
template <int N, int NACC, int NPROD>
__global__ void __launch_bounds__(128) rate_kernel(int rounds, long long *cycles, int32_t *sink) {
    extern __shared__ __align__(128) int8_t sm[];
    int8_t *sA = sm;                  // 7 slices x 4 KB
    int8_t *sB = sm + 7 * 4096;       // 7 slices x N*32
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 7 * 4096 + 7 * N * 32; i += 128) sm[i] = (int8_t)((i * 37 + 11) % 127 - 63);
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    constexpr uint32_t COLS = (N * NACC <= 32) ? 32 : (N * NACC <= 64) ? 64 : (N * NACC <= 128) ? 128 : (N * NACC <= 256) ? 256 : 512;
    static_assert(N * NACC <= 512, "TMEM");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    long long t0 = clock64();
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, N);
        const uint64_t da0 = smem_desc(smem_u32(sA), 128, 256), db0 = smem_desc(smem_u32(sB), 128, 256);
        for (int r = 0; r < rounds; r++) {
            int p = 0;
#pragma unroll
            for (int i = 0; i < 7; i++)
#pragma unroll
                for (int j = 0; j < 7; j++)
                    if (i + j < 7 && p < NPROD) {  // compile-time schedule: descriptors differ by constants
                        umma_i8(tb + ((i + j) % NACC) * N, da0 + (uint64_t)(i * 4096 >> 4), db0 + (uint64_t)(j * N * 32 >> 4), idesc, 1u);
                        p++;
                    }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(tb + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (v[0] == 0x12345678u) sink[0] = (int32_t)v[1];
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(COLS) : "memory");
}

template <int N>
static int run_check(uint32_t lbo, uint32_t sbo) {
    std::vector<int8_t> A(128 * 32), B(N * 32);
    srand(7);
    for (auto &v : A) v = (int8_t)(rand() % 255 - 127);
    for (auto &v : B) v = (int8_t)(rand() % 255 - 127);
    int8_t *dA, *dB;
    int32_t *dC;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dC, 128 * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xff, 128 * N * 4));
    check_kernel<N><<<1, 128>>>(dA, dB, dC, lbo, sbo);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("check N=%d lbo=%u sbo=%u: CUDA error %s\n", N, lbo, sbo, cudaGetErrorString(e)); return -1; }
    std::vector<int32_t> C(128 * N);
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < 128; i++)
        for (int j = 0; j < N; j++) {
            int32_t s = 0;
            for (int k = 0; k < 32; k++) s += (int32_t)A[i * 32 + k] * (int32_t)B[j * 32 + k];
            if (s != C[i * N + j]) { if (bad < 4) printf("  mismatch (%d,%d): got %d want %d\n", i, j, C[i * N + j], s); bad++; }
        }
    printf("check 128x%dx32 int8 UMMA, lbo=%u sbo=%u: %s (%d mismatches)\n", N, lbo, sbo, bad ? "FAIL" : "OK", bad);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return bad;
}

template <int N, int NACC, int NPROD>
static void run_rate(const char *what) {
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    long long *dcy; int32_t *dsink;
    CK(cudaMalloc(&dcy, sms * 8)); CK(cudaMalloc(&dsink, 64));
    const int smem = 7 * 4096 + 7 * N * 32;
    CK(cudaFuncSetAttribute(rate_kernel<N, NACC, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int rounds = 4000;
    rate_kernel<N, NACC, NPROD><<<sms, 128, smem>>>(10, dcy, dsink);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    rate_kernel<N, NACC, NPROD><<<sms, 128, smem>>>(rounds, dcy, dsink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> cy(sms);
    CK(cudaMemcpy(cy.data(), dcy, sms * 8, cudaMemcpyDeviceToHost));
    double ops = 2.0 * 128 * N * 32 * (double)NPROD * rounds * sms;
    printf("%-44s N=%3d acc=%d prod/round=%2d: %.3f ms, %.2f POPS (int8), %.1f clk per MMA (SM0), FP64-equivalent %.1f TF at 28 products\n", what, N, NACC,
           NPROD, ms, ops / (ms * 1e-3) / 1e15, (double)cy[0] / ((double)NPROD * rounds), ops / 28.0 / (ms * 1e-3) / 1e12);
    cudaFree(dcy); cudaFree(dsink);
}

int main() {
    int bad = 0;
    bad += run_check<64>(128, 256) != 0;
    bad += run_check<128>(128, 256) != 0;
    if (bad) { printf("descriptor interpretation A failed; trying swapped LBO/SBO\n"); run_check<64>(256, 128); }
    run_check_ts<64>(0);
    run_check_ts<64>(1);
    run_rate2();
    run_overlap<0>();
    run_overlap<1>();
    run_rate_ts(0);
    run_rate_ts(1);
    run_rate<64, 7, 28>("7 accumulators, N=64 (slice schedule)");
    run_rate<128, 4, 28>("4 accumulators, N=128");
    run_rate<256, 2, 28>("2 accumulators, N=256");
    run_rate<32, 7, 28>("7 accumulators, N=32");
    return 0;
}
