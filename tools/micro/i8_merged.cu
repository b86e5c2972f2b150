// Micro-benchmark for the stacked-N schedule of aqml_b200/csrc/i8pair.cuh (round 2):
//   (1) correctness of ONE UTCIMMA whose B operand is several digit planes stacked along N (N = 256: four planes of 64
//       rows one behind the other), output slices = consecutive accumulators, against a CPU product;
//   (2) rate of the 22- / 26-product schedule (8 / 9 MMAs per 32-column stage, M 128, N 64..256) with the operands
//       resident in shared memory;
//   (3) the same inside the real main loop: a producer thread refills a 4-stage ring with cp.async.bulk (36 KB per
//       stage, L2 resident source), the MMA thread consumes it -- what shared-memory bandwidth and the L2 leave of (2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8_merged tools/micro/i8_merged.cu && ./i8_merged
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t *b, uint32_t phase) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(phase) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t phase) {
    for (long long it = 0; !mbar_try_wait(b, phase); it++)
        if (it > 200000000ll) { printf("mbarrier timeout\n"); __trap(); }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

constexpr int S = 6, A_CHUNK = S * 2048, B_CHUNK = S * 1024, STAGE = 2 * (A_CHUNK + B_CHUNK);

// the kernel's stage layout: [k-chunk c][plane p][row][16 B] for A (128 rows) then for B (64 rows)
__device__ __host__ inline int offA(int c, int p, int row, int kb) { return c * A_CHUNK + p * 2048 + row * 16 + kb; }
__device__ __host__ inline int offB(int c, int p, int row, int kb) { return 2 * A_CHUNK + c * B_CHUNK + p * 1024 + row * 16 + kb; }

template <int VAR>
__device__ __forceinline__ void issue_stage(uint32_t tb, uint32_t st, uint32_t first) {
    constexpr int NMMA = VAR ? 9 : 8;
    constexpr int mi[9] = {0, 0, 1, 1, 2, VAR ? 2 : 3, VAR ? 3 : 4, VAR ? 4 : 5, 5};
    constexpr int mj[9] = {0, 4, 0, VAR ? 4 : 3, 0, VAR ? 3 : 0, 0, 0, 0};
    constexpr int mn[9] = {4, 2, VAR ? 4 : 3, 2, VAR ? 3 : 4, VAR ? 2 : 4, VAR ? 4 : 2, VAR ? 3 : 1, 2};
    const uint64_t da = smem_desc(st, A_CHUNK, 128), db = smem_desc(st + 2 * A_CHUNK, B_CHUNK, 128);
#pragma unroll
    for (int m = 0; m < NMMA; m++)
        umma_i8(tb + (mi[m] + mj[m]) * 64, da + (uint64_t)(mi[m] * 2048 >> 4), db + (uint64_t)(mj[m] * 1024 >> 4), make_idesc(128, 64 * mn[m]),
                (first && mi[m] == 0) ? 0u : 1u);
}

// (1) one stage of the schedule on real data; C[g][row][col] = level sums P_g
template <int VAR>
__global__ void __launch_bounds__(128) check_kernel(const int8_t *stage_img, int32_t *C) {
    extern __shared__ __align__(1024) int8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < STAGE; i += 128) sm[i] = stage_img[i];
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
    {   // accumulator 6 is not cleared by plane 0: zero it like the kernel's epilogue does
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + 6 * 64;
        for (int c0 = 0; c0 < 64; c0 += 8) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + c0), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        issue_stage<VAR>(tb, smem_u32(sm), 1u);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 7 * 64; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; j++) C[((c0 + j) / 64 * 128 + tid) * 64 + (c0 + j) % 64] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

// (2) / (3): warp 0 lane 0 = producer (if fill), warp 1 lane 0 = MMA issuer
template <int VAR>
__global__ void __launch_bounds__(128) rate_kernel(int rounds, int fill, const int8_t *src, long long src_bytes, long long *cycles) {
    extern __shared__ __align__(1024) int8_t sm[];
    __shared__ __align__(8) uint64_t full[4], empty[4], done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 4 * STAGE; i += 128) sm[i] = (int8_t)((i * 37 + 11) % 255 - 127);
    if (tid == 0) {
        for (int s = 0; s < 4; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    const long long t0 = clock64();
    if (warp == 0 && (tid & 31) == 0 && fill) {
        // every CTA walks its own window of the source (like the kernel: tiles of different rows), L2 resident
        long long pos = ((long long)blockIdx.x * 7919 * STAGE) % (src_bytes - STAGE);
        for (int r = 0; r < rounds; r++) {
            const int s = r & 3;
            if (r >= 4) mbar_wait(&empty[s], ((r >> 2) - 1) & 1);
            mbar_expect_tx(&full[s], STAGE);
            int8_t *st = sm + s * STAGE;
            for (int c = 0; c < 2; c++) {  // the kernel's copies: A chunk 12 KB, B chunk as 6 x 1 KB
                bulk_g2s(st + c * A_CHUNK, src + pos + c * A_CHUNK, A_CHUNK, &full[s]);
                for (int p = 0; p < S; p++) bulk_g2s(st + 2 * A_CHUNK + c * B_CHUNK + p * 1024, src + pos + 2 * A_CHUNK + c * B_CHUNK + p * 1024, 1024, &full[s]);
            }
            pos += STAGE;
            if (pos + STAGE > src_bytes) pos = 0;
        }
    } else if (warp == 1 && (tid & 31) == 0) {
        for (int r = 0; r < rounds; r++) {
            const int s = r & 3;
            if (fill) {
                mbar_wait(&full[s], (r >> 2) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            issue_stage<VAR>(tb, smem_u32(sm + s * STAGE), 0u);
            if (fill) umma_commit(&empty[s]);
        }
        umma_commit(&done);
        mbar_wait(&done, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

// (5) the main loop of (3) on a 2-CTA cluster that SHARES the A tile: each CTA loads half of the A planes (3 of 6) and multicasts
// it into both CTAs' stage buffers, its own B tile stays private: 24 KB instead of 36 KB per stage and SM from L2.  A stage
// may be refilled when BOTH CTAs' MMAs have read it: the MMA threads' commits arrive on both CTAs' `empty` barriers (count 2).
__device__ __forceinline__ void bulk_g2s_mc(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
template <int VAR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) rate_cluster_kernel(int rounds, const int8_t *src, long long src_bytes, long long *cycles) {
    extern __shared__ __align__(1024) int8_t sm[];
    __shared__ __align__(8) uint64_t full[4], empty[4], done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = tid; i < 4 * STAGE; i += 128) sm[i] = (int8_t)((i * 37 + 11) % 255 - 127);
    if (tid == 0) {
        for (int s = 0; s < 4; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); }
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    const long long t0 = clock64();
    if (warp == 0 && (tid & 31) == 0) {
        // both CTAs of a cluster walk the same A window (cluster id), their own B windows
        const long long cl = blockIdx.x >> 1;
        long long posA = (cl * 7919 * STAGE) % (src_bytes - STAGE), posB = ((long long)blockIdx.x * 104729 * STAGE) % (src_bytes - STAGE);
        for (int r = 0; r < rounds; r++) {
            const int s = r & 3;
            if (r >= 4) mbar_wait(&empty[s], ((r >> 2) - 1) & 1);
            mbar_expect_tx(&full[s], STAGE);
            int8_t *st = sm + s * STAGE;
            for (int c = 0; c < 2; c++) {
                // my half of the A chunk (3 planes = 6 KB) into BOTH CTAs; my B chunk (6 x 1 KB) into mine
                bulk_g2s_mc(st + c * A_CHUNK + rank * (A_CHUNK / 2), src + posA + c * A_CHUNK + rank * (A_CHUNK / 2), A_CHUNK / 2, &full[s], (uint16_t)3);
                for (int p = 0; p < S; p++) bulk_g2s(st + 2 * A_CHUNK + c * B_CHUNK + p * 1024, src + posB + 2 * A_CHUNK + c * B_CHUNK + p * 1024, 1024, &full[s]);
            }
            posA += STAGE; posB += STAGE;
            if (posA + STAGE > src_bytes) posA = 0;
            if (posB + STAGE > src_bytes) posB = 0;
        }
    } else if (warp == 1 && (tid & 31) == 0) {
        for (int r = 0; r < rounds; r++) {
            const int s = r & 3;
            mbar_wait(&full[s], (r >> 2) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_stage<VAR>(tb, smem_u32(sm + s * STAGE), 0u);
            umma_commit_mc(&empty[s], (uint16_t)3);
        }
        umma_commit(&done);
        mbar_wait(&done, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

template <int VAR>
static void run_rate_cluster(const int8_t *src, long long src_bytes) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    sms &= ~1;
    long long *dcy;
    CK(cudaMalloc(&dcy, sms * 8));
    CK(cudaFuncSetAttribute(rate_cluster_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * STAGE));
    const int rounds = 20000;
    rate_cluster_kernel<VAR><<<sms, 128, 4 * STAGE>>>(200, src, src_bytes, dcy);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    rate_cluster_kernel<VAR><<<sms, 128, 4 * STAGE>>>(rounds, src, src_bytes, dcy);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> cy(sms);
    CK(cudaMemcpy(cy.data(), dcy, sms * 8, cudaMemcpyDeviceToHost));
    const int nprod = VAR ? 26 : 22;
    const double ops = 2.0 * 128 * 64 * 32 * nprod * (double)rounds * sms;
    printf("%d products, 2-CTA clusters sharing the A tile by multicast (24 KB per stage and SM from L2): %.3f ms, %.2f POPS (int8), %.1f clk per stage (SM0)\n",
           nprod, ms, ops / (ms * 1e-3) / 1e15, (double)cy[0] / rounds);
    cudaFree(dcy);
}

template <int VAR>
static void run_check() {
    std::vector<int8_t> img(STAGE);
    srand(7 + VAR);
    for (auto &v : img) v = (int8_t)(rand() % 256 - 128);
    int8_t *dimg; int32_t *dC;
    CK(cudaMalloc(&dimg, STAGE)); CK(cudaMalloc(&dC, 7 * 128 * 64 * 4));
    CK(cudaMemcpy(dimg, img.data(), STAGE, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(check_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE));
    check_kernel<VAR><<<1, 128, STAGE>>>(dimg, dC);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> C(7 * 128 * 64);
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int g = 0; g < 7; g++)
        for (int r = 0; r < 128; r++)
            for (int c = 0; c < 64; c++) {
                long long ref = 0;
                for (int i = 0; i < 6; i++) {
                    const int j = g - i;
                    if (j < 0 || j >= 6) continue;
                    if (!(VAR ? (i + j <= 6) : (i + j <= 5 || (i == 3 && j == 3)))) continue;
                    for (int k = 0; k < 32; k++) ref += (long long)img[offA(k >> 4, i, r, k & 15)] * img[offB(k >> 4, j, c, k & 15)];
                }
                if (ref != C[(g * 128 + r) * 64 + c]) bad++;
            }
    printf("check stacked-N schedule, %d products (%d MMAs): %s (%lld mismatches of %d)\n", VAR ? 26 : 22, VAR ? 9 : 8, bad ? "FAILED" : "OK", bad, 7 * 128 * 64);
    cudaFree(dimg); cudaFree(dC);
}

template <int VAR>
static void run_rate(int fill, const int8_t *src, long long src_bytes) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long *dcy;
    CK(cudaMalloc(&dcy, sms * 8));
    CK(cudaFuncSetAttribute(rate_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * STAGE));
    const int rounds = 20000;
    rate_kernel<VAR><<<sms, 128, 4 * STAGE>>>(200, fill, src, src_bytes, dcy);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    rate_kernel<VAR><<<sms, 128, 4 * STAGE>>>(rounds, fill, src, src_bytes, dcy);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> cy(sms);
    CK(cudaMemcpy(cy.data(), dcy, sms * 8, cudaMemcpyDeviceToHost));
    const int nprod = VAR ? 26 : 22;
    const double ops = 2.0 * 128 * 64 * 32 * nprod * (double)rounds * sms;
    printf("%d products, %s: %.3f ms, %.2f POPS (int8), %.1f clk per 32-column stage (SM0), FP64-equivalent %.1f TF%s\n", nprod,
           fill ? "4-stage ring refilled by cp.async.bulk (36 KB per stage)" : "operands resident in shared memory", ms, ops / (ms * 1e-3) / 1e15,
           (double)cy[0] / rounds, ops / nprod / (ms * 1e-3) / 1e12,
           fill ? "" : "");
    if (fill) printf("    L2 -> SM traffic %.2f TB/s\n", (double)STAGE * rounds * sms / (ms * 1e-3) / 1e12);
    cudaFree(dcy);
}

// (4) how do FP64 instructions progress while UTCIMMA of a given N stream through the tensor pipe?  warp 0 lane 0 issues
// back-to-back MMAs (M 128, N = n_mma, operands resident); warps 1 .. nw run `chains` independent dependent-DFMA chains
// of `len` steps each and report clk per chain step.  mode 1: DFMA replaced by FFMA (FP32), mode 2: IMAD.
template <int MODE>
__global__ void __launch_bounds__(544) contention_kernel(int n_mma, int mma_rounds, int len, int chains, long long *cyc, double *sink) {
    extern __shared__ __align__(1024) int8_t sm[];
    __shared__ __align__(8) uint64_t done;
    __shared__ uint32_t tmem_base;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < STAGE; i += blockDim.x) sm[i] = (int8_t)((i * 37 + 11) % 255 - 127);
    if (tid == 0) { mbar_init(&done, 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    const long long t0 = clock64();
    if (warp == 0) {
        if (tid == 0 && n_mma > 0) {
            const uint64_t da = smem_desc(smem_u32(sm), A_CHUNK, 128), db = smem_desc(smem_u32(sm) + 2 * A_CHUNK, B_CHUNK, 128);
            const uint32_t id = make_idesc(128, n_mma);
            for (int r = 0; r < mma_rounds; r++) {
                umma_i8(tb, da, db, id, 1u);
                umma_i8(tb + 256, da, db, id, 1u);
            }
            umma_commit(&done);
            mbar_wait(&done, 0);
            cyc[blockIdx.x * 2] = clock64() - t0;
        }
    } else {
        double a[8];
        float f[8];
        int q[8];
        for (int k = 0; k < 8; k++) { a[k] = (double)(tid + k); f[k] = (float)(tid + k); q[k] = tid + k; }
        const double m = 1.0000000001, c = 1e-12;
        const long long s0 = clock64();
        for (int it = 0; it < len; it++) {
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (k < chains) {
                    if (MODE == 0) a[k] = fma(a[k], m, c);
                    else if (MODE == 1) f[k] = fmaf(f[k], 1.0000001f, 1e-9f);
                    else q[k] = q[k] * 3 + 7;
                }
        }
        const long long s1 = clock64();
        double t = 0;
        for (int k = 0; k < 8; k++) t += a[k] + f[k] + q[k];
        if (t == 12345.678) sink[0] = t;
        if (tid == 32) cyc[blockIdx.x * 2 + 1] = s1 - s0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

template <int MODE>
static void run_contention() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long *dcy; double *dsink;
    CK(cudaMalloc(&dcy, sms * 16)); CK(cudaMalloc(&dsink, 64));
    CK(cudaFuncSetAttribute(contention_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE));
    const char *name = MODE == 0 ? "DFMA" : MODE == 1 ? "FFMA" : "IMAD";
    const int len = 4000;
    for (int nw : {1, 16})
        for (int chains : {1, 4})
            for (int n_mma : {0, 64, 128, 256}) {
                // enough MMAs to outlast the FP warps: assume <= 200 clk per chain step
                const int mma_rounds = n_mma ? (int)(200.0 * len * 4 / (n_mma / 2 + 16) / 2) : 0;
                std::vector<long long> cy(sms * 2, 0);
                CK(cudaMemset(dcy, 0, sms * 16));
                contention_kernel<MODE><<<sms, 32 * (1 + nw), STAGE>>>(n_mma, mma_rounds, len, chains, dcy, dsink);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(cy.data(), dcy, sms * 16, cudaMemcpyDeviceToHost));
                printf("%s: %2d warps x %d chains, MMA N=%3d: %.1f clk per chain step (FP warp), MMA %.1f clk each\n", name, nw, chains, n_mma,
                       (double)cy[1] / len, n_mma ? (double)cy[0] / (2.0 * mma_rounds) : 0.0);
            }
    cudaFree(dcy); cudaFree(dsink);
}

// (6) what bounds the operand feed?  The main loop of (3) with switches: `mma` 0: the consumer only waits for the stage and frees
// it (pure L2 -> shared-memory copy rate); `same` 1: every CTA reads the same window (maximal overlap of the requests in L2);
// `one` 1: a stage is ONE 36 KB bulk copy instead of 2 + 12; grid = number of CTAs (one per SM): fewer SMs -> per-SM or chip-wide limit?
__device__ __forceinline__ void tensor_g2s(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
template <int VAR>
__global__ void __launch_bounds__(128) feed_kernel(int rounds, int mma, int same, int one, const int8_t *src, long long src_bytes, long long *cycles,
                                                   const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) int8_t sm[];
    __shared__ __align__(8) uint64_t full[4], empty[4], done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 4 * STAGE; i += 128) sm[i] = (int8_t)((i * 37 + 11) % 255 - 127);
    if (tid == 0) {
        for (int s = 0; s < 4; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    const long long t0 = clock64();
    if (warp == 0 && (tid & 31) == 0) {
        long long pos = same ? 0 : ((long long)blockIdx.x * 7919 * STAGE) % (src_bytes - STAGE);
        for (int r = 0; r < rounds; r++) {
            const int s = r & 3;
            if (r >= 4) mbar_wait(&empty[s], ((r >> 2) - 1) & 1);
            mbar_expect_tx(&full[s], STAGE);
            int8_t *st = sm + s * STAGE;
            if (one == 1) bulk_g2s(st, src + pos, STAGE, &full[s]);
            else if (one >= 2) {   // the pair kernel's copies: A chunks 12 KB bulk; B chunks as ONE tensor box of 6 rows x 1 KB (2), or 6 KB bulk (3)
                const int blk = (int)(pos / (6 * 2048));   // 12 KB units of the source viewed as [block][plane 2 KB]
                for (int c = 0; c < 2; c++) {
                    bulk_g2s(st + c * A_CHUNK, src + pos + c * A_CHUNK, A_CHUNK, &full[s]);
                    if (one == 2) tensor_g2s(st + 2 * A_CHUNK + c * B_CHUNK, &tmap, (blockIdx.x & 1) * 128, 0, blk + c, &full[s]);
                    else bulk_g2s(st + 2 * A_CHUNK + c * B_CHUNK, src + pos + 2 * A_CHUNK + c * B_CHUNK, B_CHUNK, &full[s]);
                }
            } else
                for (int c = 0; c < 2; c++) {
                    bulk_g2s(st + c * A_CHUNK, src + pos + c * A_CHUNK, A_CHUNK, &full[s]);
                    for (int p = 0; p < S; p++) bulk_g2s(st + 2 * A_CHUNK + c * B_CHUNK + p * 1024, src + pos + 2 * A_CHUNK + c * B_CHUNK + p * 1024, 1024, &full[s]);
                }
            pos += STAGE;
            if (pos + STAGE > src_bytes) pos = 0;
        }
    } else if (warp == 1 && (tid & 31) == 0) {
        for (int r = 0; r < rounds; r++) {
            const int s = r & 3;
            mbar_wait(&full[s], (r >> 2) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (mma) { issue_stage<VAR>(tb, smem_u32(sm + s * STAGE), 0u); umma_commit(&empty[s]); }
            else { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory"); }
        }
        if (mma) { umma_commit(&done); mbar_wait(&done, 0); }
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

static void run_feed(const int8_t *src, long long src_bytes) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long *dcy;
    CK(cudaMalloc(&dcy, sms * 8));
    CK(cudaFuncSetAttribute(feed_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * STAGE));
    const int rounds = 20000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CUtensorMap tmap;
    {
        typedef CUresult (*encode_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                     const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        const cuuint64_t dims[3] = {256, 6, (cuuint64_t)(src_bytes / (6 * 2048))}, strides[2] = {2048, 6 * 2048};
        const cuuint32_t box[3] = {128, 6, 1}, estr[3] = {1, 1, 1};
        CUresult r = ((encode_t)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void *)src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
    }
    const int grids[4] = {sms, sms / 2, sms / 4, 8};
    for (int mma = 0; mma < 2; mma++)
        for (int same = 0; same < 1; same++)
            for (int one = 0; one < 4; one++)
                for (int gi = 0; gi < 4; gi += 3) {
                    const int g = grids[gi];
                    feed_kernel<0><<<g, 128, 4 * STAGE>>>(200, mma, same, one, src, src_bytes, dcy, tmap);
                    CK(cudaDeviceSynchronize());
                    CK(cudaEventRecord(e0));
                    feed_kernel<0><<<g, 128, 4 * STAGE>>>(rounds, mma, same, one, src, src_bytes, dcy, tmap);
                    CK(cudaEventRecord(e1));
                    CK(cudaDeviceSynchronize());
                    float ms = 0;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    long long cy0;
                    CK(cudaMemcpy(&cy0, dcy, 8, cudaMemcpyDeviceToHost));
                    printf("feed: mma %d same-window %d copies %s CTAs %3d: %.1f clk per 36 KB stage (CTA 0), %.2f TB/s L2 -> SM in total, %.1f B/clk per SM\n",
                           mma, same, one == 0 ? "2 x 12 KB + 12 x 1 KB" : one == 1 ? "1 x 36 KB" : one == 2 ? "2 x 12 KB + 2 tensor boxes (6 x 1 KB)" : "2 x 12 KB + 2 x 6 KB", g, (double)cy0 / rounds, (double)STAGE * rounds * g / (ms * 1e-3) / 1e12, (double)STAGE * rounds / (double)cy0);
                }
    cudaFree(dcy);
}

int main(int argc, char **argv) {
    if (argc > 1 && argv[1][0] == 'f') {
        const long long src_bytes = 64ll << 20;
        int8_t *src;
        CK(cudaMalloc(&src, src_bytes));
        CK(cudaMemset(src, 1, src_bytes));
        run_feed(src, src_bytes);
        cudaFree(src);
        return 0;
    }
    if (argc > 1 && argv[1][0] == 'c') {
        run_contention<0>();
        run_contention<1>();
        run_contention<2>();
        return 0;
    }
    run_check<0>();
    run_check<1>();
    const long long src_bytes = 64ll << 20;  // 64 MB: L2 resident (126 MB)
    int8_t *src;
    CK(cudaMalloc(&src, src_bytes));
    CK(cudaMemset(src, 1, src_bytes));
    run_rate<0>(0, src, src_bytes);
    run_rate<1>(0, src, src_bytes);
    run_rate<0>(1, src, src_bytes);
    run_rate<1>(1, src, src_bytes);
    run_rate_cluster<0>(src, src_bytes);
    run_rate_cluster<1>(src, src_bytes);
    cudaFree(src);
    return 0;
}
