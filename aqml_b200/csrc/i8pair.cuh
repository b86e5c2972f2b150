// Local Gaussian kernel on the 5th-generation tensor cores: FP64-accurate a.b from exact int8 slice products
// (tcgen05.mma kind::i8, accumulators in TMEM) -- the B200-native engine of calc_lk_gaussian (fkernels.f90:23-93).
//
// FP64 DMMA peaks at 37 TF on this chip; the int8 pipe does 4.1 POP/s (profiles/r01_micro_i8_umma.txt).  An FP64
// row x is split once, when the representation is packed, into S = 6 signed 8-bit digits of a per-row fixed
// point number:  x[k] = 2^e * sum_{s=0..5} q_s[k] 256^-(s+1) + r,  -128 <= q_s <= 127,  |r| <= 0.502 * 2^(e-48)
// (e: |x| 2^-e < 0.498 for the row maximum).  A digit is q = floor(u + 128/255), which keeps every remainder in
// [-128/255, 127/255) and therefore every following digit inside the int8 range (a plain round-to-nearest digit can
// be +128).  Then
//     a.b = 2^(ea+eb) * sum_g 256^-(g+2) * P_g,      P_g = sum_{i+j=g} q^a_i . q^b_j
// over the products with i + j <= 5 PLUS (3,3) -- 22 products (variant 0) -- or every product with i + j <= 6 -- 26
// (variant 1, AQML_I8_PRODUCTS=26).  Every P_g is an EXACT int32 (<= 6 products: 6 * ld * 128^2 < 2^31 for ld <= 21840).  Why (3,3):
// the dropped level i + j = 6 contains the square q_3 . q_3, which does not average out for a ~ b (exactly the
// pairs with a large kernel value): without it the error of a.b is 6 x larger (tests/test_cpu_host.py::test_digit_plane_algebra).  Errors
// left: 22 products: the four mixed pairs of level 6, <= 2^-48 of 2^(ea+eb) per column, and the digits beyond the
// 6th, 2^-49 of 2^ea per element -- measured on aSLATM rows 1e-13 .. 1e-12 relative per exp term at sigma = dmax/2.4
// (the 7 x 7-bit / 28-product scheme of round 1: 2e-13 .. 6e-13); 26 products: 1e-14 .. 3e-14.
//
// The products are issued as 8 (9) MMAs: A plane i meets the B planes 0 .. n-1, which lie one behind the other in
// the stage buffer, as ONE instruction with N = 64 n stacked rows (split at 256) whose 64-column output slices are
// the consecutive accumulators g = i .. i+n-1.  N >= 128 runs at the full UTCIMMA rate (64 clk per 128 columns; a
// single N = 64 product takes 48 clk, not 32: tools/micro/i8_umma.cu), so a stage costs ~720 (832) clk instead of
// the 28 x 48 = 1344 of round 1's one-MMA-per-product schedule.  Accumulator 6 is not touched by plane 0 (whose
// MMAs clear the others at the first stage of a tile): the epilogue zeroes it with tcgen05.st after draining it.
//
// Kernel: persistent, one CTA per SM; 128 x 64 tiles of atom pairs (A rows = TMEM lanes) handed out in raster order by an
// atomic counter; roles synchronised by mbarriers only:
//   warp 0   producer: per stage two active 16-column k-chunks of all 6 digit planes of A (2 x 12 KB, contiguous in the
//            pre-tiled slice buffer: cp.async.bulk) and of B (2 boxes of 6 planes x 1 KB through a tensor map: ONE
//            cp.async.bulk.tensor each -- six 1 KB bulk copies cost ~57 clk apiece and bounded the ring at 1150 clk per stage,
//            profiles/r02_micro_feed.txt) onto an mbarrier; 5 stages x 36 KB;
//   warp 1   issues the 8 UTCIMMA (M 128, N 64..256, K 32 = two k-chunks via the descriptor's leading byte offset, so the
//            exact 16-column block sparsity of the DMMA engine carries over) into 7 accumulators (7 x 64 = 448 TMEM columns);
//            tcgen05.commit frees the stage / signals the end.  Both role warps run their loops converged and let one
//            elected lane issue (descriptors stay in uniform registers), and the issuer tests the NEXT stage's barrier before
//            it issues this stage, so that nothing but the commit separates two stages' MMAs;
//   warp 2   tile metadata (segments, norms -- as 64-bit fixed point for the integer epilogue --, exponents, active k-chunk
//            list) two tiles ahead, triple buffered;
//   warps 3-18  epilogue (4 per TMEM lane quarter, 16 columns each): tcgen05.ld of the 7 accumulators into one exact 64-bit
//            integer per pair, TMEM handed back (the next tile's MMAs start), then -- on INTEGERS, because an FP64 instruction
//            takes ~800 clk while UTCIMMA stream (i8exp.cuh) -- d^2 = |a|^2+|b|^2-2a.b in fixed point, T = d^2/(2 sigma^2) log2 e
//            by one 64 x 64 -> high multiply, 2^-T by table + polynomial, the sigma ladder by shifts of T, the sum over the
//            atoms of a molecule pair as aligned 64-bit integers per (lane, column run), one conversion + RED.F64 per lane
//            and run; or K[s][i][j] = the assembled double for the global kernel (store epilogue).  Tiles with a NaN / Inf
//            norm, sigma sets that are no ladder and AQML_I8_EPILOGUE=fp64 use FP64 arithmetic (the round-1/2 epilogue).
// Operands are staged in the canonical K-major no-swizzle core-matrix layout (8 rows x 16 B), which is exactly how
// the slice buffer is written.
// Measured (K(test, train) 800 x 4000 molecules, 4 sigmas, profiles/r02_pair_kernel_steps.txt): 16.0 ms at the start of
// round 2 -> 11.0 ms; per 32-column stage 1435 -> 950 clk (716 = the UTCIMMA alone).
#pragma once
#include "i8digits.cuh"
#include "i8exp.cuh"
#include <cuda.h>   // CUtensorMap (encoded on the host through cudaGetDriverEntryPoint: no link against libcuda)

namespace aqml {

constexpr int I8_BM = 128, I8_BN = 64;           // tile: A rows x B rows
constexpr int I8_STAGES = 5;                     // operand ring: 4 stages in flight cover the L2 / DRAM latency of the B stream
constexpr int I8_KT_LIST = 512;                  // active k-chunk list capacity (ld <= 8192; longer rows walk all chunks)
constexpr int I8_A_CHUNK = I8_S * 2048;          // bytes of one k-chunk (16 columns) of an A tile, all planes
constexpr int I8_B_CHUNK = I8_S * 1024;
constexpr int I8_STAGE_BYTES = 2 * (I8_A_CHUNK + I8_B_CHUNK);  // 43008
constexpr int I8_EPI_WARPS = 16;
// FP64 arithmetic and the tensor pipe exclude each other on this chip (tools/micro/i8_umma.cu: dense FP64 + UTCIMMA take the
// sum of their times; FP32 overlaps fully).  Measured: overlapping the epilogue's exp / sums with the next tile's MMAs (0)
// still beats running them strictly after each other (1): 331 vs 353 ms per step.
constexpr int I8_SERIAL_EPILOGUE = 0;
// integer epilogue, end of a column run: 1 = every lane converts its own partial sum and issues its own RED.F64 (8 instructions);
// 0 = the lanes of one A molecule first add their partial sums by segmented shuffle reductions (i8_flush_int, ~100 instructions,
// 1/6 of the REDs).  The epilogue is issue bound, the L2 atomic units are not the limit (measured: see DESIGN 4.1b).
#ifndef AQML_I8_LANE_RED
#define AQML_I8_LANE_RED 1
#endif
constexpr int I8_META = 3;                       // tile metadata buffers (prepared up to two tiles ahead)
constexpr int I8_THREADS = 32 * (3 + I8_EPI_WARPS);  // producer, MMA issuer, tile metadata, 16 epilogue warps
constexpr int I8_MAXSIG = 8;
constexpr int I8_NACC = 7;                       // accumulators: levels i + j = 0 .. 6
constexpr int I8_MAXGROUPS = 16;                // element groups per launch (their tensor maps travel as kernel parameters)
constexpr int I8_MAX_LD = 21840;                 // int32 accumulators: <= 6 products each, 6 * ld * 128^2 < 2^31 (longer rows: FP64 engine)

struct I8Group {
    const int8_t *slA, *slB;      // [row block of 128][KT + 1][plane][128 rows x 16 B]; k-chunk KT is all zero
    const int *expA, *expB;       // per packed row: exponent e
    const double *normA, *normB;  // |row|^2 (FP64)
    const int *segA, *segB;       // molecule per packed row, -1 = padding
    int nrA, nrB;                 // packed rows covered by seg / norm / exp
    const unsigned *maskA, *maskB;
    int maskw;
    int KT;
    int tilesM, tilesN, tile_begin;
    int zcol, segA_off, segA_n, sym;
};

struct I8Params {
    const I8Group *groups;
    int ngroups;
    const double *isig;
    int nsig, zstride;
    double *out;
    long long ldo, slab;
    unsigned long long *stats;
    int chain, order[16], nsq[16];
    int debug;  // AQML_I8_DEBUG: MMA-thread wait timers into stats[2..6]
    int *counter;  // zeroed before the launch: next tile to hand out
    int gm;     // tile rows per raster band
    int store;  // 0: sum exp over the atoms of each molecule pair (local kernels); 1: K[s][seg_a][seg_b] = exp (global kernels)
    // per group: the B operand's slice buffer as a 3-D tensor {2 KB plane of a 128-row block, 6 planes, row block x k-chunk}; the
    // box {half a plane = 64 rows, 6 planes, 1} is one B k-chunk of a tile in ONE copy instruction (the per-instruction cost of
    // cp.async.bulk, ~57 clk, bounded the operand ring when the six 1 KB pieces were six copies: profiles/r02_micro_feed.txt)
    CUtensorMap tmB[I8_MAXGROUPS];
    const unsigned long long *exp2tab;  // 2^(-j/4096) as Q53, j = 0 .. 4095 (integer epilogue, i8exp.cuh)
};

struct I8Meta {  // everything the roles need to know about one tile
    int tm, tn, nact, KT, zcol, sym, dense;  // dense: every k-chunk 0 .. KT-1 is active (no list)
    int group;
    int special;                       // a row norm of this tile is NaN / Inf: the integer epilogue takes its FP64 side path
    // integer d^2 (ladder of sigmas, local kernel, finite norms in range): normA / normB then hold |row|^2 2^(60 - E) as 64-bit
    // integers (E: binary exponent of the tile's largest norm), T 2^48 = d^2-in-these-units * K, K = Kq 2^-(64 + kr)
    int fix, kr, eoff;                 // eoff = 13 - E:  2 a.b in the norms' units = W 2^(ea + eb + eoff)
    unsigned long long Kq;
    const int8_t *srcA, *srcB;
    int segA[I8_BM], expA[I8_BM];
    double normA[I8_BM];
    int segB[I8_BN + 1];
    double normB[I8_BN];               // |b|^2
    int expB[I8_BN];                   // e_b
    unsigned short kt[I8_KT_LIST + 2];
};

struct I8Smem {
    unsigned long long full[I8_STAGES], empty[I8_STAGES], accum, tmem_free, meta_full[I8_META], meta_empty[I8_META];
    unsigned tmem_base;
    union {
        double exptab[256];                      // 2^(j/128), 2^(j/16384)  (FP64 epilogue)
        unsigned long long exp2tab[i8x::TAB_N];  // 2^(-j/4096) 2^53  (integer epilogue)
    };
    I8Meta meta[I8_META];
};


namespace i8 {
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *b, unsigned phase) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(phase) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(unsigned long long *b, unsigned phase) {  // non-blocking
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(phase) : "memory");
    return ok != 0;
}
// bounded: a protocol bug becomes a launch failure, never a hung GPU
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned phase) {
    for (long long it = 0; !mbar_try_wait(b, phase); it++)
        if (it > 30000000ll) __trap();
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: 8 rows x 16 B core matrices; sbo = byte stride between 8-row groups, lbo = between the two
// 16-byte k-chunks of one MMA (checked bit-exactly by tools/micro/i8_umma.cu)
// one B k-chunk of a tile: box {128 x 8 B, 6 planes, 1} of the group's tensor map -> [plane][64 rows x 16 B] in shared memory
__device__ __forceinline__ void tensor_g2s(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr, unsigned lbo, unsigned sbo) {
    return (unsigned long long)((addr & 0x3FFFF) >> 4) | ((unsigned long long)(lbo >> 4) << 16) |
           ((unsigned long long)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// s32 accumulate, s8 x s8, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr unsigned idesc(int M, int N) { return (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24); }
}  // namespace i8

// exp(v), v <= 0 (or -inf, NaN): the table-based exp_tab of slatm_units.cuh with every special case and the final scaling
// done on the bit patterns -- 9 FP64 instructions.  While UTCIMMA stream through the tensor pipe the FP64 pipe delivers only
// ~0.3 warp-instructions per clock and SM (2 when the tensor pipe idles; profiles/r02_micro_contention.txt), i.e. every FP64
// instruction of the epilogue costs ~4 clk per tile and warp: the epilogue is written for few FP64 instructions.
__device__ __forceinline__ double i8_exp(double v, const double *tab) {
    const double magic = 6755399441055744.0;
    double t = fma(v, 23637.115549924776 /* 2^14 log2(e) */, magic);
    const int n = __double2loint(t);
    t -= magic;
    double r = fma(t, -4.2306346131226746e-05 /* ln2 / 2^14, upper 27 bits: t * hi is exact */, v);
    r = fma(t, -3.384964780863074e-13, r);
    double p = fma(r, 1.0 / 6.0, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double res = (tab[(n >> 7) & 127] * tab[128 + (n & 127)]) * p;   // in [1, 2.02)
    int hi = __double2hiint(res) + ((n >> 14) << 20);                     // * 2^(n >> 14): exponent field, n >> 14 >= -1022 below
    int lo = __double2loint(res);
    const unsigned vh = (unsigned)__double2hiint(v);
    if ((vh & 0x7ff00000u) == 0x7ff00000u) {              // -inf -> 0, NaN -> NaN
        const bool isnan = ((vh & 0xfffffu) | (unsigned)__double2loint(v)) != 0u;
        hi = isnan ? 0x7ff80000 : 0;
        lo = 0;
    } else if (vh > 0xC0862000u) {                         // v < -708: 0 (the reference's exp gives a denormal down to -745)
        hi = 0;
        lo = 0;
    }
    return __hiloint2double(hi, lo);
}

// sum the lanes of each A molecule (contiguous lanes, last lane run_end) and add the heads' totals into K
__device__ __noinline__ void i8_flush(double r, int lane, int run_end, bool head, double *dst) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_down_sync(0xffffffffu, r, o);
        if (lane + o <= run_end) r += u;
    }
    if (head) atomicAdd(dst, r);
}

// integer epilogue: the lanes of one A molecule (lanes head_lane .. run_end) hold acc 2^-(53 + nacc); agree on the smallest
// exponent, add the aligned integers, ONE conversion to double per molecule pair, RED into K
__device__ __noinline__ void i8_flush_int(unsigned long long acc, int nacc, int lane, int run_end, int head_lane, bool head, double *dst) {
    int nm = nacc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_down_sync(0xffffffffu, nm, o);
        if (lane + o <= run_end) nm = min(nm, u);
    }
    nm = __shfl_sync(0xffffffffu, nm, head_lane);
    const int sh = min(nacc - nm, 63);
    acc >>= sh;
    unsigned lo = (unsigned)acc, hi = (unsigned)(acc >> 32);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned ul = __shfl_down_sync(0xffffffffu, lo, o), uh = __shfl_down_sync(0xffffffffu, hi, o);
        if (lane + o <= run_end) {
            const unsigned long long t = (((unsigned long long)hi << 32) | lo) + (((unsigned long long)uh << 32) | ul);
            lo = (unsigned)t;
            hi = (unsigned)(t >> 32);
        }
    }
    const unsigned long long tot = ((unsigned long long)hi << 32) | lo;   // < 2^62: <= 16 x 32 terms of <= 2^53
    if (head && tot != 0ull && nm < i8x::N_ZERO) {                         // a sum below 2^-1022 is dropped
        const double d = (double)tot;                                      // tot >= 2^52 - few: exponent field >= 1074
        atomicAdd(dst, __hiloint2double(__double2hiint(d) - ((53 + nm) << 20), __double2loint(d)));
    }
}

// integer epilogue: this lane's partial sum acc 2^-(53 + nacc) straight into K (one conversion, one RED)
__device__ __forceinline__ void i8_lane_red(unsigned long long acc, int nacc, double *dst) {
    if (acc != 0ull && nacc < i8x::N_ZERO) {                               // a sum below 2^-1022 is dropped
        const double d = (double)acc;                                      // acc >= 2^52 - few: exponent field >= 1074
        atomicAdd(dst, __hiloint2double(__double2hiint(d) - ((53 + nacc) << 20), __double2loint(d)));
    }
}

// ---- slicing (row exponent + digit rule: i8digits.cuh) ------------------------------------------------------------
// thread = one row of a 128-row block; per 16-column chunk: 6 planes x 16 B
__global__ void __launch_bounds__(128) i8_slice_kernel(const double *x, long long ld, int nrows, const int *expo, int KT, int kt_per_block,
                                                       int8_t *sl) {
    const int rb = blockIdx.x, r = threadIdx.x, row = rb * 128 + r;
    const int kt0 = blockIdx.y * kt_per_block, kt1 = min(KT, kt0 + kt_per_block);
    const bool valid = row < nrows;
    const double sc = valid ? i8::pow2(48 - expo[row]) : 0.0;  // |e| <= 1000: exact power of two
    const double *xr = x + (long long)(valid ? row : 0) * ld;
    for (int kt = kt0; kt < kt1; kt++) {
        double t[16];
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
            const double2 v = valid ? *reinterpret_cast<const double2 *>(xr + kt * 16 + c) : make_double2(0.0, 0.0);
            t[c] = v.x;
            t[c + 1] = v.y;
        }
        i8_split16(t, sc, sl + (((long long)rb * (KT + 1) + kt) * I8_S) * 2048 + r * 16);
    }
}

// ---- peak of the pipe, measured where the bench runs (aqml_measure_i8_peak) -----------------------------------------
// One CTA per SM, one thread issues `rounds` x 2 UTCIMMA of M 128, N 256, K 32 into the two halves of TMEM from operands
// that stay in shared memory: the best rate kind::i8 reaches on this chip (profiles/r01_micro_i8_umma.txt: 128 clk per MMA).
__global__ void __launch_bounds__(128) i8_peak_kernel(int rounds) {
    __shared__ __align__(1024) int8_t sA[128 * 32], sB[256 * 32];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 32; i += 128) sA[i] = (int8_t)((i * 37 + 11) % 255 - 127);
    for (int i = tid; i < 256 * 32; i += 128) sB[i] = (int8_t)((i * 53 + 7) % 255 - 127);
    if (tid == 0) { i8::mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(i8::smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tmem_base;
    if (tid == 0) {
        const unsigned long long da = i8::smem_desc(i8::smem_u32(sA), 2048, 128), db = i8::smem_desc(i8::smem_u32(sB), 4096, 128);
        constexpr unsigned ID = i8::idesc(128, 256);
        for (int r = 0; r < rounds; r++) {
            i8::umma(tb, da, db, ID, 1u);
            i8::umma(tb + 256, da, db, ID, 1u);
        }
        i8::umma_commit(&bar);
        i8::mbar_wait(&bar, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

// ---- the pair kernel -----------------------------------------------------------------------------------------
// Persistent: one CTA per SM; tiles are handed out in raster order by an atomic counter.
// Roles: warp 0 producer, warp 1 MMA issuer, warp 2 tile metadata (segments, norms, exponents, active k-chunk list,
// two tiles ahead, triple buffered), warps 3-18 epilogue.  The epilogue first drains TMEM into registers (phase 1),
// hands TMEM back (the next tile's MMAs start while exp / sums / REDs of this tile run), and the operand ring
// keeps filling across tile boundaries.
// VAR 0: 22 products (i + j <= 5 and (3,3)); VAR 1: 26 products (i + j <= 6)
// EPI 1: exp and sums on integers (i8exp.cuh; default), EPI 0: the FP64 epilogue (AQML_I8_EPILOGUE=fp64)
template <int VAR, int EPI>
__global__ void __launch_bounds__(I8_THREADS, 1) i8_pair_kernel(const __grid_constant__ I8Params P, int total_tiles) {
    extern __shared__ __align__(1024) unsigned char i8_smem[];
    unsigned char *stages = i8_smem;
    I8Smem &S = *reinterpret_cast<I8Smem *>(i8_smem + I8_STAGES * I8_STAGE_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < I8_STAGES; s++) { i8::mbar_init(&S.full[s], 1); i8::mbar_init(&S.empty[s], 1); }
        i8::mbar_init(&S.accum, 1);
        i8::mbar_init(&S.tmem_free, I8_EPI_WARPS);
        for (int b = 0; b < I8_META; b++) { i8::mbar_init(&S.meta_full[b], 1); i8::mbar_init(&S.meta_empty[b], I8_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (EPI == 0) fill_exp_table(S.exptab, tid, I8_THREADS);
    else
        for (int i = tid; i < i8x::TAB_N; i += I8_THREADS) S.exp2tab[i] = P.exp2tab[i];
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(i8::smem_u32(&S.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = S.tmem_base;

    if (warp == 2) {
        // ---- tile metadata, one tile ahead ----
        // tiles are handed out dynamically (one atomic per tile): the tiles in flight are always consecutive in the raster
        // order, so the CTAs that share a B tile read it within one L2 residency and the tail is balanced
        int tl = 0;
        for (;;) {
            int tile = 0;
            if (lane == 0) tile = atomicAdd(P.counter, 1);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            const int b = tl % I8_META;
            I8Meta &M = S.meta[b];
            if (tl >= I8_META) i8::mbar_wait(&S.meta_empty[b], ((tl / I8_META) - 1) & 1);
            if (tile >= total_tiles) {  // sentinel: every role stops here
                if (lane == 0) M.tm = -1;
                __syncwarp();
                if (lane == 0) i8::mbar_arrive(&S.meta_full[b]);
                break;
            }
            int gi = 0;
            while (gi + 1 < P.ngroups && tile >= P.groups[gi + 1].tile_begin) gi++;
            const I8Group G = P.groups[gi];
            const int t = tile - G.tile_begin;
            const int GM = P.gm;  // tile rows per raster band (A planes stay L2 resident while B streams)
            const int band = t / (GM * G.tilesN), rem = t % (GM * G.tilesN);
            const int gm = min(GM, G.tilesM - band * GM);
            const int tm = band * GM + rem % gm, tn = rem / gm;
            if (G.sym && (tn + 1) * I8_BN <= tm * I8_BM) continue;  // entirely below the diagonal
            int any = 0, special = 0;
            for (int i = lane; i < I8_BM + I8_BN; i += 32) {
                double nrm;
                if (i < I8_BM) {
                    const int r = tm * I8_BM + i;
                    int sg = (r < G.nrA) ? G.segA[r] : -1;
                    sg = (sg >= G.segA_off && sg - G.segA_off < G.segA_n) ? sg - G.segA_off : -1;
                    M.segA[i] = sg;
                    any |= (sg >= 0);
                    M.normA[i] = nrm = (r < G.nrA) ? G.normA[r] : 0.0;
                    M.expA[i] = (r < G.nrA) ? G.expA[r] : 0;
                } else {
                    const int c = i - I8_BM, r = tn * I8_BN + c;
                    M.segB[c] = (r < G.nrB) ? G.segB[r] : -1;
                    M.normB[c] = nrm = (r < G.nrB) ? G.normB[r] : 0.0;
                    M.expB[c] = (r < G.nrB) ? G.expB[r] : 0;
                }
                special |= ((__double2hiint(nrm) & 0x7ff00000) == 0x7ff00000);
            }
            special = __any_sync(0xffffffffu, special);
            // fixed-point norms for the integer d^2 (see I8Meta::fix): decided per tile, warp-uniform
            int fix = 0, kr = 0, eoff = 0;
            unsigned long long Kq = 0ull;
            if (EPI == 1 && P.chain && !P.store && !special) {
                // integer instructions only (an FP64 instruction takes ~800 clk while the MMAs stream, and this warp must stay two tiles ahead)
                __syncwarp();
                unsigned hmax = 0u;   // norms are >= 0: their high words order like the values
                for (int i = lane; i < I8_BM + I8_BN; i += 32) hmax = max(hmax, (unsigned)__double2hiint(i < I8_BM ? M.normA[i] : M.normB[i - I8_BM]));
                hmax = __reduce_max_sync(0xffffffffu, hmax);
                const int E = (int)(hmax >> 20) - 1022;                                         // largest norm = f 2^E, f in [1/2, 1)
                const double cz = -P.isig[P.order[0] * P.zstride + G.zcol] * 1.4426950408889634;   // the one FP64 instruction
                const int eK = ((__double2hiint(cz) >> 20) & 0x7ff) + (E - 12);                 // K = cz 2^(E - 12) = m 2^(eK - 1075)
                if (E > -900 && E < 900 && cz > 0.0 && cz < 1e300 && eK >= 1022 - 63 && eK <= 1021) {   // 2^-63 <= K < 1/2
                    fix = 1; kr = 1022 - eK; eoff = 13 - E;
                    Kq = (((unsigned long long)(((unsigned)__double2hiint(cz) & 0xfffffu) | 0x100000u) << 32) | (unsigned)__double2loint(cz)) << 11;
                    for (int i = lane; i < I8_BM + I8_BN; i += 32) {
                        double *q = i < I8_BM ? &M.normA[i] : &M.normB[i - I8_BM];
                        // |row|^2 = m 2^(e - 1075)  ->  m 2^(e - 1015 - E) units of 2^(E - 60); e - 1023 < E: a left shift of <= 7 or a right shift
                        const unsigned h = (unsigned)__double2hiint(*q), l = (unsigned)__double2loint(*q);
                        const int e = (int)(h >> 20), sft = e - 1015 - E;
                        const unsigned long long m = e ? ((((unsigned long long)((h & 0xfffffu) | 0x100000u)) << 32) | l) : 0ull;   // denormal: 0
                        *reinterpret_cast<unsigned long long *>(q) = sft >= 0 ? (m << sft) : (m >> min(-sft, 63));
                    }
                }
            }
            if (!__any_sync(0xffffffffu, any)) continue;  // no wanted row in this tile (row-block sharding)
            // active 16-column k-chunks: both operands have a non-zero there (exact: a zero chunk adds 0 to a.b)
            int n = 0;
            const bool masked = G.maskA && G.maskB && G.maskw <= 32 && G.KT <= I8_KT_LIST;
            if (masked) {
                unsigned bits = 0;
                if (lane < G.maskw) {
                    const int nb64A = (G.nrA + 63) / 64;
                    unsigned am = G.maskA[(long long)(2 * tm) * G.maskw + lane];
                    if (2 * tm + 1 < nb64A) am |= G.maskA[(long long)(2 * tm + 1) * G.maskw + lane];
                    bits = am & G.maskB[(long long)tn * G.maskw + lane];
                }
                for (int w = 0; w < G.maskw; w++) {
                    unsigned bb = __shfl_sync(0xffffffffu, bits, w);
                    const int cnt = __popc(bb);
                    if (lane < cnt) {
                        for (int k = 0; k < lane; k++) bb &= bb - 1;
                        M.kt[n + lane] = (unsigned short)(w * 32 + __ffs(bb) - 1);
                    }
                    n += cnt;
                }
            } else {
                n = G.KT;  // no masks (or more chunks than the list holds): all of them, generated by the producer
            }
            if (lane == 0) {
                if (masked && (n & 1)) M.kt[n] = (unsigned short)G.KT;  // pad the last stage with the all-zero chunk
                M.nact = n; M.tm = tm; M.tn = tn; M.KT = G.KT; M.zcol = G.zcol; M.sym = G.sym; M.dense = masked ? 0 : 1; M.group = gi; M.special = special; M.fix = fix; M.kr = kr; M.eoff = eoff; M.Kq = Kq;
                M.srcA = G.slA + (long long)tm * (G.KT + 1) * I8_A_CHUNK;
                M.srcB = G.slB + (long long)(tn >> 1) * (G.KT + 1) * I8_A_CHUNK + (tn & 1) * 1024;
                if (P.stats) {  // in units of 64 x 64 x 16 k-tiles, like the DMMA engine
                    atomicAdd(P.stats, 2ull * (unsigned long long)n);
                    atomicAdd(P.stats + 1, 2ull * (unsigned long long)G.KT);
                }
            }
            __syncwarp();
            if (lane == 0) i8::mbar_arrive(&S.meta_full[b]);
            tl++;
        }
    } else if (warp == 0) {
        // ---- producer ----  (whole warp converged, one elected lane issues the copies: see the MMA issuer)
        {
            unsigned leader;
            asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(leader));
            int s = 0;           // ring slot of the next stage (continues across tiles) and the parity its `empty` barrier is waited with
            unsigned par = 1;    // first round: nothing to wait for (a fresh barrier passes a wait on parity 1)
            for (int tl = 0;; tl++) {
                const int b = tl % I8_META;
                const I8Meta &M = S.meta[b];
                i8::mbar_wait(&S.meta_full[b], (tl / I8_META) & 1);
                if (M.tm < 0) break;
                const int nst = (M.nact + 1) >> 1;
                const int8_t *srcA = M.srcA;
                const CUtensorMap *tmB = &P.tmB[M.group];
                const int bx = (M.tn & 1) * 128, bz = (M.tn >> 1) * (M.KT + 1), dense = M.dense, KT = M.KT;
                for (int k = 0; k < nst; k++) {
                    i8::mbar_wait(&S.empty[s], par);
                    unsigned char *st = stages + s * I8_STAGE_BYTES;
                    const int kt0 = dense ? min(2 * k, KT) : (int)M.kt[2 * k], kt1 = dense ? min(2 * k + 1, KT) : (int)M.kt[2 * k + 1];
                    if (leader) {
                        i8::mbar_expect_tx(&S.full[s], I8_STAGE_BYTES);
                        i8::bulk_g2s(st, srcA + (long long)kt0 * I8_A_CHUNK, I8_A_CHUNK, &S.full[s]);
                        i8::tensor_g2s(st + 2 * I8_A_CHUNK, tmB, bx, 0, bz + kt0, &S.full[s]);
                        i8::bulk_g2s(st + I8_A_CHUNK, srcA + (long long)kt1 * I8_A_CHUNK, I8_A_CHUNK, &S.full[s]);
                        i8::tensor_g2s(st + 2 * I8_A_CHUNK + I8_B_CHUNK, tmB, bx, 0, bz + kt1, &S.full[s]);
                    }
                    __syncwarp();
                    if (++s == I8_STAGES) { s = 0; par ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ----
        // The whole warp walks the loop converged (descriptors, waits: warp-uniform values the compiler keeps in uniform
        // registers) and ONE elected lane issues: ~3 instructions per UTCIMMA.  With the loop inside `if (lane == 0)` every MMA
        // carried a ~15-instruction register-to-uniform-register loop; once the epilogue warps of this scheduler were issue
        // bound (integer epilogue) the issuer could not keep the tensor pipe fed (1120 instead of 750 clk per stage).
        {
            // A plane i against the stacked B planes j0 .. j0 + n - 1 (N = 64 n) -> accumulators i + j0 .. i + j0 + n - 1
            constexpr int NMMA = VAR ? 9 : 8;
            // VAR 0: plane 0 first (it clears the accumulators at k = 0), then the short MMAs, the two N = 256 ones last: the tensor
            // pipe's queue then holds the most work while the issuer does its bookkeeping between two stages
            constexpr int mi[9] = {0, 0, VAR ? 1 : 5, VAR ? 1 : 4, VAR ? 2 : 1, VAR ? 2 : 1, VAR ? 3 : 2, VAR ? 4 : 3, 5};
            constexpr int mj[9] = {0, 4, 0, VAR ? 4 : 0, VAR ? 0 : 3, VAR ? 3 : 0, 0, 0, 0};
            constexpr int mn[9] = {4, 2, VAR ? 4 : 1, 2, VAR ? 3 : 2, VAR ? 2 : 3, 4, VAR ? 3 : 4, 2};
            unsigned leader;
            asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(leader));
            const unsigned stage0 = i8::smem_u32(stages);
            int it = 0, s = 0;       // stages issued so far; ring slot and parity of the next one (no division in the loop)
            unsigned par = 0;
            bool ready = false;      // the next stage's operands were already seen complete (looked ahead while the previous MMAs ran)
            long long w_meta = 0, w_tmem = 0, w_full = 0;
            const long long t_begin = clock64();
            for (int tl = 0;; tl++) {
                const int b = tl % I8_META;
                long long c0 = P.debug ? clock64() : 0;
                i8::mbar_wait(&S.meta_full[b], (tl / I8_META) & 1);
                if (P.debug) w_meta += clock64() - c0;
                if (S.meta[b].tm < 0) {
                    if (P.debug && leader) {
                        atomicAdd(P.stats + 2, (unsigned long long)(clock64() - t_begin));
                        atomicAdd(P.stats + 3, (unsigned long long)w_meta);
                        atomicAdd(P.stats + 4, (unsigned long long)w_tmem);
                        atomicAdd(P.stats + 5, (unsigned long long)w_full);
                        atomicAdd(P.stats + 6, (unsigned long long)it);
                    }
                    break;
                }
                const int nst = (S.meta[b].nact + 1) >> 1;
                if (P.debug) c0 = clock64();
                i8::mbar_wait(&S.tmem_free, tl & 1);  // accumulator 6 cleared (tl = 0) / the previous tile's accumulators are in registers
                if (P.debug) w_tmem += clock64() - c0;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int k = 0; k < nst; k++, it++) {
                    if (!ready) {
                        if (P.debug) c0 = clock64();
                        i8::mbar_wait(&S.full[s], par);
                        if (P.debug) w_full += clock64() - c0;
                    }
                    const unsigned a0 = stage0 + s * I8_STAGE_BYTES, b0 = a0 + 2 * I8_A_CHUNK;
                    const unsigned long long da = i8::smem_desc(a0, I8_A_CHUNK, 128), db = i8::smem_desc(b0, I8_B_CHUNK, 128);
                    const int s0 = s;
                    if (++s == I8_STAGES) { s = 0; par ^= 1u; }
                    // look ahead: is the NEXT stage (of this or the next tile: the ring is continuous) already complete?  Then nothing
                    // but the commit stands between this stage's last MMA and the next stage's first one -- the tensor pipe's queue
                    // is short, and a wait + descriptor arithmetic between the stages showed up as ~180 idle clk per stage
                    ready = i8::mbar_test_wait(&S.full[s], par);
                    if (leader) {
#pragma unroll
                        for (int m = 0; m < NMMA; m++)  // plane 0 comes first and clears accumulators 0 .. 5 at k = 0 (6: the epilogue)
                            i8::umma(tb + (mi[m] + mj[m]) * I8_BN, da + (unsigned long long)(mi[m] * 2048 >> 4),
                                     db + (unsigned long long)(mj[m] * 1024 >> 4), i8::idesc(I8_BM, I8_BN * mn[m]), (k > 0 || mi[m] > 0) ? 1u : 0u);
                        i8::umma_commit(&S.empty[s0]);  // stage reusable once these MMAs have read it
                    }
                    __syncwarp();
                }
                if (leader) {
                    if (nst > 0) i8::umma_commit(&S.accum);
                    else i8::mbar_arrive(&S.accum);
                }
                __syncwarp();
            }
        }
    } else {
        // ---- epilogue: 16 warps; warp (q, h) owns TMEM lanes / tile rows 32q .. 32q+31 and columns 16h .. 16h+15 ----
        // (many warps with little state each: the exp / Horner chains are latency bound, not issue bound)
        constexpr int EC = I8_BN / 4;  // 16 columns per warp
        const int q = warp & 3, cbase = ((warp - 3) >> 2) * EC, row = q * 32 + lane;
        const unsigned trow = tb + ((unsigned)(q * 32) << 16) + cbase;
        auto clear_acc6 = [&]() {  // plane 0's MMAs never touch level 6: it accumulates from zero
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
                         ::"r"(trow + 6 * I8_BN), "r"(0u) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        };
        clear_acc6();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) i8::mbar_arrive(&S.tmem_free);
        for (int tl = 0;; tl++) {
            const int b = tl % I8_META;
            const I8Meta &M = S.meta[b];
            i8::mbar_wait(&S.meta_full[b], (tl / I8_META) & 1);
            if (M.tm < 0) break;
            const int nst = (M.nact + 1) >> 1;
            const int sr = M.segA[row];
            // lanes of one molecule are contiguous: head lane and last lane of my run
            const int prev = __shfl_up_sync(0xffffffffu, sr, 1);
            const bool head = ((lane == 0) || (prev != sr)) && sr >= 0;
            const unsigned heads = __ballot_sync(0xffffffffu, (lane == 0) || (prev != sr));
            const unsigned above = (lane == 31) ? 0u : (heads >> (lane + 1));
            const int run_end = above ? (lane + __ffs(above) - 1) : 31;
            const int head_lane = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
            const double na = M.normA[row];
            const int ea = M.expA[row];
            const int growA = M.tm * I8_BM + row;
            const unsigned lastmask = __ballot_sync(0xffffffffu, lane < EC && (lane == EC - 1 || M.segB[cbase + lane + 1] != M.segB[cbase + lane]));
            const long long e0 = clock64();
            i8::mbar_wait(&S.accum, tl & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long e1 = clock64();
            // phase 1: drain my 16 columns of the 7 accumulators into ONE exact 64-bit integer per pair (integer pipe only), then
            // hand TMEM back: the next tile's MMAs run while everything below happens.
            //   sum_g P_g 256^-g = 2^-32 W + (dropped fraction),  W = P0 2^32 + P1 2^24 + P2 2^16 + P3 2^8 + P4 + floor((P5 2^8 + P6) / 2^16):
            // levels 5 and 6 come first (their sum is floored once, exactly as a 128-bit sum shifted right would be), and the loads
            // go out two at a time so that one TMEM round trip covers two accumulators.
            long long Wacc[EC];
#pragma unroll
            for (int j = 0; j < EC; j++) Wacc[j] = 0;
            if (nst > 0) {
                auto ld16 = [&](unsigned (&v)[EC], int g) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                 : "r"(trow + g * I8_BN));
                };
                unsigned va[EC], vb[EC];
                ld16(va, 5);
                ld16(vb, 6);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < EC; j++) Wacc[j] = ((((long long)(int)va[j]) << 8) + (long long)(int)vb[j]) >> 16;
                ld16(va, 0);
                ld16(vb, 1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < EC; j++) Wacc[j] += (((long long)(int)va[j]) << 32) + (((long long)(int)vb[j]) << 24);
                ld16(va, 2);
                ld16(vb, 3);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < EC; j++) Wacc[j] += (((long long)(int)va[j]) << 16) + (((long long)(int)vb[j]) << 8);
                ld16(va, 4);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < EC; j++) Wacc[j] += (long long)(int)va[j];
                clear_acc6();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (!I8_SERIAL_EPILOGUE && lane == 0) i8::mbar_arrive(&S.tmem_free);
            const long long e2 = clock64();
            if (P.debug & 2) {  // measurement only: no exp / sums (how much of the tile time is the FP64 epilogue?)
                __syncwarp();
                if (lane == 0) i8::mbar_arrive(&S.meta_empty[b]);
                continue;
            }
            // integer epilogue, ladder of sigmas, local kernel: V[j] = T of the widest sigma as Q16.48; per sigma (a left shift of V)
            // a term (M, n) per pair, column runs added as aligned 64-bit integers, one RED per lane and run
            auto chain_loop = [&](unsigned long long (&V)[EC]) {
#pragma unroll 1
                for (int si = 0; si < P.nsig; si++) {
                    double *outS = P.out + (long long)P.order[si] * P.slab + (long long)max(sr, 0) * P.ldo;
                    const int sq = si ? P.nsq[si] : 0;
                    unsigned long long acc = 0ull;
                    int nacc = 4095;
#pragma unroll
                    for (int j = 0; j < EC; j++) {
                        V[j] = i8x::fix_shl(V[j], sq);
                        unsigned long long Mt;
                        int nt;
                        i8x::exp2neg(V[j], S.exp2tab, Mt, nt);
                        i8x::add_term(acc, nacc, Mt, nt);
                        if ((lastmask >> j) & 1u) {  // warp-uniform: column cbase + j closes a molecule of B
                            const int sc = M.segB[cbase + j];
                            if (sc >= 0) {
                                if (AQML_I8_LANE_RED) { if (sr >= 0) i8_lane_red(acc, nacc, outS + sc); }
                                else i8_flush_int(acc, nacc, lane, run_end, head_lane, head, outS + sc);
                            }
                            acc = 0ull;
                            nacc = 4095;
                        }
                    }
                }
            };
            if (EPI == 1 && M.fix) {
                // d^2 and T on integers too (no FP64 instruction in this tile's epilogue except the conversions at the REDs): the norms
                // arrive as 64-bit integers in units of 2^(E - 60), 2 a.b = W 2^(ea + eb - 47) is shifted into the same units, and
                // T 2^48 = d^2 K by one 64 x 64 -> high multiply (K = Kq 2^-(64 + kr), prepared per tile by the metadata warp)
                unsigned long long V[EC];
                const long long NA = reinterpret_cast<const long long *>(M.normA)[row];
                const int eaE = ea + M.eoff, kr = M.kr;
                const unsigned long long Kq = M.Kq;
#pragma unroll
                for (int j = 0; j < EC; j++) {
                    const int col = cbase + j;
                    const long long W = Wacc[j];
                    const int sh = eaE + M.expB[col];
                    const long long Ws = (long long)((unsigned long long)W << i8x::clamp63(sh)) >> i8x::clamp63(-sh);
                    long long D = NA + reinterpret_cast<const long long *>(M.normB)[col] - Ws;
                    D = D < 0 ? 0 : D;                                                          // rounding noise of a.b for a = b
                    const unsigned long long F = __umul64hi((unsigned long long)D, Kq) >> kr;
                    const bool out = (M.sym && M.tn * I8_BN + col <= growA) || sr < 0 || M.segB[col] < 0;   // pairs that do not count
                    V[j] = out ? ~0ull : F;
                }
                chain_loop(V);
                __syncwarp();
                if (I8_SERIAL_EPILOGUE && lane == 0) i8::mbar_arrive(&S.tmem_free);
                if (P.debug && lane == 0) {
                    const long long e3 = clock64();
                    if (warp == 3) {
                        atomicAdd(P.stats + 7, (unsigned long long)(e1 - e0));
                        atomicAdd(P.stats + 8, (unsigned long long)(e2 - e1));
                        atomicAdd(P.stats + 9, (unsigned long long)(e3 - e2));
                    }
                    atomicAdd(P.stats + 16 + (warp - 3), (unsigned long long)(e3 - e2));
                    atomicAdd(P.stats + 32 + (warp - 3), (unsigned long long)(e1 - e0));
                }
                if (lane == 0) i8::mbar_arrive(&S.meta_empty[b]);
                continue;
            }
            // d^2 of my 16 pairs: a.b = 2^(ea + eb - 64) V, V = hi 2^32 + lo; W = V >> 16 has <= 62 bits, ONE conversion to double
            // (rounding 2^-53 of a.b, as in an FP64 dot product), one FMA with the power of two built from the exponents
            double x[EC];
#pragma unroll
            for (int j = 0; j < EC; j++) {
                const int col = cbase + j;
                const long long W = Wacc[j];
                const int e2 = max(-1022, min(1023, ea + M.expB[col] - 47));                 // -2 a.b = W * -2^(ea + eb - 47)
                const double sc = __hiloint2double((int)(0x80000000u | ((unsigned)(e2 + 1023) << 20)), 0);
                double xx = fma((double)W, sc, na + M.normB[col]);                           // |a|^2 + |b|^2 - 2 a.b
                int xh = __double2hiint(xx), xl = __double2loint(xx);
                if (xh < 0 && (xh & 0x7ff00000) != 0x7ff00000) { xh = 0; xl = 0; }           // clamp rounding noise, keep NaN
                // pairs that do not count: on / below the diagonal of a symmetric kernel, padding rows / columns
                if ((M.sym && M.tn * I8_BN + col <= growA) || sr < 0 || M.segB[col] < 0) { xh = 0x7ff00000; xl = 0; }
                x[j] = __hiloint2double(xh, xl);
            }
            if (EPI == 1) {
                // phase 2 on integers (i8exp.cuh): per sigma a term (M, n) per pair from T = d^2 / (2 sigma^2) log2(e) as Q16.48 --
                // ladder: T of the widest sigma once, the next sigma is a left shift; otherwise one FP64 multiply per (pair, sigma).
                // Local kernels: column runs added as aligned 64-bit integers, row runs by i8_flush_int; global kernels: the
                // double is assembled and stored.  The three cases are separate straight-line loops (no per-pair branches).
                constexpr double LOG2E = 1.4426950408889634;
                auto fixT = [](double T) { return i8x::fix_from_bits((unsigned)__double2hiint(T) & 0x7fffffffu, (unsigned)__double2loint(T)); };
                if (M.special) {
                    // a NaN / Inf row norm in this tile (non-finite input): plain FP64 per pair, so that NaN and exp(-Inf) = 0 reach K
                    // exactly as in the reference (fkernels.f90:84,351); the integer path would turn NaN into 0
#pragma unroll 1
                    for (int s = 0; s < P.nsig; s++) {
                        double *outS = P.out + (long long)s * P.slab + (long long)max(sr, 0) * P.ldo;
                        const double f = P.isig[s * P.zstride + M.zcol];
#pragma unroll
                        for (int j = 0; j < EC; j++) {
                            const int sc = M.segB[cbase + j];
                            if (sr < 0 || sc < 0) continue;
                            const bool counts = !(M.sym && M.tn * I8_BN + cbase + j <= growA);
                            const double e = counts ? exp(x[j] * f) : 0.0;
                            if (P.store) outS[sc] = e;
                            else if (counts) atomicAdd(outS + sc, e);
                        }
                    }
                } else if (P.store) {
#pragma unroll 1
                    for (int si = 0; si < P.nsig; si++) {
                        const int s = P.chain ? P.order[si] : si;
                        double *outS = P.out + (long long)s * P.slab + (long long)max(sr, 0) * P.ldo;
                        // ladder: T = d^2 * c of the widest sigma times 2^(squarings so far) -- exact, so one multiply serves here too
                        const double cs = -P.isig[s * P.zstride + M.zcol] * LOG2E;
#pragma unroll
                        for (int j = 0; j < EC; j++) {
                            unsigned long long Mt;
                            int nt;
                            i8x::exp2neg(fixT(x[j] * cs), S.exp2tab, Mt, nt);
                            const int sc = M.segB[cbase + j];
                            if (sr >= 0 && sc >= 0) outS[sc] = __longlong_as_double((long long)i8x::term_bits(Mt, nt));
                        }
                    }
                } else if (P.chain) {
                    unsigned long long V[EC];
                    const double c0 = -P.isig[P.order[0] * P.zstride + M.zcol] * LOG2E;
#pragma unroll
                    for (int j = 0; j < EC; j++) V[j] = fixT(x[j] * c0);
                    chain_loop(V);
                } else {
#pragma unroll 1
                    for (int si = 0; si < P.nsig; si++) {
                        double *outS = P.out + (long long)si * P.slab + (long long)max(sr, 0) * P.ldo;
                        const double cs = -P.isig[si * P.zstride + M.zcol] * LOG2E;
                        unsigned long long acc = 0ull;
                        int nacc = 4095;
#pragma unroll
                        for (int j = 0; j < EC; j++) {
                            unsigned long long Mt;
                            int nt;
                            i8x::exp2neg(fixT(x[j] * cs), S.exp2tab, Mt, nt);
                            i8x::add_term(acc, nacc, Mt, nt);
                            if ((lastmask >> j) & 1u) {
                                const int sc = M.segB[cbase + j];
                                if (sc >= 0) {
                                    if (AQML_I8_LANE_RED) { if (sr >= 0) i8_lane_red(acc, nacc, outS + sc); }
                                    else i8_flush_int(acc, nacc, lane, run_end, head_lane, head, outS + sc);
                                }
                                acc = 0ull;
                                nacc = 4095;
                            }
                        }
                    }
                }
            } else {
                // phase 2: per sigma exp (chain: squarings of the previous sigma's values, in place), then the sums over
                // molecule pairs: column runs sequentially in registers, row runs by a segmented shuffle reduction, one RED
                // per head lane
#pragma unroll 1
                for (int si = 0; si < P.nsig; si++) {
                    const int s = P.chain ? P.order[si] : si;
                    double *outS = P.out + (long long)s * P.slab + (long long)max(sr, 0) * P.ldo;
                    double run = 0.0;
                    if (P.store) {
                        // global kernel (fkernels.f90:344-355): every pair is one element of K; 16 consecutive columns per lane
                        if (!P.chain || si == 0) {
                            const double f = P.isig[s * P.zstride + M.zcol];
#pragma unroll
                            for (int j = 0; j < EC; j++) {
                                const double e = i8_exp(x[j] * f, S.exptab);
                                if (P.chain) x[j] = e;
                                const int sc = M.segB[cbase + j];
                                if (sr >= 0 && sc >= 0) outS[sc] = e;
                            }
                        } else {
                            for (int k = 0; k < P.nsq[si]; k++)
#pragma unroll
                                for (int j = 0; j < EC; j++) x[j] *= x[j];
#pragma unroll
                            for (int j = 0; j < EC; j++) {
                                const int sc = M.segB[cbase + j];
                                if (sr >= 0 && sc >= 0) outS[sc] = x[j];
                            }
                        }
                    } else if (P.chain) {
                        if (si == 0) {
                            const double f = P.isig[s * P.zstride + M.zcol];
#pragma unroll
                            for (int j = 0; j < EC; j++) x[j] = i8_exp(x[j] * f, S.exptab);  // +inf * f = -inf -> 0
                        } else {
                            for (int k = 0; k < P.nsq[si]; k++)
#pragma unroll
                                for (int j = 0; j < EC; j++) x[j] *= x[j];
                        }
#pragma unroll
                        for (int j = 0; j < EC; j++) {
                            run += x[j];
                            if ((lastmask >> j) & 1u) {  // warp-uniform: column cbase + j closes a molecule of B
                                const int sc = M.segB[cbase + j];
                                if (sc >= 0) i8_flush(run, lane, run_end, head, outS + sc);
                                run = 0.0;
                            }
                        }
                    } else {
                        const double f = P.isig[s * P.zstride + M.zcol];
#pragma unroll
                        for (int j = 0; j < EC; j++) {
                            run += i8_exp(x[j] * f, S.exptab);
                            if ((lastmask >> j) & 1u) {
                                const int sc = M.segB[cbase + j];
                                if (sc >= 0) i8_flush(run, lane, run_end, head, outS + sc);
                                run = 0.0;
                            }
                        }
                    }
                }
            }  // EPI
            __syncwarp();
            if (I8_SERIAL_EPILOGUE && lane == 0) i8::mbar_arrive(&S.tmem_free);
            if (P.debug && lane == 0) {
                const long long e3 = clock64();
                if (warp == 3) {
                    atomicAdd(P.stats + 7, (unsigned long long)(e1 - e0));
                    atomicAdd(P.stats + 8, (unsigned long long)(e2 - e1));
                    atomicAdd(P.stats + 9, (unsigned long long)(e3 - e2));
                }
                atomicAdd(P.stats + 16 + (warp - 3), (unsigned long long)(e3 - e2));   // per epilogue warp: exp + sums, wait
                atomicAdd(P.stats + 32 + (warp - 3), (unsigned long long)(e1 - e0));
            }
            if (lane == 0) i8::mbar_arrive(&S.meta_empty[b]);  // this tile's metadata buffer may be refilled
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

}  // namespace aqml
