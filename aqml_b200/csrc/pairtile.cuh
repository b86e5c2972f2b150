// Pair-tile engine: every kernel-matrix / distance routine of the hot path is one instance of
//
//     for a 64x64 tile of (row i of A, row j of B):  v_ij = reduce_k f(A[i,k], B[j,k])
//     then an epilogue on v_ij
//
//   MODE_DOT : v = max(|a|^2 + |b|^2 - 2 a.b, 0); the cross term runs on the FP64 tensor path
//              (DMMA.8x8x4), operands staged through a 3-deep cp.async shared-memory pipeline; only
//              the k-tiles that are non-zero in both operand blocks are visited (exact block sparsity).
//   MODE_L1  : v = sum_k |a_k - b_k| on the FP64 CUDA cores (not a contraction), same staging.
//
//   EPI_EXP_STORE  : K[s][i][j] = exp(v * isig[s])          global Gaussian/Laplacian, all sigmas in one pass
//   EPI_DIST_STORE : D[i][j]   = sqrt(v) | v                l2 / l1 distance matrices
//   EPI_DOT_STORE  : G[i][j]   = a.b                        linear kernel (no norm transform)
//   EPI_MAX        : out[slot] = max(out[slot], v)          kernel-width heuristic (never stores D)
//   EPI_LOCAL      : K[s][mol(i)][mol(j)] += exp(v*isig[s][z_i])   local kernels: per-warp staged,
//                    segmented (molecule-pair) reduction, then one FP64 RED per molecule pair.
//
// Work is described by a small table of PairGroups (one per element for the element-matched
// local kernels -- the block-sparse mask of the design: only same-element atom pairs are ever
// contracted, and only over the columns in that element's support).
#pragma once
#include "common.cuh"

namespace aqml {

// CTA tile (16*MI) x 64, 4 warps (2 x 2, warp tile (8*MI) x 32), 3-stage pipeline.  MI=8: 128x64 tile,
// 92 KB smem, <=240 regs -> 2 CTAs/SM.  MI=4: 64x64 tile, 61 KB smem, <=168 regs -> 3 CTAs/SM.  Several
// resident CTAs matter because the epilogue (exp + segmented reduction) and the per-k-tile barrier leave
// the FP64 units idle: another CTA's DMMA stream fills those gaps (measured: 76% busy with one 128x128 CTA,
// 84% with 2 x (128x64), 84% on half the work with 3 x (64x64); profiles/history).
#ifndef AQML_MI
#define AQML_MI 4
#endif
constexpr int MI = AQML_MI;        // 8x8 MMA row blocks per warp: warp tile (8*MI) x 32
constexpr int WROWS = 8 * MI;
constexpr int BM = 2 * WROWS, BN = 64, BK = 16;
#ifndef AQML_CTAS
#define AQML_CTAS ((AQML_MI == 8) ? 2 : 3)
#endif
constexpr int CTAS_PER_SM = AQML_CTAS;
constexpr int PITCH = BK + 4;  // doubles; (PITCH mod 16) == 4 -> conflict-free 8x4 fragment loads
#ifndef AQML_STAGES
#define AQML_STAGES 3
#endif
constexpr int STAGES = AQML_STAGES;
constexpr int WARPS_N = BN / 32;
constexpr int NT = 64 * WARPS_N;  // 2 (M) x WARPS_N (N) warps
constexpr int STAGE_DOUBLES = (BM + BN) * PITCH;
constexpr int PIPE_BYTES = STAGES * STAGE_DOUBLES * 8;  // 61440 for the default 64x64 tile, 3 stages
constexpr int EPITCH = 40;     // per-warp staging pitch (doubles) for the segmented epilogue
static_assert((NT / 32) * WROWS * EPITCH * 8 <= PIPE_BYTES, "epilogue staging must fit in the pipeline buffers");

enum { MODE_DOT = 0, MODE_L1 = 1 };
enum { EPI_EXP_STORE = 0, EPI_DIST_STORE = 1, EPI_MAX = 2, EPI_LOCAL = 3, EPI_DOT_STORE = 4 };

struct PairGroup {
    const double *A, *B;          // packed operands, row pitch ld (multiple of BK), rows padded to 128
    const double *normA, *normB;  // |row|^2 (MODE_DOT)
    const int *segA, *segB;       // output index per packed row, -1 = padding; null -> identity < nA/nB
    const int *zA;                // sigma column per packed A row; null -> zcol
    long long ld;
    int nA, nB;                   // valid rows when seg* is null
    int tilesM, tilesN;
    int tile_begin;               // first flattened tile id of this group
    int zcol;
    int out_slot;                 // EPI_MAX
    int segA_off, segA_n;         // keep rows with segA_off <= seg < segA_off+segA_n, output row = seg - segA_off
    int sym;                      // A == B: skip tiles below the diagonal (EPI_MAX); EPI_LOCAL additionally keeps only
                                  // pairs with packed column > packed row (the caller mirrors K and adds the i == j terms)
    // Feature-space block sparsity: one bit per 16-column k-tile and per 64-row block, set when the block
    // holds any non-zero.  A k-tile whose A-block or B-block is all zero contributes exactly 0 to a.b (and to
    // sum|a-b| when both are zero), so it is skipped -- bit-identical results, fewer DMMAs.  null = dense.
    const unsigned *maskA, *maskB;
    int maskw;                    // 32-bit words per 64-row block
    // EPI_MAX pruning (triangle inequality): boundX[r] >= distance of packed row r (and of every later row of its
    // tile -- rows are sorted by it) to a common reference point.  No pair of a tile can reach bound_min when
    // boundA[first row] + boundB[first row] < bound_min, a value some pair is already known to attain.  null = off.
    const double *boundA, *boundB;
    double bound_min;
};

struct PairParams {
    const PairGroup *groups;
    int ngroups;
    const double *isig;  // device (nsig, zstride)
    int nsig, zstride;
    double *out;
    long long ldo, slab;  // row pitch, per-sigma stride (elements)
    int *sm_arrive;       // [>= #SMs] zeroed per launch, or null: stagger the two resident CTAs of an SM
    unsigned long long *stats;  // [2] executed / dense k-tile counts (summed over tiles), or null
    // Sigma chain: when every sigma is an exact power-of-two multiple of the widest one (the usual
    // coefficient ladder 0.5, 1, 2, 4 ...), exp(v*isig[s]) = exp(v*isig[base])^(4^k) exactly in real
    // arithmetic: one exp for the widest sigma, then squarings.  Every squaring doubles the relative error of the one exp,
    // so the chain is used only while the squarings ADD UP to <= 8 over the whole ladder (2^8 ulp = 3e-14 in FP64, ~1e-11
    // in the int8 engine's FP32-pair arithmetic); wider ladders get one exp per sigma.
    const double *isig_host;  // host copy of isig, valid during the call (set by plan_sigma_chain; the int8 engine derives its
                              // fixed-point constants from it)
    int chain;            // 1: use order/nsq below
    int order[16];        // evaluation order: widest sigma first
    int nsq[16];          // squarings that take order[i-1]'s values to order[i]'s
};

constexpr int KT_LIST_MAX = 1024;  // active-k-tile list capacity (ld <= 16384 columns)

template <int MODE, int EPI>
__global__ void __launch_bounds__(NT, CTAS_PER_SM) pair_tile_kernel(const PairParams P) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int sSegA[BM], sSegB[BN], sZA[BM];
    __shared__ double sNA[BM], sNB[BN];
    __shared__ double sRed[NT / 32];
    __shared__ unsigned short sKt[KT_LIST_MAX];
    __shared__ int sNact;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int g = lane >> 2, t4 = lane & 3;

    // ---- locate group and tile -------------------------------------------------------------
    int gi = 0;
    while (gi + 1 < P.ngroups && (int)blockIdx.x >= P.groups[gi + 1].tile_begin) gi++;
    const PairGroup G = P.groups[gi];
    const int t = blockIdx.x - G.tile_begin;
    constexpr int GM = 32;  // tile rows per raster band: the A panels (32 x 64 rows) stay L2-resident while B streams
    const int band = t / (GM * G.tilesN), rem = t % (GM * G.tilesN);
    const int gm = min(GM, G.tilesM - band * GM);
    const int tm = band * GM + rem % gm, tn = rem / gm;
    if ((EPI == EPI_MAX || EPI == EPI_LOCAL) && G.sym && (tn + 1) * BN <= tm * BM) return;  // tile entirely below the diagonal
    if (EPI == EPI_MAX && G.boundA && G.boundA[tm * BM] + G.boundB[tn * BN] < G.bound_min) return;  // cannot hold the maximum

    // ---- de-phase the two CTAs that share an SM --------------------------------------------------
    // All tiles of a group cost the same, so the two resident CTAs would run in lockstep and reach
    // their (tensor-idle) epilogues together.  The second CTA to arrive on each SM waits half a tile
    // once; from then on one CTA's epilogue / barrier bubbles overlap the other's DMMA stream.
    if (P.sm_arrive) {
        if (tid == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const int k = atomicAdd(P.sm_arrive + smid, 1);
            if (k >= 1 && k < CTAS_PER_SM) {
                // one full tile of one CTA takes ~ (active k-tiles) * MI*16*16 * CTAS_PER_SM cycles of DMMA issue
                const long long tile = (long long)(G.ld / BK) * (MI * 256) * CTAS_PER_SM / (G.maskA ? 2 : 1);
                const long long wait = tile * k / CTAS_PER_SM;
                const long long t0 = clock64();
                while (clock64() - t0 < wait) __nanosleep(1000);
            }
        }
        __syncthreads();
    }

    // ---- tile metadata ---------------------------------------------------------------------
    int any = 0;
    for (int i = tid; i < BM + BN; i += NT) {
        if (i < BM) {
            int r = tm * BM + i;
            int s = G.segA ? G.segA[r] : (r < G.nA ? r : -1);
            s = (s >= G.segA_off && s - G.segA_off < G.segA_n) ? s - G.segA_off : -1;
            sSegA[i] = s;
            any |= (s >= 0);
            sNA[i] = G.normA ? G.normA[r] : 0.0;
            sZA[i] = G.zA ? G.zA[r] : G.zcol;
        } else {
            int c = i - BM, r = tn * BN + c;
            sSegB[c] = G.segB ? G.segB[r] : (r < G.nB ? r : -1);
            sNB[c] = G.normB ? G.normB[r] : 0.0;
        }
    }
    if (!__syncthreads_or(any)) return;  // no row of this tile is wanted (row-block sharding)

    // ---- main loop: 4-stage cp.async pipeline ------------------------------------------------
    const double *gA = G.A + (long long)tm * BM * G.ld;
    const double *gB = G.B + (long long)tn * BN * G.ld;
    const int KT = (int)(G.ld / BK);

    // ---- active k-tile list ------------------------------------------------------------------------
    const bool use_list = (G.maskA != nullptr) && (G.maskB != nullptr) && KT <= KT_LIST_MAX && G.maskw <= 32;
    int nact = KT;
    if (use_list) {
        if (warp == 0) {
            unsigned bits = 0;
            if (lane < G.maskw) {
                unsigned a0 = 0;
#pragma unroll
                for (int h = 0; h < BM / 64; h++) a0 |= G.maskA[(long long)(tm * (BM / 64) + h) * G.maskw + lane];
                const unsigned b = G.maskB[(long long)tn * (BN / 64) * G.maskw + lane];
                bits = (MODE == MODE_DOT) ? (a0 & b) : (a0 | b);
                const int hi = KT - lane * 32;  // clip bits beyond KT
                if (hi < 32) bits &= (hi <= 0) ? 0u : ((1u << hi) - 1u);
            }
            int cnt = __popc(bits), incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            int off = incl - cnt;
            while (bits) {
                int bpos = __ffs(bits) - 1;
                bits &= bits - 1;
                sKt[off++] = (unsigned short)(lane * 32 + bpos);
            }
            if (lane == 31) sNact = incl;
        }
        __syncthreads();
        nact = sNact;
    }
    if (P.stats && tid == 0) {
        atomicAdd(P.stats, (unsigned long long)nact);
        atomicAdd(P.stats + 1, (unsigned long long)KT);
    }

    // one quarter of a stage per call: the copies are spread over the four k-steps of the k-tile being
    // computed instead of being issued back to back behind the barrier (LDGSTS issue throttling)
    constexpr int CHUNKS = (BM + BN) * 8 / NT;  // 16-byte chunks per thread per stage
    static_assert(CHUNKS % 4 == 0, "stage copy must split into four parts");
    auto load_part = [&](int stage, int it, int part) {
        const int kt = use_list ? (int)sKt[it] : it;
        double *sA = smem + stage * STAGE_DOUBLES;
#pragma unroll
        for (int i = part * (CHUNKS / 4); i < (part + 1) * (CHUNKS / 4); i++) {
            const int c = tid + i * NT, row = c >> 3, ch = c & 7;  // rows 0..BM-1 -> A, BM.. -> B (contiguous in smem)
            const double *src = (row < BM) ? gA + (long long)row * G.ld : gB + (long long)(row - BM) * G.ld;
            cp_async16(sA + row * PITCH + ch * 2, src + kt * BK + ch * 2);
        }
    };
    auto load_stage = [&](int stage, int it) {
#pragma unroll
        for (int part = 0; part < 4; part++) load_part(stage, it, part);
    };

    double acc[MI][4][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nact) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nact; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nk = kt + STAGES - 1;
        const bool pf = nk < nact;
        const double *sA = smem + (kt % STAGES) * STAGE_DOUBLES;
        const double *sB = sA + BM * PITCH;
        if (MODE == MODE_DOT) {
            const double *pa = sA + (wm * WROWS + g) * PITCH + t4;
            const double *pb = sB + (wn * 32 + g) * PITCH + t4;
#pragma unroll
            for (int kk = 0; kk < BK / 4; kk++) {
                if (pf) load_part(nk % STAGES, nk, kk);
                double a[MI], b[4];
#pragma unroll
                for (int mi = 0; mi < MI; mi++) a[mi] = pa[mi * 8 * PITCH + kk * 4];
#pragma unroll
                for (int ni = 0; ni < 4; ni++) b[ni] = pb[ni * 8 * PITCH + kk * 4];
#pragma unroll
                for (int mi = 0; mi < MI; mi++)
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
        } else {
            // same accumulator ownership as the DMMA C fragment: rows 8*mi+g, cols 8*ni+2*t4+{0,1}
            const double *pa = sA + (wm * WROWS + g) * PITCH;
            const double *pb = sB + (wn * 32 + 2 * t4) * PITCH;
#pragma unroll
            for (int kk = 0; kk < BK; kk += 2) {
                if (pf && (kk & 3) == 0) load_part(nk % STAGES, nk, kk >> 2);
                double2 a[MI], b[4][2];
#pragma unroll
                for (int mi = 0; mi < MI; mi++) a[mi] = *reinterpret_cast<const double2 *>(pa + mi * 8 * PITCH + kk);
#pragma unroll
                for (int ni = 0; ni < 4; ni++) {
                    b[ni][0] = *reinterpret_cast<const double2 *>(pb + (ni * 8) * PITCH + kk);
                    b[ni][1] = *reinterpret_cast<const double2 *>(pb + (ni * 8 + 1) * PITCH + kk);
                }
#pragma unroll
                for (int mi = 0; mi < MI; mi++)
#pragma unroll
                    for (int ni = 0; ni < 4; ni++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            double v = acc[mi][ni][e];
                            v += fabs(a[mi].x - b[ni][e].x);  // k ascending, as the reference's sum(abs(..))
                            v += fabs(a[mi].y - b[ni][e].y);
                            acc[mi][ni][e] = v;
                        }
            }
        }
        cp_async_commit();  // the prefetch issued piecewise above (possibly empty: keeps the group count uniform)
    }
    cp_async_wait<0>();
    __syncthreads();  // pipeline buffers are free from here on

    // ---- v_ij -------------------------------------------------------------------------------
    if (MODE == MODE_DOT && EPI != EPI_DOT_STORE) {
#pragma unroll
        for (int mi = 0; mi < MI; mi++) {
            double na = sNA[wm * WROWS + mi * 8 + g];
#pragma unroll
            for (int ni = 0; ni < 4; ni++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    double nb = sNB[wn * 32 + ni * 8 + 2 * t4 + e];
                    const double d2 = na + nb - 2.0 * acc[mi][ni][e];
                    acc[mi][ni][e] = (d2 < 0.0) ? 0.0 : d2;  // clamp the cancellation noise, keep NaN (the reference propagates it)
                }
        }
    }

    if (EPI == EPI_LOCAL && G.sym && tn * BN < (tm + 1) * BM) {
        // diagonal-crossing tile of a symmetric kernel: drop pairs on or below the diagonal
        // (v = +inf -> exp(v * isig) = 0 for every sigma, also through the squaring chain)
#pragma unroll
        for (int mi = 0; mi < MI; mi++) {
            const int r = tm * BM + wm * WROWS + mi * 8 + g;
#pragma unroll
            for (int ni = 0; ni < 4; ni++)
#pragma unroll
                for (int e = 0; e < 2; e++)
                    if (tn * BN + wn * 32 + ni * 8 + 2 * t4 + e <= r) acc[mi][ni][e] = __longlong_as_double(0x7ff0000000000000LL);
        }
    }

    // ---- epilogues ----------------------------------------------------------------------------
    if (EPI == EPI_EXP_STORE) {
        for (int si = 0; si < P.nsig; si++) {
            const int s = P.chain ? P.order[si] : si;
            const bool do_exp = !P.chain || si == 0;  // warp-uniform
            if (!do_exp)
                for (int q = 0; q < P.nsq[si]; q++)
#pragma unroll
                    for (int mi = 0; mi < MI; mi++)
#pragma unroll
                        for (int ni = 0; ni < 4; ni++) {
                            acc[mi][ni][0] *= acc[mi][ni][0];
                            acc[mi][ni][1] *= acc[mi][ni][1];
                        }
            double *outS = P.out + (long long)s * P.slab;
#pragma unroll
            for (int mi = 0; mi < MI; mi++) {
                int row = wm * WROWS + mi * 8 + g;
                int sr = sSegA[row];
                double f = P.isig[s * P.zstride + sZA[row]];
#pragma unroll
                for (int ni = 0; ni < 4; ni++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        double v = acc[mi][ni][e];
                        if (do_exp) {
                            v = exp(v * f);
                            if (P.chain) acc[mi][ni][e] = v;
                        }
                        int sc = sSegB[wn * 32 + ni * 8 + 2 * t4 + e];
                        if (sr >= 0 && sc >= 0) outS[(long long)sr * P.ldo + sc] = v;
                    }
            }
        }
    } else if (EPI == EPI_DIST_STORE || EPI == EPI_DOT_STORE) {
#pragma unroll
        for (int mi = 0; mi < MI; mi++) {
            int sr = sSegA[wm * WROWS + mi * 8 + g];
            if (sr < 0) continue;
#pragma unroll
            for (int ni = 0; ni < 4; ni++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    int sc = sSegB[wn * 32 + ni * 8 + 2 * t4 + e];
                    if (sc >= 0) {
                        double v = acc[mi][ni][e];
                        if (MODE == MODE_DOT && EPI == EPI_DIST_STORE) {
                            // sqrt amplifies the expansion's cancellation error near d = 0 (duplicate / nearly equal
                            // rows): recompute those few elements by direct differences, as the reference does
                            const int row = wm * WROWS + mi * 8 + g, col = wn * 32 + ni * 8 + 2 * t4 + e;
                            if (v < 1e-6 * (sNA[row] + sNB[col])) {
                                const double *pa = gA + (long long)row * G.ld, *pb = gB + (long long)col * G.ld;
                                double s2 = 0.0;
                                for (int k = 0; k < (int)G.ld; k++) {
                                    const double t = pa[k] - pb[k];
                                    s2 += t * t;
                                }
                                v = s2;
                            }
                            v = sqrt(v);
                        }
                        P.out[(long long)sr * P.ldo + sc] = v;
                    }
                }
        }
    } else if (EPI == EPI_MAX) {
        double m = 0.0;
#pragma unroll
        for (int mi = 0; mi < MI; mi++) {
            if (sSegA[wm * WROWS + mi * 8 + g] < 0) continue;
#pragma unroll
            for (int ni = 0; ni < 4; ni++)
#pragma unroll
                for (int e = 0; e < 2; e++)
                    if (sSegB[wn * 32 + ni * 8 + 2 * t4 + e] >= 0) m = fmax(m, acc[mi][ni][e]);
        }
        m = warp_max(m);
        if (lane == 0) sRed[warp] = m;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < NT / 32; w++) m = fmax(m, sRed[w]);
            // v >= 0: IEEE order == unsigned integer order of the bit patterns
            atomicMax(reinterpret_cast<unsigned long long *>(P.out) + G.out_slot, (unsigned long long)__double_as_longlong(m));
        }
    } else {  // EPI_LOCAL
        // Each warp owns a 64x32 patch of atom pairs.  Per sigma: exp -> warp-private smem patch
        // (conflict-free 16-byte stores), then lane c walks column c down the rows, summing the rows
        // of one molecule (row segment); at each row-segment end a segmented shuffle reduction sums
        // the lanes of one molecule (column segment) and the segment head issues one RED.F64.
        double *patch = smem + warp * (WROWS * EPITCH);
        const int colseg = sSegB[wn * 32 + lane];
        const int prevseg = __shfl_up_sync(0xffffffffu, colseg, 1);
        const bool head = (lane == 0) || (prevseg != colseg);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        // last lane of my run
        const unsigned above = (lane == 31) ? 0u : (heads >> (lane + 1));
        const int run_end = above ? (lane + __ffs(above) - 1) : 31;
        const int *rowseg = sSegA + wm * WROWS;
        // row-segment structure of this warp's 64 rows as bit masks (identical for all lanes).  The rows
        // are walked as two independent 32-row chains (ILP), so row 31 always closes a segment.
        constexpr int HALF = WROWS / 2;
        unsigned last_lo, last_hi, valid_lo, valid_hi;
        {
            const bool in = lane < HALF;
            int r0 = in ? lane : 0, r1 = r0 + HALF;
            int s0 = rowseg[r0], s1 = rowseg[r1];
            int n0 = (r0 == HALF - 1) ? -2 : rowseg[r0 + 1], n1 = (r1 == WROWS - 1) ? -2 : rowseg[r1 + 1];
            last_lo = __ballot_sync(0xffffffffu, in && n0 != s0);
            last_hi = __ballot_sync(0xffffffffu, in && n1 != s1);
            valid_lo = __ballot_sync(0xffffffffu, in && s0 >= 0);
            valid_hi = __ballot_sync(0xffffffffu, in && s1 >= 0);
        }
        for (int si = 0; si < P.nsig; si++) {
            const int s = P.chain ? P.order[si] : si;
            const bool do_exp = !P.chain || si == 0;  // warp-uniform
            if (!do_exp)
                for (int q = 0; q < P.nsq[si]; q++)
#pragma unroll
                    for (int mi = 0; mi < MI; mi++)
#pragma unroll
                        for (int ni = 0; ni < 4; ni++) {
                            acc[mi][ni][0] *= acc[mi][ni][0];
                            acc[mi][ni][1] *= acc[mi][ni][1];
                        }
#pragma unroll
            for (int mi = 0; mi < MI; mi++) {
                int row = mi * 8 + g;
                double f = P.isig[s * P.zstride + sZA[wm * WROWS + row]];
#pragma unroll
                for (int ni = 0; ni < 4; ni++) {
                    double2 v;
                    v.x = acc[mi][ni][0];
                    v.y = acc[mi][ni][1];
                    if (do_exp) {
                        v.x = exp(v.x * f);
                        v.y = exp(v.y * f);
                        if (P.chain) { acc[mi][ni][0] = v.x; acc[mi][ni][1] = v.y; }
                    }
                    *reinterpret_cast<double2 *>(patch + row * EPITCH + ni * 8 + 2 * t4) = v;
                }
            }
            __syncwarp();
            double *outS = P.out + (long long)s * P.slab;
            auto flush = [&](double v, int sr) {  // called warp-uniformly
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    double u = __shfl_down_sync(0xffffffffu, v, o);
                    if (lane + o <= run_end) v += u;
                }
                if (head && colseg >= 0) atomicAdd(outS + (long long)sr * P.ldo + colseg, v);
            };
            double run0 = 0.0, run1 = 0.0;
#pragma unroll 1
            for (int rb = 0; rb < HALF; rb += 8) {
                double u8[8], w8[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    u8[j] = patch[(rb + j) * EPITCH + lane];
                    w8[j] = patch[(HALF + rb + j) * EPITCH + lane];
                }
                const unsigned lb0 = (last_lo >> rb) & 0xffu, vb0 = (valid_lo >> rb) & 0xffu;
                const unsigned lb1 = (last_hi >> rb) & 0xffu, vb1 = (valid_hi >> rb) & 0xffu;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if ((vb0 >> j) & 1u) run0 += u8[j];
                    if ((vb1 >> j) & 1u) run1 += w8[j];
                    if ((lb0 >> j) & 1u) {
                        if ((vb0 >> j) & 1u) flush(run0, rowseg[rb + j]);
                        run0 = 0.0;
                    }
                    if ((lb1 >> j) & 1u) {
                        if ((vb1 >> j) & 1u) flush(run1, rowseg[HALF + rb + j]);
                        run1 = 0.0;
                    }
                }
            }
            __syncwarp();
        }
    }
}

template <int MODE, int EPI>
void launch_pair_tiles(const PairParams &P, int total_tiles, cudaStream_t stream) {
    if (total_tiles <= 0) return;
    AQML_CUDA(cudaFuncSetAttribute(pair_tile_kernel<MODE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_BYTES));
    pair_tile_kernel<MODE, EPI><<<total_tiles, NT, PIPE_BYTES, stream>>>(P);
    AQML_LAUNCH_CHECK();
}

}  // namespace aqml
