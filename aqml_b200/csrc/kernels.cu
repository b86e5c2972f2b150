// Kernel-matrix assembly: host orchestration + C ABI (see include/aqml_b200.h for the reference
// interfaces each entry point replaces: fkernels.f90, fdistance.f90, algo/aqml.py:204-255).
#include "pairtile.cuh"
#include "packing.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

namespace aqml {

std::string &last_error_ref() {
    static thread_local std::string s;
    return s;
}
std::atomic<long long> g_launches{0};

void require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        fail(AQML_ECUDA, "no CUDA device available (aqml_b200 has no CPU fallback): %s",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    // Keep freed workspace blocks in the stream-ordered pool instead of returning them to the driver at
    // every synchronisation (default release threshold 0): a step re-uses the same ~9 GB of operands.
    static std::atomic<unsigned> configured{0};
    int dev = 0;
    AQML_CUDA(cudaGetDevice(&dev));
    if (dev < 32 && !(configured.load() & (1u << dev))) {
        cudaMemPool_t pool;
        AQML_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long keep = ~0ull;
        AQML_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        configured.fetch_or(1u << dev);
    }
}

// =============================================================================================
// device kernels: packing
// =============================================================================================
// One block per packed row: gather (optional row map / column map), subtract an optional mean,
// zero-pad to the 16-column pitch, and produce |row|^2.
__global__ void __launch_bounds__(256) pack_rows_kernel(PackJob J) {
    const int r = blockIdx.x;
    const double *sp;
    if (J.srcrows) {
        sp = J.srcrows[r];
    } else {
        int src = J.rowmap ? J.rowmap[r] : (r < J.nvalid ? r : -1);
        sp = (src >= 0) ? J.src + (long long)src * J.lds : nullptr;
    }
    double *dp = J.dst + (long long)r * J.ldd;
    double ss = 0.0;
    __shared__ unsigned smask[32];
    const bool want_mask = J.mask && J.maskw <= 32;
    if (want_mask && threadIdx.x < 32) smask[threadIdx.x] = 0u;
    if (want_mask) __syncthreads();
    for (int c = threadIdx.x; c < J.ldd; c += blockDim.x) {
        double v = 0.0;
        if (sp && c < J.ncols) {
            int sc = J.colmap ? J.colmap[c] : c;
            v = sp[sc];
            if (J.mean) v -= J.mean[c];
        }
        dp[c] = v;
        ss += v * v;
        if (want_mask && v != 0.0) atomicOr(&smask[c >> 9], 1u << ((c >> 4) & 31));
    }
    if (want_mask) {
        __syncthreads();
        if ((int)threadIdx.x < J.maskw && smask[threadIdx.x])
            atomicOr(J.mask + (long long)(r >> 6) * J.maskw + threadIdx.x, smask[threadIdx.x]);
    }
    __shared__ double red[8];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0 && J.norm) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
        J.norm[r] = t;
    }
}

// Column sums and non-zero flags over a list of rows (support detection + centring).
// grid (row chunks, ceil(D/256)); thread = one column over CS_ROWS rows.
constexpr int CS_ROWS = 64;
__global__ void __launch_bounds__(256) colstats_kernel(const double *x, long long ldx, const int *rows, int nrows, int D,
                                                       double *colsum, unsigned *colnz) {
    int c = blockIdx.y * 256 + threadIdx.x;
    if (c >= D) return;
    int r0 = blockIdx.x * CS_ROWS, r1 = min(nrows, r0 + CS_ROWS);
    double s = 0.0;
    unsigned nz = 0;
    for (int i = r0; i < r1; i++) {
        int r = rows ? rows[i] : i;
        double v = x[(long long)r * ldx + c];
        if (fabs(v) <= 1.7976931348623157e308) s += v;  // a NaN / Inf must stay in its own row (centring is only a shift)
        nz |= (v != 0.0);
    }
    if (colsum) atomicAdd(colsum + c, s);
    if (nz) atomicOr(colnz + c, 1u);
}

// r[i] = distance of row i to the point c: Euclidean (p = 2) or sum of absolute differences (p = 1), by direct
// differences (no cancellation).  One warp per row.
__global__ void __launch_bounds__(256) row_dist_kernel(int p, const double *x, long long ld, int n, const double *c, int ncols,
                                                       double *r) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const double *xr = x + (long long)row * ld;
    double s = 0.0;
    for (int k = lane; k < ncols; k += 32) {
        const double d = xr[k] - c[k];
        s += (p == 2) ? d * d : fabs(d);
    }
    s = warp_sum(s);
    if (lane == 0) r[row] = (p == 2) ? sqrt(s) : s;
}

__global__ void scale_gather_kernel(const double *colsum, const int *colmap, int n, double scale, double *mean) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mean[i] = colsum[colmap ? colmap[i] : i] * scale;
}

// K[k][a][b] += cnt1[a][e] * (nas2[b] - cnt2[b][e]) * exp(1e9 * isig[k][z_e])  -- the reference gives
// element-mismatched pairs d2 = 1e9 (fkernels.f90:73), which is not exactly 0 for very wide sigmas.
__global__ void mismatch_kernel(const int *cnt1, const int *cnt2, const int *nas2, const double *term, int nel, int nm1,
                                int nm2, int nsig, double *K) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nm1 * nm2) return;
    int a = (int)(i / nm2), b = (int)(i % nm2);
    for (int k = 0; k < nsig; k++) {
        double s = 0.0;
        for (int e = 0; e < nel; e++)
            s += (double)cnt1[a * nel + e] * (double)(nas2[b] - cnt2[b * nel + e]) * term[k * nel + e];
        K[((long long)k * nm1 + a) * nm2 + b] += s;
    }
}

void launch_pack(const PackJob &J, int nrows_packed, cudaStream_t st) {
    if (nrows_packed <= 0) return;
    pack_rows_kernel<<<nrows_packed, 256, 0, st>>>(J);
    AQML_LAUNCH_CHECK();
}

void launch_colstats(const double *x, long long ldx, const int *rows, int nrows, int D, double *colsum, unsigned *colnz,
                     cudaStream_t st) {
    if (nrows <= 0 || D <= 0) return;
    dim3 grid((nrows + CS_ROWS - 1) / CS_ROWS, (D + 255) / 256);  // row chunks on x (no 65535 limit)
    colstats_kernel<<<grid, 256, 0, st>>>(x, ldx, rows, nrows, D, colsum, colnz);
    AQML_LAUNCH_CHECK();
}

// =============================================================================================
// host helpers
// =============================================================================================
// Pack a molecule-sorted row list into 128-row tiles (-1 padding only at the very end).  Molecules may
// straddle tile / warp-patch boundaries: their partial sums meet in K through the FP64 REDs.  (Aligning
// tiles to molecule boundaries wasted ~5% of the rows per side, ~9% of the DMMA work, for hydrogen.)
void build_tiles(const std::vector<int> &rows, const std::vector<int> &mol_of_row, std::vector<int> &rowmap,
                 std::vector<int> &seg) {
    rowmap = rows;
    seg = mol_of_row;
    size_t padded = (size_t)round_up((int64_t)rows.size(), BM);
    rowmap.resize(padded, -1);
    seg.resize(padded, -1);
}

// Order molecules by composition class (which elements they contain, once / more than once) so that a
// 64/128-row block of packed rows holds molecules with the same set of non-zero many-body blocks: whole
// k-tiles of such a block are exactly zero and the pair kernel skips them.  Heavier elements are the
// more significant digits, so the rare ones (F, then O, N ...) separate first.
std::vector<int> class_sorted_molecules(const int *zs, const int *nas, int nm) {
    std::vector<unsigned long long> key(nm, 0ull);
    std::vector<int> zlist;
    size_t a0 = 0;
    {
        std::vector<char> seen(4096, 0);
        size_t A = 0;
        for (int m = 0; m < nm; m++) A += nas[m];
        for (size_t a = 0; a < A; a++) {
            int z = zs[a] % 1000;
            if (z >= 0 && z < 4096) seen[z] = 1;
        }
        for (int z = 0; z < 4096; z++) if (seen[z]) zlist.push_back(z);
    }
    std::vector<int> rank(4096, -1);
    for (size_t i = 0; i < zlist.size(); i++) rank[zlist[i]] = (int)i;
    std::vector<int> cnt(zlist.size());
    for (int m = 0; m < nm; m++) {
        std::fill(cnt.begin(), cnt.end(), 0);
        for (int i = 0; i < nas[m]; i++) {
            int z = zs[a0 + i] % 1000;
            if (z >= 0 && z < 4096) cnt[rank[z]]++;
        }
        a0 += nas[m];
        unsigned long long k = 0;
        for (size_t e = zlist.size(); e-- > 0;) k = k * 3ull + (unsigned long long)std::min(cnt[e], 2);
        key[m] = k;
    }
    std::vector<int> order(nm);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return key[x] < key[y]; });
    return order;
}

static std::vector<int> mol_index(const int *nas, int nm, int &A) {
    std::vector<int> m;
    A = 0;
    for (int i = 0; i < nm; i++) {
        AQML_REQUIRE(nas[i] >= 0, "negative atom count for molecule %d", i);
        A += nas[i];
    }
    m.reserve(A);
    for (int i = 0; i < nm; i++) m.insert(m.end(), nas[i], i);
    return m;
}

unsigned long long *pair_stats_buffer();

// Evaluation plan for the sigma ladder: all rows of the host table (nsig x zstride) must be exact
// power-of-two multiples of the widest sigma's row; otherwise every sigma gets its own exp.
void plan_sigma_chain(PairParams &P, const std::vector<double> &isig) {
    P.chain = 0;
    P.isig_host = isig.data();
    const int S = P.nsig, Z = P.zstride;
    if (S < 2 || S > 16 || Z < 1 || (int)isig.size() != S * Z) return;
    std::vector<int> ord(S);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return std::fabs(isig[(size_t)a * Z]) < std::fabs(isig[(size_t)b * Z]); });
    int nsq[16] = {0};
    for (int i = 1; i < S; i++) {
        int k = -1;
        for (int z = 0; z < Z; z++) {
            double a = isig[(size_t)ord[i] * Z + z], b = isig[(size_t)ord[i - 1] * Z + z];
            if (!(std::isfinite(a) && std::isfinite(b)) || b == 0.0) return;
            int e = 0;
            double m = std::frexp(a / b, &e);           // a/b = m * 2^e, m in [0.5,1)
            if (m != 0.5 || e - 1 < 0 || e - 1 > 4) return;   // ratio must be 2^k, 0 <= k <= 4 (the kernels unroll <= 4 squarings per step)
            if (a != std::ldexp(b, e - 1)) return;       // exact
            if (k < 0) k = e - 1; else if (k != e - 1) return;
        }
        nsq[i] = k;  // ratio 2^k: k squarings ( x -> x^(2^k) )
    }
    // the narrowest sigma's value is the widest one's after sum(k) squarings: its error is 2^sum(k) times that of the exp
    int total = 0;
    for (int i = 1; i < S; i++) total += nsq[i];
    if (total > 8) return;
    P.chain = 1;
    for (int i = 0; i < S; i++) { P.order[i] = ord[i]; P.nsq[i] = nsq[i]; }
}

static void run_pairs(int mode, int epi, const std::vector<PairGroup> &groups, PairParams P, Workspace &ws) {
    int total = 0;
    std::vector<PairGroup> gs = groups;
    for (auto &g : gs) { g.tile_begin = total; total += g.tilesM * g.tilesN; }
    if (total == 0) return;
    P.groups = ws.upload(gs);
    P.ngroups = (int)gs.size();
    cudaStream_t st = ws.stream();
    P.sm_arrive = (total >= 1024) ? ws.zeros<int>(1024) : nullptr;
    P.stats = pair_stats_buffer();
#define AQML_DISPATCH(M, E) if (mode == M && epi == E) { launch_pair_tiles<M, E>(P, total, st); return; }
    AQML_DISPATCH(MODE_DOT, EPI_EXP_STORE)
    AQML_DISPATCH(MODE_DOT, EPI_DIST_STORE)
    AQML_DISPATCH(MODE_DOT, EPI_MAX)
    AQML_DISPATCH(MODE_DOT, EPI_LOCAL)
    AQML_DISPATCH(MODE_DOT, EPI_DOT_STORE)
    AQML_DISPATCH(MODE_L1, EPI_EXP_STORE)
    AQML_DISPATCH(MODE_L1, EPI_DIST_STORE)
    AQML_DISPATCH(MODE_L1, EPI_MAX)
    AQML_DISPATCH(MODE_L1, EPI_LOCAL)
#undef AQML_DISPATCH
    fail(AQML_EINVAL, "bad pair-tile mode");
}

void run_pairs_public(int mode, int epi, const std::vector<PairGroup> &groups, PairParams P, Workspace &ws) {
    run_pairs(mode, epi, groups, P, ws);
}

// device-side counters of executed vs dense k-tiles (bench.py reports the executed DMMA work)
unsigned long long *pair_stats_buffer() {
    static unsigned long long *buf[32] = {nullptr};
    int dev = 0;
    AQML_CUDA(cudaGetDevice(&dev));
    if (dev >= 32) return nullptr;
    if (!buf[dev]) {
        AQML_CUDA(cudaMalloc((void **)&buf[dev], 512));  // [0..1] k-tile counts, [2..] engine debug timers
        AQML_CUDA(cudaMemset(buf[dev], 0, 512));
    }
    return buf[dev];
}

static PairGroup blank_group() {
    PairGroup g;
    memset(&g, 0, sizeof(g));
    g.segA_n = 0x7fffffff;
    return g;
}

// dense (n, D) rows -> packed [round_up(n,128)][round_up(D,16)] (+ norms), optional mean
struct Packed {
    double *x = nullptr, *norm = nullptr;
    long long ld = 0;
    int rows = 0, tiles = 0;
};
static Packed pack_dense(Workspace &ws, const double *src, long long lds, int n, int D, const double *mean, bool want_norm) {
    Packed p;
    p.ld = round_up(std::max(D, 1), BK);
    p.tiles = (n + BM - 1) / BM;
    p.rows = p.tiles * BM;
    p.x = ws.alloc<double>((size_t)p.rows * p.ld);
    p.norm = want_norm ? ws.alloc<double>(p.rows) : nullptr;
    PackJob J{};
    J.src = src; J.lds = lds; J.rowmap = nullptr; J.colmap = nullptr; J.mean = mean;
    J.dst = p.x; J.ldd = p.ld; J.norm = p.norm; J.ncols = D; J.nvalid = n;
    launch_pack(J, p.rows, ws.stream());
    return p;
}

static const double *column_mean(Workspace &ws, const double *x, long long ldx, int n, int D) {
    double *sum = ws.zeros<double>(D);
    unsigned *nz = ws.zeros<unsigned>(D);
    launch_colstats(x, ldx, nullptr, n, D, sum, nz, ws.stream());
    double *mean = ws.alloc<double>(D);
    scale_gather_kernel<<<(D + 255) / 256, 256, 0, ws.stream()>>>(sum, nullptr, D, 1.0 / std::max(n, 1), mean);
    AQML_LAUNCH_CHECK();
    return mean;
}

static double pruned_max(int p, const double *A, int nA, const double *B, int nB, long long ld, bool same, Workspace &ws);

static std::vector<double> inv_sigma_table(int kind, const double *sigmas, size_t n) {
    std::vector<double> t(n);
    for (size_t i = 0; i < n; i++)
        t[i] = (kind == AQML_KERNEL_GAUSSIAN) ? -0.5 / (sigmas[i] * sigmas[i]) : -1.0 / sigmas[i];
    return t;
}

// =============================================================================================
// global kernels / distances
// =============================================================================================
// epi: EPI_EXP_STORE (K (nsig,na,nb)), EPI_DIST_STORE (D (na,nb)), EPI_MAX (out[0] bits)
static void global_pairs(int kind, int epi, const double *a, int na, long long lda, const double *b, int nb, long long ldb,
                         int D, const double *sigmas, int nsig, int center, double *out, cudaStream_t st) {
    AQML_REQUIRE(na >= 0 && nb >= 0 && D >= 1, "bad shape na=%d nb=%d D=%d", na, nb, D);
    AQML_REQUIRE(lda >= D && ldb >= D, "row pitch smaller than D");
    if (na == 0 || nb == 0) return;
    require_device();
    Workspace ws(st);
    const int mode = (kind == AQML_KERNEL_GAUSSIAN) ? MODE_DOT : MODE_L1;
    const double *mean = (mode == MODE_DOT && center && epi != EPI_DOT_STORE) ? column_mean(ws, b, ldb, nb, D) : nullptr;
    const bool same = (a == b && na == nb && lda == ldb);
    Packed pb = pack_dense(ws, b, ldb, nb, D, mean, mode == MODE_DOT);
    Packed pa = same ? pb : pack_dense(ws, a, lda, na, D, mean, mode == MODE_DOT);
    if (epi == EPI_MAX) {
        const double v = pruned_max(mode == MODE_DOT ? 2 : 1, pa.x, na, pb.x, nb, pa.ld, same, ws);
        AQML_CUDA(cudaMemcpyAsync(out, &v, 8, cudaMemcpyHostToDevice, st));
        AQML_CUDA(cudaStreamSynchronize(st));
        return;
    }
    PairGroup g = blank_group();
    g.A = pa.x; g.B = pb.x; g.normA = pa.norm; g.normB = pb.norm; g.ld = pa.ld;
    g.nA = na; g.nB = nb; g.tilesM = pa.rows / BM; g.tilesN = pb.rows / BN;
    g.sym = 0;
    PairParams P{};
    P.out = out; P.ldo = nb; P.slab = (long long)na * nb; P.nsig = nsig; P.zstride = 1;
    if (epi == EPI_EXP_STORE) {
        AQML_REQUIRE(nsig >= 1 && sigmas, "need at least one sigma");
        for (int s = 0; s < nsig; s++) AQML_REQUIRE(sigmas[s] != 0.0, "sigma[%d] is zero", s);
        std::vector<double> tab = inv_sigma_table(kind, sigmas, nsig);
        P.isig = ws.upload(tab);
        plan_sigma_chain(P, tab);
    }
    if (epi == EPI_EXP_STORE && mode == MODE_DOT) {
        // default engine for the Gaussian kernel: int8 digit planes on the tcgen05 tensor cores (i8pair.cuh)
        I8Side sa{pa.x, pa.norm, nullptr, nullptr, 0, pa.rows, pa.ld}, sb{pb.x, pb.norm, nullptr, nullptr, 0, pb.rows, pb.ld};
        if (i8_global_pair(sa, na, sb, nb, P, ws)) return;
    }
    run_pairs(mode, epi, {g}, P, ws);
}

// host-buffer wrapper: K column-major (na, nb) == row-major (nb, na) of the swapped problem
static void global_host(int kind, int epi, const double *a, int na, const double *b, int nb, int D, double *k, double sigma) {
    AQML_REQUIRE(a && b && k, "null pointer");
    AQML_REQUIRE(na >= 0 && nb >= 0 && D >= 1, "bad shape");
    if (na == 0 || nb == 0) return;
    require_device();
    cudaStream_t st = 0;
    Workspace ws(st);
    double *da = ws.upload(a, (size_t)na * D), *db = ws.upload(b, (size_t)nb * D);
    double *dk = ws.alloc<double>((size_t)na * nb);
    global_pairs(kind, epi, db, nb, D, da, na, D, D, &sigma, 1, 1, dk, st);
    AQML_CUDA(cudaMemcpyAsync(k, dk, sizeof(double) * (size_t)na * nb, cudaMemcpyDeviceToHost, st));
    AQML_CUDA(cudaStreamSynchronize(st));
}

// =============================================================================================
// local kernels
// =============================================================================================
struct ElementRows {
    std::vector<int> rowmap, seg;  // tile-padded
    int n = 0;                      // valid rows
};

static void local_gaussian(const double *x1, long long ldx1, const double *x2, long long ldx2, const int *zs1,
                           const int *zs2, const int *nas1, const int *nas2, int zmax, int iab, const double *sigmas,
                           int nm1, int nm2, int S, int D, int center, double *K, cudaStream_t st) {
    AQML_REQUIRE(x1 && x2 && zs1 && zs2 && nas1 && nas2 && sigmas && K, "null pointer");
    AQML_REQUIRE(nm1 >= 0 && nm2 >= 0 && S >= 1 && D >= 1 && zmax >= 1, "bad shape");
    AQML_REQUIRE(ldx1 >= D && ldx2 >= D, "row pitch smaller than D");
    require_device();
    int A1, A2;
    std::vector<int> mol1 = mol_index(nas1, nm1, A1), mol2 = mol_index(nas2, nm2, A2);
    for (int i = 0; i < A1; i++) AQML_REQUIRE(zs1[i] >= 1 && zs1[i] <= zmax, "zs1[%d]=%d outside 1..zmax", i, zs1[i]);
    for (int i = 0; i < A2; i++) AQML_REQUIRE(zs2[i] >= 1 && zs2[i] <= zmax, "zs2[%d]=%d outside 1..zmax", i, zs2[i]);
    for (int i = 0; i < S * zmax; i++) AQML_REQUIRE(sigmas[i] != 0.0, "sigma is zero");
    AQML_CUDA(cudaMemsetAsync(K, 0, sizeof(double) * (size_t)S * nm1 * nm2, st));
    if (nm1 == 0 || nm2 == 0 || A1 == 0 || A2 == 0) return;
    Workspace ws(st);
    std::vector<double> isig = inv_sigma_table(AQML_KERNEL_GAUSSIAN, sigmas, (size_t)S * zmax);
    PairParams P{};
    P.isig = ws.upload(isig); P.nsig = S; P.zstride = zmax;
    plan_sigma_chain(P, isig);
    P.out = K; P.ldo = nm2; P.slab = (long long)nm1 * nm2;
    std::vector<PairGroup> groups;

    if (iab) {
        // every atom pair, width of the first argument's atom (fkernels.f90:76): one group, full D
        std::vector<int> rows1(A1), rows2(A2), rm1, sg1, rm2, sg2;
        std::iota(rows1.begin(), rows1.end(), 0);
        std::iota(rows2.begin(), rows2.end(), 0);
        build_tiles(rows1, mol1, rm1, sg1);
        build_tiles(rows2, mol2, rm2, sg2);
        std::vector<int> z1p(rm1.size());
        for (size_t i = 0; i < rm1.size(); i++) z1p[i] = rm1[i] >= 0 ? zs1[rm1[i]] - 1 : 0;
        const double *mean = center ? column_mean(ws, x2, ldx2, A2, D) : nullptr;
        long long ld = round_up(D, BK);
        auto pack = [&](const double *x, long long ldx, const std::vector<int> &rm, double *&px, double *&pn) {
            px = ws.alloc<double>(rm.size() * (size_t)ld);
            pn = ws.alloc<double>(rm.size());
            PackJob J{};
            J.src = x; J.lds = ldx; J.rowmap = ws.upload(rm); J.colmap = nullptr; J.mean = mean;
            J.dst = px; J.ldd = ld; J.norm = pn; J.ncols = D; J.nvalid = 0;
            launch_pack(J, (int)rm.size(), st);
        };
        double *p1, *n1, *p2, *n2;
        pack(x1, ldx1, rm1, p1, n1);
        pack(x2, ldx2, rm2, p2, n2);
        PairGroup g = blank_group();
        g.A = p1; g.B = p2; g.normA = n1; g.normB = n2; g.ld = ld;
        g.segA = ws.upload(sg1); g.segB = ws.upload(sg2); g.zA = ws.upload(z1p);
        g.tilesM = (int)rm1.size() / BM; g.tilesN = (int)rm2.size() / BN;
        groups.push_back(g);
        run_pairs(MODE_DOT, EPI_LOCAL, groups, P, ws);
        return;
    }

    // ---- element-matched: one group per element present on both sides ---------------------------
    // rows of an element: molecules in composition-class order (block sparsity), atoms in list order
    std::vector<std::vector<int>> r1(zmax + 1), r2(zmax + 1);
    {
        std::vector<int> off1(nm1 + 1, 0), off2(nm2 + 1, 0);
        for (int m = 0; m < nm1; m++) off1[m + 1] = off1[m] + nas1[m];
        for (int m = 0; m < nm2; m++) off2[m + 1] = off2[m] + nas2[m];
        for (int m : class_sorted_molecules(zs1, nas1, nm1))
            for (int i = off1[m]; i < off1[m + 1]; i++) r1[zs1[i]].push_back(i);
        for (int m : class_sorted_molecules(zs2, nas2, nm2))
            for (int i = off2[m]; i < off2[m + 1]; i++) r2[zs2[i]].push_back(i);
    }
    std::vector<int> els;
    for (int z = 1; z <= zmax; z++) if (!r1[z].empty() && !r2[z].empty()) els.push_back(z);
    const int ne = (int)els.size();

    // element-mismatched pairs: exp(1e9 * isig) (exactly 0 unless sigma is huge)
    {
        std::vector<int> allz;
        for (int z = 1; z <= zmax; z++) if (!r1[z].empty()) allz.push_back(z);
        std::vector<double> term((size_t)S * allz.size());
        bool any = false;
        for (int k = 0; k < S; k++)
            for (size_t e = 0; e < allz.size(); e++) {
                term[k * allz.size() + e] = std::exp((double)1.e9f * isig[(size_t)k * zmax + allz[e] - 1]);
                any |= term[k * allz.size() + e] != 0.0;
            }
        if (any) {
            int nel = (int)allz.size();
            std::vector<int> c1((size_t)nm1 * nel, 0), c2((size_t)nm2 * nel, 0), zi(zmax + 1, -1);
            for (int e = 0; e < nel; e++) zi[allz[e]] = e;
            for (int i = 0; i < A1; i++) c1[(size_t)mol1[i] * nel + zi[zs1[i]]]++;
            for (int i = 0; i < A2; i++) if (zi[zs2[i]] >= 0) c2[(size_t)mol2[i] * nel + zi[zs2[i]]]++;
            long long n = (long long)nm1 * nm2;
            mismatch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws.upload(c1), ws.upload(c2), ws.upload(nas2, nm2),
                                                                        ws.upload(term), nel, nm1, nm2, S, K);
            AQML_LAUNCH_CHECK();
        }
    }
    if (ne == 0) return;

    // column support (and sums for centring) per element over both sets
    double *colsum = ws.zeros<double>((size_t)ne * D);
    unsigned *colnz = ws.zeros<unsigned>((size_t)ne * D);
    std::vector<int *> d_rows1(ne), d_rows2(ne);
    for (int e = 0; e < ne; e++) {
        int z = els[e];
        d_rows1[e] = ws.upload(r1[z]);
        d_rows2[e] = ws.upload(r2[z]);
        launch_colstats(x1, ldx1, d_rows1[e], (int)r1[z].size(), D, colsum + (size_t)e * D, colnz + (size_t)e * D, st);
        launch_colstats(x2, ldx2, d_rows2[e], (int)r2[z].size(), D, colsum + (size_t)e * D, colnz + (size_t)e * D, st);
    }
    std::vector<unsigned> nz((size_t)ne * D);
    AQML_CUDA(cudaMemcpyAsync(nz.data(), colnz, sizeof(unsigned) * nz.size(), cudaMemcpyDeviceToHost, st));
    AQML_CUDA(cudaStreamSynchronize(st));

    std::vector<I8Pair> i8pairs;
    for (int e = 0; e < ne; e++) {
        int z = els[e];
        std::vector<int> colmap;
        for (int c = 0; c < D; c++) if (nz[(size_t)e * D + c]) colmap.push_back(c);
        if (colmap.empty()) colmap.push_back(0);  // all-zero rows: d2 = 0 for every pair
        int Dz = (int)colmap.size();
        long long ld = round_up(Dz, BK);
        int *d_colmap = ws.upload(colmap);
        double *mean = nullptr;
        if (center) {
            mean = ws.alloc<double>(Dz);
            double scale = 1.0 / (double)(r1[z].size() + r2[z].size());
            scale_gather_kernel<<<(Dz + 255) / 256, 256, 0, st>>>(colsum + (size_t)e * D, d_colmap, Dz, scale, mean);
            AQML_LAUNCH_CHECK();
        }
        std::vector<int> m1(r1[z].size()), m2(r2[z].size()), rm1, sg1, rm2, sg2;
        for (size_t i = 0; i < m1.size(); i++) m1[i] = mol1[r1[z][i]];
        for (size_t i = 0; i < m2.size(); i++) m2[i] = mol2[r2[z][i]];
        build_tiles(r1[z], m1, rm1, sg1);
        build_tiles(r2[z], m2, rm2, sg2);
        const int mw = mask_words(ld);
        const bool sparse = !center && mw <= 32;  // centring turns structural zeros into -mean
        auto pack = [&](const double *x, long long ldx, const std::vector<int> &rm, double *&px, double *&pn,
                        unsigned *&pm) {
            px = ws.alloc<double>(rm.size() * (size_t)ld);
            pn = ws.alloc<double>(rm.size());
            pm = sparse ? ws.zeros<unsigned>(rm.size() / 64 * (size_t)mw) : nullptr;
            PackJob J{};
            J.src = x; J.lds = ldx; J.rowmap = ws.upload(rm); J.colmap = d_colmap; J.mean = mean;
            J.dst = px; J.ldd = ld; J.norm = pn; J.ncols = Dz; J.nvalid = 0; J.mask = pm; J.maskw = mw;
            launch_pack(J, (int)rm.size(), st);
        };
        double *p1, *n1, *p2, *n2;
        unsigned *k1, *k2;
        pack(x1, ldx1, rm1, p1, n1, k1);
        pack(x2, ldx2, rm2, p2, n2, k2);
        PairGroup g = blank_group();
        g.A = p1; g.B = p2; g.normA = n1; g.normB = n2; g.ld = ld;
        g.maskA = k1; g.maskB = k2; g.maskw = mw;
        g.segA = ws.upload(sg1); g.segB = ws.upload(sg2); g.zA = nullptr; g.zcol = z - 1;
        g.tilesM = (int)rm1.size() / BM; g.tilesN = (int)rm2.size() / BN;
        groups.push_back(g);
        I8Pair pr;
        pr.a = I8Side{p1, n1, g.segA, k1, mw, (int)rm1.size(), ld};
        pr.b = I8Side{p2, n2, g.segB, k2, mw, (int)rm2.size(), ld};
        pr.zcol = z - 1;
        i8pairs.push_back(pr);
    }
    // default engine: int8 digit planes on the tcgen05 tensor cores; FP64 DMMA when disabled / not applicable
    if (!i8_local_pairs(i8pairs, P, ws)) run_pairs(MODE_DOT, EPI_LOCAL, groups, P, ws);
}

static void local_laplacian(const double *x1, long long ldx1, const double *x2, long long ldx2, const int *nas1,
                            const int *nas2, const double *sigmas, int nm1, int nm2, int S, int D, double *K,
                            cudaStream_t st) {
    AQML_REQUIRE(x1 && x2 && nas1 && nas2 && sigmas && K, "null pointer");
    AQML_REQUIRE(nm1 >= 0 && nm2 >= 0 && S >= 1 && D >= 1, "bad shape");
    require_device();
    int A1, A2;
    std::vector<int> mol1 = mol_index(nas1, nm1, A1), mol2 = mol_index(nas2, nm2, A2);
    for (int i = 0; i < S; i++) AQML_REQUIRE(sigmas[i] != 0.0, "sigma is zero");
    AQML_CUDA(cudaMemsetAsync(K, 0, sizeof(double) * (size_t)S * nm1 * nm2, st));
    if (A1 == 0 || A2 == 0) return;
    Workspace ws(st);
    std::vector<int> rows1(A1), rows2(A2), rm1, sg1, rm2, sg2;
    std::iota(rows1.begin(), rows1.end(), 0);
    std::iota(rows2.begin(), rows2.end(), 0);
    build_tiles(rows1, mol1, rm1, sg1);
    build_tiles(rows2, mol2, rm2, sg2);
    long long ld = round_up(D, BK);
    auto pack = [&](const double *x, long long ldx, const std::vector<int> &rm) {
        double *px = ws.alloc<double>(rm.size() * (size_t)ld);
        PackJob J{};
        J.src = x; J.lds = ldx; J.rowmap = ws.upload(rm); J.dst = px; J.ldd = ld; J.ncols = D;
        launch_pack(J, (int)rm.size(), st);
        return px;
    };
    PairGroup g = blank_group();
    g.A = pack(x1, ldx1, rm1); g.B = pack(x2, ldx2, rm2); g.ld = ld;
    g.segA = ws.upload(sg1); g.segB = ws.upload(sg2);
    g.tilesM = (int)rm1.size() / BM; g.tilesN = (int)rm2.size() / BN;
    PairParams P{};
    std::vector<double> tab = inv_sigma_table(AQML_KERNEL_LAPLACIAN, sigmas, S);
    P.isig = ws.upload(tab); P.nsig = S; P.zstride = 1;
    plan_sigma_chain(P, tab);
    P.out = K; P.ldo = nm2; P.slab = (long long)nm1 * nm2;
    run_pairs(MODE_L1, EPI_LOCAL, {g}, P, ws);
}

// max_{i in A, j in B} |a_i - b_j|_p without touching most pairs (get_kernel_width / get_dsmax: the reference forms
// the whole distance matrix, aqml.py:220-237).  d(i,j) <= r_i + r_j for the distances r to ANY common point (here
// the column mean of B); one farthest-point sweep (the row farthest from the mean, its farthest partner, that
// row's farthest partner) gives a pair distance L that is attained.  Only rows with r_i + max r_B >= L can take
// part in the maximum (aSLATM: ~5 % of the H / C rows -- the norms are heavy-tailed), and among those, sorted by
// r, whole tiles are skipped by the same test.  The surviving pairs go through the ordinary pair-tile kernel, so
// the result is the same number the unpruned pass produces.  Operands: packed rows (pitch ld, zero padded
// columns).  Returns d^2 (p = 2) or the L1 distance (p = 1).
static double pruned_max_segs(int p, const std::vector<RowSeg> &SA, const std::vector<RowSeg> &SB, long long ld, bool same,
                              Workspace &ws) {
    cudaStream_t st = ws.stream();
    auto total = [](const std::vector<RowSeg> &S) { long long n = 0; for (auto &s : S) n += s.n; return n; };
    const long long nA = total(SA), nB = total(SB);
    if (nA <= 0 || nB <= 0) return 0.0;
    AQML_REQUIRE(nA < (1ll << 31) && nB < (1ll << 31), "too many rows");
    const int ncols = (int)ld;
    // reference point: column mean over B
    double *sum = ws.zeros<double>(ncols);
    unsigned *nzs = ws.zeros<unsigned>(ncols);
    for (auto &sg : SB) launch_colstats(sg.x, ld, nullptr, sg.n, ncols, sum, nzs, st);
    double *mean = ws.alloc<double>(ncols);
    scale_gather_kernel<<<(ncols + 255) / 256, 256, 0, st>>>(sum, nullptr, ncols, 1.0 / (double)nB, mean);
    AQML_LAUNCH_CHECK();
    auto sweep = [&](const std::vector<RowSeg> &S, long long n, const double *c) {
        double *r = ws.alloc<double>((size_t)n);
        long long o = 0;
        for (auto &sg : S) {
            if (sg.n > 0) {
                row_dist_kernel<<<(sg.n + 7) / 8, 256, 0, st>>>(p, sg.x, ld, sg.n, c, ncols, r + o);
                AQML_LAUNCH_CHECK();
            }
            o += sg.n;
        }
        std::vector<double> h((size_t)n);
        AQML_CUDA(cudaMemcpyAsync(h.data(), r, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
        AQML_CUDA(cudaStreamSynchronize(st));
        return h;
    };
    auto rowptr = [&](const std::vector<RowSeg> &S, long long i) -> const double * {
        for (auto &sg : S) {
            if (i < sg.n) return sg.x + i * ld;
            i -= sg.n;
        }
        return nullptr;
    };
    auto argmax = [](const std::vector<double> &v) { return (long long)(std::max_element(v.begin(), v.end()) - v.begin()); };
    std::vector<double> rA = sweep(SA, nA, mean);
    std::vector<double> rB = same ? rA : sweep(SB, nB, mean);
    for (double v : rA) if (!(v == v)) return std::numeric_limits<double>::quiet_NaN();  // NaN rows: the reference's max is NaN too
    for (double v : rB) if (!(v == v)) return std::numeric_limits<double>::quiet_NaN();
    const long long i0 = argmax(rA);
    std::vector<double> t1 = sweep(SB, nB, rowptr(SA, i0));
    const long long j1 = argmax(t1);
    std::vector<double> t2 = sweep(SA, nA, rowptr(SB, j1));
    const double L = std::max(t1[j1], t2[argmax(t2)]) * (1.0 - 1e-9);
    const double rAmax = rA[i0], rBmax = rB[argmax(rB)];
    auto select = [&](const std::vector<RowSeg> &S, const std::vector<double> &r, double partner_max, std::vector<const double *> &ptrs,
                      std::vector<double> &bound) {
        std::vector<long long> rows;
        for (long long i = 0; i < (long long)r.size(); i++)
            if (r[i] + partner_max >= L) rows.push_back(i);
        std::sort(rows.begin(), rows.end(), [&](long long a, long long b) { return r[a] > r[b] || (r[a] == r[b] && a < b); });
        const size_t padded = (size_t)round_up((int64_t)rows.size(), BM);
        bound.assign(padded, -1.0);
        ptrs.assign(padded, nullptr);
        for (size_t k = 0; k < rows.size(); k++) { bound[k] = r[rows[k]]; ptrs[k] = rowptr(S, rows[k]); }
        return (int)rows.size();
    };
    std::vector<const double *> ca, cb;
    std::vector<double> ba, bb;
    const int nca = select(SA, rA, rBmax, ca, ba);
    const int ncb = same ? nca : select(SB, rB, rAmax, cb, bb);
    if (nca == 0 || ncb == 0) return (p == 2) ? L * L : L;  // cannot happen (i0 / j1 always qualify); be safe
    auto gather = [&](const std::vector<const double *> &ptrs, double *&px, double *&pn) {
        px = ws.alloc<double>(ptrs.size() * (size_t)ld);
        pn = (p == 2) ? ws.alloc<double>(ptrs.size()) : nullptr;
        PackJob J{};
        J.srcrows = ws.upload(ptrs); J.colmap = nullptr; J.mean = (p == 2) ? mean : nullptr;
        J.dst = px; J.ldd = ld; J.norm = pn; J.ncols = ncols;
        launch_pack(J, (int)ptrs.size(), st);
    };
    double *pa, *na_, *pb, *nb_;
    gather(ca, pa, na_);
    if (same) { pb = pa; nb_ = na_; } else gather(cb, pb, nb_);
    PairGroup g = blank_group();
    g.A = pa; g.B = pb; g.normA = na_; g.normB = nb_; g.ld = ld; g.nA = nca; g.nB = ncb;
    g.tilesM = (int)(ca.size() / BM); g.tilesN = (int)((same ? ca.size() : cb.size()) / BN);
    g.sym = same ? 1 : 0;
    g.boundA = ws.upload(ba); g.boundB = same ? g.boundA : ws.upload(bb); g.bound_min = L;
    unsigned long long *bits = ws.zeros<unsigned long long>(1);
    PairParams P{};
    P.out = reinterpret_cast<double *>(bits);
    run_pairs(p == 2 ? MODE_DOT : MODE_L1, EPI_MAX, {g}, P, ws);
    double v = 0.0;
    AQML_CUDA(cudaMemcpyAsync(&v, bits, 8, cudaMemcpyDeviceToHost, st));
    AQML_CUDA(cudaStreamSynchronize(st));
    return v;
}

static double pruned_max(int p, const double *A, int nA, const double *B, int nB, long long ld, bool same, Workspace &ws) {
    return pruned_max_segs(p, {RowSeg{A, nA}}, {RowSeg{B, nB}}, ld, same, ws);
}

double pruned_max_segs_public(int p, const std::vector<RowSeg> &SA, const std::vector<RowSeg> &SB, long long ld, bool same,
                              Workspace &ws) {
    return pruned_max_segs(p, SA, SB, ld, same, ws);
}

double pruned_max_public(int p, const double *A, int nA, const double *B, int nB, long long ld, bool same, Workspace &ws) {
    return pruned_max(p, A, nA, B, nB, ld, same, ws);
}

// per-element max distance, algo/aqml.py:204-255 (cab=False)
static void kernel_width(int p, const double *x, long long ldx, const int *zs, int A, int D, int zmax, double *dso,
                         cudaStream_t st) {
    AQML_REQUIRE(x && zs && dso && A >= 1 && D >= 1 && zmax >= 1, "bad argument");
    AQML_REQUIRE(p == 1 || p == 2, "p must be 1 or 2");
    require_device();
    for (int i = 0; i < A; i++) AQML_REQUIRE(zs[i] >= 1 && zs[i] <= zmax, "zs[%d]=%d outside 1..zmax", i, zs[i]);
    Workspace ws(st);
    std::vector<std::vector<int>> rz(zmax + 1);
    for (int i = 0; i < A; i++) rz[zs[i]].push_back(i);
    std::vector<int> els;
    for (int z = 1; z <= zmax; z++) if (!rz[z].empty()) els.push_back(z);
    const int ne = (int)els.size();
    unsigned *colnz = ws.zeros<unsigned>((size_t)ne * D);
    double *colsum = ws.zeros<double>((size_t)ne * D);
    std::vector<int *> d_rows(ne);
    for (int e = 0; e < ne; e++) {
        d_rows[e] = ws.upload(rz[els[e]]);
        launch_colstats(x, ldx, d_rows[e], (int)rz[els[e]].size(), D, colsum + (size_t)e * D, colnz + (size_t)e * D, st);
    }
    std::vector<unsigned> nz((size_t)ne * D);
    AQML_CUDA(cudaMemcpyAsync(nz.data(), colnz, sizeof(unsigned) * nz.size(), cudaMemcpyDeviceToHost, st));
    AQML_CUDA(cudaStreamSynchronize(st));
    std::vector<double> mx(ne);
    for (int e = 0; e < ne; e++) {
        std::vector<int> colmap;
        for (int c = 0; c < D; c++) if (nz[(size_t)e * D + c]) colmap.push_back(c);
        if (colmap.empty()) colmap.push_back(0);
        int Dz = (int)colmap.size(), n = (int)rz[els[e]].size();
        long long ld = round_up(Dz, BK);
        int tiles = (n + BM - 1) / BM;
        int *d_colmap = ws.upload(colmap);
        std::vector<int> rm(rz[els[e]]);
        rm.resize((size_t)tiles * BM, -1);
        double *px = ws.alloc<double>(rm.size() * (size_t)ld);
        PackJob J{};
        J.src = x; J.lds = ldx; J.rowmap = ws.upload(rm); J.colmap = d_colmap; J.mean = nullptr;
        J.dst = px; J.ldd = ld; J.norm = nullptr; J.ncols = Dz;
        launch_pack(J, (int)rm.size(), st);
        mx[e] = pruned_max(p, px, n, px, n, ld, true, ws);
    }
    for (int z = 0; z < zmax; z++) dso[z] = 1.e-9;
    for (int e = 0; e < ne; e++) dso[els[e] - 1] = (p == 2) ? std::sqrt(mx[e]) : mx[e];
}

}  // namespace aqml

// =============================================================================================
// C ABI
// =============================================================================================
using namespace aqml;

extern "C" {

int aqml_version(void) { return 100; }
const char *aqml_last_error(void) { return last_error_ref().c_str(); }
int aqml_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int64_t aqml_launch_count(void) { return (int64_t)g_launches.load(); }

void aqml_tile_shape(int *bm, int *bn, int *bk) {
    if (bm) *bm = BM;
    if (bn) *bn = BN;
    if (bk) *bk = BK;
}

int aqml_pair_stats(int64_t *executed_ktiles, int64_t *dense_ktiles, int reset) {
    return guarded([&] {
        require_device();
        unsigned long long *buf = pair_stats_buffer();
        unsigned long long h[2] = {0, 0};
        if (buf) {
            AQML_CUDA(cudaDeviceSynchronize());
            AQML_CUDA(cudaMemcpy(h, buf, 16, cudaMemcpyDeviceToHost));
            if (reset) AQML_CUDA(cudaMemset(buf, 0, 16));
        }
        if (executed_ktiles) *executed_ktiles = (int64_t)h[0];
        if (dense_ktiles) *dense_ktiles = (int64_t)h[1];
    });
}

int aqml_fgaussian_kernel(const double *a, int na, const double *b, int nb, int D, double *k, double sigma) {
    return guarded([&] {
        AQML_REQUIRE(sigma != 0.0, "sigma is zero");
        global_host(AQML_KERNEL_GAUSSIAN, EPI_EXP_STORE, a, na, b, nb, D, k, sigma);
    });
}
int aqml_flaplacian_kernel(const double *a, int na, const double *b, int nb, int D, double *k, double sigma) {
    return guarded([&] {
        AQML_REQUIRE(sigma != 0.0, "sigma is zero");
        global_host(AQML_KERNEL_LAPLACIAN, EPI_EXP_STORE, a, na, b, nb, D, k, sigma);
    });
}
int aqml_fl2_distance(const double *A, const double *B, int n1, int n2, int nc, double *Dm) {
    return guarded([&] { global_host(AQML_KERNEL_GAUSSIAN, EPI_DIST_STORE, A, n1, B, n2, nc, Dm, 1.0); });
}
int aqml_fl1_distance(const double *A, const double *B, int n1, int n2, int nc, double *Dm) {
    return guarded([&] { global_host(AQML_KERNEL_LAPLACIAN, EPI_DIST_STORE, A, n1, B, n2, nc, Dm, 1.0); });
}
int aqml_flinear_kernel(const double *a, int na, const double *b, int nb, int D, double *k) {
    return guarded([&] { global_host(AQML_KERNEL_GAUSSIAN, EPI_DOT_STORE, a, na, b, nb, D, k, 1.0); });
}
int aqml_distance_dev(int p, const double *a, int na, int64_t lda, const double *b, int nb, int64_t ldb, int D, double *out,
                      void *stream) {
    return guarded([&] {
        AQML_REQUIRE(p == 0 || p == 1 || p == 2, "p must be 0 (dot product), 1 (L1) or 2 (L2)");
        AQML_REQUIRE(a && b && out, "null pointer");
        global_pairs(p == 1 ? AQML_KERNEL_LAPLACIAN : AQML_KERNEL_GAUSSIAN, p == 0 ? EPI_DOT_STORE : EPI_DIST_STORE, a, na, lda,
                     b, nb, ldb, D, nullptr, 0, 1, out, (cudaStream_t)stream);
    });
}

int aqml_global_kernel_dev(int kind, const double *a, int na, int64_t lda, const double *b, int nb, int64_t ldb, int D,
                           const double *sigmas, int nsig, int center, double *K, void *stream) {
    return guarded([&] {
        AQML_REQUIRE(kind == AQML_KERNEL_GAUSSIAN || kind == AQML_KERNEL_LAPLACIAN, "bad kernel kind %d", kind);
        AQML_REQUIRE(a && b && K && sigmas, "null pointer");
        global_pairs(kind, EPI_EXP_STORE, a, na, lda, b, nb, ldb, D, sigmas, nsig, center, K, (cudaStream_t)stream);
    });
}

int aqml_rows_colsum_dev(const double *x, int64_t ld, int n, int ncols, double *colsum, void *stream) {
    return guarded([&] {
        AQML_REQUIRE(x && colsum && n >= 0 && ncols >= 1 && ld >= ncols, "bad argument");
        require_device();
        if (n == 0) return;
        cudaStream_t st = (cudaStream_t)stream;
        Workspace ws(st);
        launch_colstats(x, ld, nullptr, n, ncols, colsum, ws.zeros<unsigned>(ncols), st);
    });
}

int aqml_rows_dist_dev(int p, const double *x, int64_t ld, int n, const double *c, int ncols, double *r, void *stream) {
    return guarded([&] {
        AQML_REQUIRE((p == 1 || p == 2) && x && c && r && n >= 0 && ncols >= 1 && ld >= ncols, "bad argument");
        require_device();
        if (n == 0) return;
        row_dist_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(p, x, ld, n, c, ncols, r);
        AQML_LAUNCH_CHECK();
    });
}

int aqml_max_distance_dev(int p, const double *a, int na, int64_t lda, const double *b, int nb, int64_t ldb, int D,
                          double *dmax, void *stream) {
    return guarded([&] {
        AQML_REQUIRE(p == 1 || p == 2, "p must be 1 or 2");
        AQML_REQUIRE(a && b && dmax && na >= 1 && nb >= 1, "bad argument");
        cudaStream_t st = (cudaStream_t)stream;
        require_device();
        unsigned long long *bits = nullptr;
        AQML_CUDA(cudaMallocAsync((void **)&bits, 8, st));
        AQML_CUDA(cudaMemsetAsync(bits, 0, 8, st));
        try {
            global_pairs(p == 2 ? AQML_KERNEL_GAUSSIAN : AQML_KERNEL_LAPLACIAN, EPI_MAX, a, na, lda, b, nb, ldb, D, nullptr, 0,
                         1, reinterpret_cast<double *>(bits), st);
            double v = 0.0;
            AQML_CUDA(cudaMemcpyAsync(&v, bits, 8, cudaMemcpyDeviceToHost, st));
            AQML_CUDA(cudaStreamSynchronize(st));
            *dmax = (p == 2) ? std::sqrt(v) : v;
        } catch (...) {
            cudaFreeAsync(bits, st);
            throw;
        }
        cudaFreeAsync(bits, st);
    });
}

int aqml_calc_lk_gaussian_dev(const double *x1, int64_t ldx1, const double *x2, int64_t ldx2, const int *zs1,
                              const int *zs2, const int *nas1, const int *nas2, int zmax, int iab, const double *sigmas,
                              int nm1, int nm2, int nsigmas, int D, int center, double *kernels, void *stream) {
    return guarded([&] {
        local_gaussian(x1, ldx1, x2, ldx2, zs1, zs2, nas1, nas2, zmax, iab, sigmas, nm1, nm2, nsigmas, D, center, kernels,
                       (cudaStream_t)stream);
    });
}

int aqml_calc_lk_gaussian(const double *x1, const double *x2, const int *zs1, const int *zs2, const int *nas1,
                          const int *nas2, int zmax, int iab, const double *sigmas, int nm1, int nm2, int nsigmas, int D,
                          double *kernels) {
    return guarded([&] {
        AQML_REQUIRE(x1 && x2 && nas1 && nas2 && kernels, "null pointer");
        AQML_REQUIRE(nm1 >= 0 && nm2 >= 0 && nsigmas >= 1 && D >= 1, "bad shape");
        require_device();
        int A1, A2;
        mol_index(nas1, nm1, A1);
        mol_index(nas2, nm2, A2);
        cudaStream_t st = 0;
        Workspace ws(st);
        const bool same = (x1 == x2 && A1 == A2);
        double *d1 = ws.upload(x1, (size_t)A1 * D);
        double *d2 = same ? d1 : ws.upload(x2, (size_t)A2 * D);
        size_t nk = (size_t)nsigmas * nm1 * nm2;
        double *dk = ws.alloc<double>(nk);
        local_gaussian(d1, D, d2, D, zs1, zs2, nas1, nas2, zmax, iab, sigmas, nm1, nm2, nsigmas, D, 1, dk, st);
        AQML_CUDA(cudaMemcpyAsync(kernels, dk, sizeof(double) * nk, cudaMemcpyDeviceToHost, st));
        AQML_CUDA(cudaStreamSynchronize(st));
    });
}

int aqml_calc_lk_laplacian_dev(const double *x1, int64_t ldx1, const double *x2, int64_t ldx2, const int *nas1,
                               const int *nas2, const double *sigmas, int nm1, int nm2, int nsigmas, int D, double *kernels,
                               void *stream) {
    return guarded([&] {
        local_laplacian(x1, ldx1, x2, ldx2, nas1, nas2, sigmas, nm1, nm2, nsigmas, D, kernels, (cudaStream_t)stream);
    });
}

int aqml_calc_lk_laplacian(const double *x1, const double *x2, const int *nas1, const int *nas2, const double *sigmas,
                           int nm1, int nm2, int nsigmas, int D, double *kernels) {
    return guarded([&] {
        AQML_REQUIRE(x1 && x2 && nas1 && nas2 && kernels, "null pointer");
        AQML_REQUIRE(nm1 >= 0 && nm2 >= 0 && nsigmas >= 1 && D >= 1, "bad shape");
        require_device();
        int A1, A2;
        mol_index(nas1, nm1, A1);
        mol_index(nas2, nm2, A2);
        cudaStream_t st = 0;
        Workspace ws(st);
        const bool same = (x1 == x2 && A1 == A2);
        double *d1 = ws.upload(x1, (size_t)A1 * D);
        double *d2 = same ? d1 : ws.upload(x2, (size_t)A2 * D);
        size_t nk = (size_t)nsigmas * nm1 * nm2;
        double *dk = ws.alloc<double>(nk);
        local_laplacian(d1, D, d2, D, nas1, nas2, sigmas, nm1, nm2, nsigmas, D, dk, st);
        AQML_CUDA(cudaMemcpyAsync(kernels, dk, sizeof(double) * nk, cudaMemcpyDeviceToHost, st));
        AQML_CUDA(cudaStreamSynchronize(st));
    });
}

int aqml_kernel_width_dev(int p, const double *x, int64_t ldx, const int *zs, int A, int D, int zmax, double *dso,
                          void *stream) {
    return guarded([&] { kernel_width(p, x, ldx, zs, A, D, zmax, dso, (cudaStream_t)stream); });
}

}  // extern "C"
