// Pair-wise / bond aSLATM (last "next" row): atom rows and bond rows of `calc_local_spectrum`,
// coreml/cml/representation/fslatm_pairwise.f90:525-641, with its own two- and three-body terms
// (`calc_sbop_local` :372-434, `calc_sbot_local` :191-301 -- they differ from fslatm.f90: weights w2 / w3, free powers, no
// 1/2 and 1/3 prefactors, the atom's nuclear charge in the one-body slot, every ordered (i, k) neighbour combination in the
// three-body sum, separate cut-offs for atoms and bonds).  The file is not part of the reference's build and cannot run as
// shipped; the semantics are those of oracle/np_pairwise.py (PARITY UNPINNED, see its header).
//
// One CTA per output row -- an atom, or a (atom, neighbour slot) bond.  The molecule sits in shared memory.  For a three-body
// type the candidate (i, k) combinations are laid out in reference order (i ascending, k ascending); the threads compute the
// geometry of a chunk of them (one combination per thread: angle at the centre, the two other cosines, the distance product,
// with the oracle's operation order and no FMA contraction in the cut-off decisions and angles), then every thread owns grid
// points and walks the chunk in order: one owner and one summation order per output element, as in the reference loop.
// This is a "next"-row kernel for molecules of tens of atoms (density-matrix learning in xb.py), one exp per term and grid
// point; it is FP64-ALU bound and makes no attempt at the aSLATM work-unit kernel's recurrences.
#include "common.cuh"

#include <cmath>
#include <vector>

namespace aqml {
namespace {

constexpr int PW_THREADS = 128, PW_CHUNK = 128;

struct PwParams {
    const double *coords;   // (A, 3)
    const int *zs;          // (A)
    const int *mol_off;     // (nmol + 1)
    const int *atom_mol;    // (A)
    const int *nbrs;        // (A, nbmax) molecule-local indices, -1 padded (bond rows)
    int A, nbmax, na_max;
    const int *mbs1, *mbs2, *mbs3;
    int n1, n2, n3;
    int nsa, nsb, ns3;      // grid points: two-body (atom rows), two-body (bond rows), three-body
    const double *xs_a, *xs0_a, *xs_b, *xs0_b, *xs3, *cos3;   // grids from the host libm: xs, w2 c'/(xs**rpower2) dgrid (without z1 z2), angles, cos
    double racut, rbcut, inv_sigma, c3, rpower3;               // c3 = w3 dgrid / sqrt(2 sigma^2 pi)
    double *ys, *ys2;       // (A, n), (A, nbmax, nu)
    int n, nu;
};

__device__ __forceinline__ double dist2_rn(const double *a, const double *b) {
    const double dx = __dsub_rn(a[0], b[0]), dy = __dsub_rn(a[1], b[1]), dz = __dsub_rn(a[2], b[2]);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
// cosine of the angle at b (fslatm_pairwise.f90:51-63)
__device__ __forceinline__ double cos_angle_rn(const double *a, const double *b, const double *c) {
    double v1x = __dsub_rn(a[0], b[0]), v1y = __dsub_rn(a[1], b[1]), v1z = __dsub_rn(a[2], b[2]);
    double v2x = __dsub_rn(c[0], b[0]), v2y = __dsub_rn(c[1], b[1]), v2z = __dsub_rn(c[2], b[2]);
    const double n1 = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(v1x, v1x), __dmul_rn(v1y, v1y)), __dmul_rn(v1z, v1z)));
    const double n2 = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(v2x, v2x), __dmul_rn(v2y, v2y)), __dmul_rn(v2z, v2z)));
    v1x = __ddiv_rn(v1x, n1); v1y = __ddiv_rn(v1y, n1); v1z = __ddiv_rn(v1z, n1);
    v2x = __ddiv_rn(v2x, n2); v2y = __ddiv_rn(v2y, n2); v2z = __ddiv_rn(v2z, n2);
    return __dadd_rn(__dadd_rn(__dmul_rn(v1x, v2x), __dmul_rn(v1y, v2y)), __dmul_rn(v1z, v2z));
}

// blockIdx.x < A: atom row; otherwise bond row (atom, slot)
__global__ void __launch_bounds__(PW_THREADS) pairwise_spectrum_kernel(const PwParams P) {
    extern __shared__ __align__(16) double pw_sm[];
    double *mc = pw_sm;                                        // [na][3]
    double *g_ang = mc + 3 * P.na_max, *g_cak = g_ang + PW_CHUNK, *g_cai = g_cak + PW_CHUNK, *g_rp = g_cai + PW_CHUNK;
    int *mz = reinterpret_cast<int *>(g_rp + PW_CHUNK);        // [na]
    int *l1 = mz + P.na_max, *l3 = l1 + P.na_max;              // candidate lists of a three-body type
    int *g_ok = l3 + P.na_max;                                 // [PW_CHUNK]
    __shared__ int s_n1, s_n3;

    const int tid = threadIdx.x;
    const bool bond = blockIdx.x >= P.A;
    const int atom = bond ? (blockIdx.x - P.A) / P.nbmax : blockIdx.x, slot = bond ? (blockIdx.x - P.A) % P.nbmax : 0;
    const int mol = P.atom_mol[atom], o = P.mol_off[mol], na = P.mol_off[mol + 1] - o, ia = atom - o;
    int ja = -1;
    if (bond) {
        // the reference walks the first count(nbrs > -1) slots of the atom (:605) and skips the atom itself
        int cnt = 0;
        for (int q = 0; q < P.nbmax; q++) cnt += (P.nbrs[(long long)atom * P.nbmax + q] > -1);
        ja = (slot < cnt) ? P.nbrs[(long long)atom * P.nbmax + slot] : -1;
        if (ja == ia || ja < 0 || ja >= na) return;          // the row stays zero (pre-cleared)
    }
    for (int i = tid; i < na; i += PW_THREADS) {
        mc[3 * i] = P.coords[3 * (long long)(o + i)];
        mc[3 * i + 1] = P.coords[3 * (long long)(o + i) + 1];
        mc[3 * i + 2] = P.coords[3 * (long long)(o + i) + 2];
        mz[i] = P.zs[o + i];
    }
    __syncthreads();
    const int zi = mz[ia];
    double *row = bond ? P.ys2 + ((long long)atom * P.nbmax + slot) * P.nu : P.ys + (long long)atom * P.n;
    const int ns2 = bond ? P.nsb : P.nsa;
    const double *xs2 = bond ? P.xs_b : P.xs_a, *xs02 = bond ? P.xs0_b : P.xs0_a;
    const double cut2 = bond ? P.rbcut : P.racut;

    // one-body (atom rows only; the bond rows' first n1 slots stay zero, :597-620)
    if (!bond)
        for (int m = tid; m < P.n1; m += PW_THREADS)
            if (P.mbs1[m] == zi) row[m] = (double)zi;

    // two-body
    for (int m = 0; m < P.n2; m++) {
        const int z1 = P.mbs2[2 * m], z2 = P.mbs2[2 * m + 1];
        if (zi != z1) continue;
        if (bond && mz[ja] != z2) continue;
        const double zz = (double)((z1 % 1000) * (z2 % 1000));
        for (int g = tid; g < ns2; g += PW_THREADS) {
            const double x = xs2[g], w = zz * xs02[g];
            double acc = 0.0;
            if (bond) {
                const double r = dist2_rn(mc + 3 * ia, mc + 3 * ja);
                if (r < cut2 * cut2) { const double t = x - sqrt(r); acc += w * exp(P.inv_sigma * (t * t)); }
            } else {
                for (int j = 0; j < na; j++) {
                    if (mz[j] != z2 || j == ia) continue;
                    const double r = dist2_rn(mc + 3 * ia, mc + 3 * j);
                    if (r < cut2 * cut2) { const double t = x - sqrt(r); acc += w * exp(P.inv_sigma * (t * t)); }
                }
            }
            row[P.n1 + m * ns2 + g] = acc;
        }
    }

    // three-body: centre ia of type z2, i of type z1 (the bond partner for a bond row), k of type z3
    for (int m = 0; m < P.n3; m++) {
        const int z1 = P.mbs3[3 * m], z2 = P.mbs3[3 * m + 1], z3 = P.mbs3[3 * m + 2];
        if (zi != z2) continue;
        if (bond && mz[ja] != z1) continue;
        __syncthreads();
        if (tid == 0) {
            int c1 = 0, c3 = 0;
            if (bond) l1[c1++] = ja;
            else
                for (int i = 0; i < na; i++)
                    if (mz[i] == z1 && i != ia) l1[c1++] = i;
            for (int i = 0; i < na; i++)
                if (mz[i] == z3 && i != ia) l3[c3++] = i;
            s_n1 = c1;
            s_n3 = c3;
        }
        __syncthreads();
        const int c1 = s_n1, c3 = s_n3, total = c1 * c3;
        const double c0 = P.c3 * (double)((z1 % 1000) * (z2 % 1000) * (z3 % 1000));
        constexpr int GP = 4;   // grid points per thread and pass (ns3 <= GP * PW_THREADS is checked on the host)
        double acc[GP];
#pragma unroll
        for (int q = 0; q < GP; q++) acc[q] = 0.0;
        for (int base = 0; base < total; base += PW_CHUNK) {
            __syncthreads();
            const int p = base + tid;
            if (tid < PW_CHUNK) {
                int ok = 0;
                if (p < total) {
                    const int i = l1[p / c3], k = l3[p % c3];
                    const double dij = __dsqrt_rn(dist2_rn(mc + 3 * i, mc + 3 * ia)), djk = __dsqrt_rn(dist2_rn(mc + 3 * ia, mc + 3 * k));
                    const double eps = 2.220446049250313e-16;
                    if (k != i && dij > eps && dij <= cut2 && djk > eps && djk <= P.racut) {
                        ok = 1;
                        double ca = cos_angle_rn(mc + 3 * i, mc + 3 * ia, mc + 3 * k);
                        ca = fmin(1.0, fmax(-1.0, ca));
                        g_ang[tid] = acos(ca);
                        g_cak[tid] = cos_angle_rn(mc + 3 * i, mc + 3 * k, mc + 3 * ia);
                        g_cai[tid] = cos_angle_rn(mc + 3 * k, mc + 3 * i, mc + 3 * ia);
                        const double dik = __dsqrt_rn(dist2_rn(mc + 3 * i, mc + 3 * k));
                        g_rp[tid] = pow(__dmul_rn(__dmul_rn(dij, dik), djk), P.rpower3);
                    }
                }
                g_ok[tid] = ok;
            }
            __syncthreads();
            const int cnt = min(PW_CHUNK, total - base);
#pragma unroll
            for (int q = 0; q < GP; q++) {
                const int g = tid + q * PW_THREADS;
                if (g >= P.ns3) continue;
                const double x = P.xs3[g], cx = P.cos3[g] * c0;
                double a = acc[q];
                for (int e = 0; e < cnt; e++) {
                    if (!g_ok[e]) continue;
                    const double t = x - g_ang[e];
                    a += (c0 + cx * g_cak[e] * g_cai[e]) / g_rp[e] * exp((t * t) * P.inv_sigma);
                }
                acc[q] = a;
            }
        }
#pragma unroll
        for (int q = 0; q < GP; q++) {
            const int g = tid + q * PW_THREADS;
            if (g < P.ns3) row[P.n1 + P.n2 * ns2 + m * P.ns3 + g] = acc[q];
        }
    }
}

std::vector<double> linspace(double x0, double x1, int nx) {   // fslatm_pairwise.f90:12-28
    std::vector<double> xs(nx);
    const double step = (x1 - x0) / (nx - 1);
    for (int i = 0; i < nx; i++) xs[i] = x0 + i * step;
    return xs;
}

const double PW_R0 = (double)0.1f;   // `real*8, parameter :: r0 = 0.1` (:5): a single-precision literal
const double PW_PI = 4.0 * std::atan(1.0);

void pw_dims(const double *rscut, double dgrid, int n1, int n2, int n3, int *nsx, int *n, int *nu) {
    AQML_REQUIRE(rscut && dgrid > 0 && rscut[0] > PW_R0 && rscut[1] > PW_R0, "bad cut-off / grid step");
    nsx[0] = (int)((rscut[0] - PW_R0) / dgrid) + 1;
    nsx[1] = (int)((rscut[1] - PW_R0) / dgrid) + 1;
    nsx[2] = (int)((PW_PI + 40.0 * (PW_PI / 180.0)) / dgrid) + 1;
    *n = n1 + n2 * nsx[0] + n3 * nsx[2];
    *nu = n1 + n2 * nsx[1] + n3 * nsx[2];
}

}  // namespace
}  // namespace aqml

using namespace aqml;

extern "C" {

int aqml_pairwise_dims(const double *rscut, double dgrid, int n1, int n2, int n3, int *nsx, int *n, int *nu) {
    return guarded([&] {
        AQML_REQUIRE(nsx && n && nu && n1 >= 0 && n2 >= 0 && n3 >= 0, "bad argument");
        pw_dims(rscut, dgrid, n1, n2, n3, nsx, n, nu);
    });
}

int aqml_pairwise_local_spectrum_dev(const double *coords, const int *zs, const int *nas, int nmol, const int *nbrs, int nbmax,
                                     const int *mbs1, int n1, const int *mbs2, int n2, const int *mbs3, int n3, const double *rscut,
                                     double dgrid, double sigma, double w2, double rpower2, double w3, double rpower3, double *ys,
                                     double *ys2, void *stream) {
    return guarded([&] {
        AQML_REQUIRE(coords && zs && nas && nmol >= 1 && ys && mbs1 && mbs2 && mbs3 && sigma > 0, "bad argument");
        AQML_REQUIRE(nbmax == 0 || (nbrs && ys2), "bond rows need the neighbour table and their output");
        require_device();
        cudaStream_t st = (cudaStream_t)stream;
        Workspace ws(st);
        int nsx[3], n = 0, nu = 0;
        pw_dims(rscut, dgrid, n1, n2, n3, nsx, &n, &nu);
        AQML_REQUIRE(nsx[0] >= 2 && nsx[1] >= 2 && nsx[2] >= 2, "grid with fewer than two points");
        AQML_REQUIRE(nsx[2] <= 4 * PW_THREADS, "three-body grid of %d points: more than %d are not supported", nsx[2], 4 * PW_THREADS);
        std::vector<int> off(nmol + 1, 0), amol;
        int na_max = 0;
        for (int i = 0; i < nmol; i++) {
            AQML_REQUIRE(nas[i] >= 1, "empty molecule");
            off[i + 1] = off[i] + nas[i];
            na_max = std::max(na_max, nas[i]);
            for (int a = 0; a < nas[i]; a++) amol.push_back(i);
        }
        const int A = off[nmol];
        PwParams P;
        memset(&P, 0, sizeof(P));
        P.coords = ws.upload(coords, (size_t)3 * A);
        P.zs = ws.upload(zs, (size_t)A);
        P.mol_off = ws.upload(off);
        P.atom_mol = ws.upload(amol);
        P.nbrs = nbmax ? ws.upload(nbrs, (size_t)A * nbmax) : nullptr;
        P.A = A; P.nbmax = nbmax; P.na_max = na_max;
        P.mbs1 = ws.upload(mbs1, (size_t)n1); P.mbs2 = ws.upload(mbs2, (size_t)2 * n2); P.mbs3 = ws.upload(mbs3, (size_t)3 * n3);
        P.n1 = n1; P.n2 = n2; P.n3 = n3;
        P.nsa = nsx[0]; P.nsb = nsx[1]; P.ns3 = nsx[2];
        // grids with the host libm, as the Fortran evaluates them (:418-424, :250-262)
        const double cn = 1.0 / std::sqrt(2 * sigma * sigma * PW_PI);
        auto grid2 = [&](int nx, std::vector<double> &xs, std::vector<double> &xs0) {
            xs = linspace(PW_R0, rscut[0], nx);   // both two-body grids end at rscut(1) (:418)
            xs0.resize(nx);
            for (int i = 0; i < nx; i++) xs0[i] = w2 * cn / std::pow(xs[i], rpower2) * dgrid;
        };
        std::vector<double> xa, x0a, xb, x0b, x3, c3;
        grid2(nsx[0], xa, x0a);
        grid2(nsx[1], xb, x0b);
        x3 = linspace(-20.0 * PW_PI / 180.0, PW_PI + 20.0 * PW_PI / 180.0, nsx[2]);
        c3.resize(nsx[2]);
        for (int i = 0; i < nsx[2]; i++) c3[i] = std::cos(x3[i]);
        P.xs_a = ws.upload(xa); P.xs0_a = ws.upload(x0a); P.xs_b = ws.upload(xb); P.xs0_b = ws.upload(x0b);
        P.xs3 = ws.upload(x3); P.cos3 = ws.upload(c3);
        P.racut = rscut[0]; P.rbcut = rscut[1];
        P.inv_sigma = -1.0 / (2 * sigma * sigma);
        P.c3 = w3 * dgrid * cn;
        P.rpower3 = rpower3;
        P.ys = ys; P.ys2 = ys2; P.n = n; P.nu = nu;
        AQML_CUDA(cudaMemsetAsync(ys, 0, (size_t)A * n * sizeof(double), st));
        if (nbmax) AQML_CUDA(cudaMemsetAsync(ys2, 0, (size_t)A * nbmax * nu * sizeof(double), st));
        const size_t smem = (size_t)(3 * na_max + 4 * PW_CHUNK) * sizeof(double) + (size_t)(3 * na_max + PW_CHUNK) * sizeof(int);
        AQML_REQUIRE(smem <= 200 * 1024, "molecule of %d atoms exceeds the shared-memory budget", na_max);
        AQML_CUDA(cudaFuncSetAttribute(pairwise_spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pairwise_spectrum_kernel<<<A + A * nbmax, PW_THREADS, smem, st>>>(P);
        AQML_LAUNCH_CHECK();
        AQML_CUDA(cudaStreamSynchronize(st));   // the host vectors behind the uploads go out of scope
    });
}

}  // extern "C"
