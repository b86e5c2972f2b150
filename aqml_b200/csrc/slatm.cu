// SLATM / aSLATM generation (replaces coreml/cml/representation/fslatm.f90 + the Python double loop of
// representation/core.py:430-554) and the fused element-packed representation (aqml_rep).
//
// One thread block per atom (local) or per (molecule, element) (global).  The block
//   1. loads the molecule, builds the centre's neighbour list sorted by element (list order kept),
//   2. two-body:   thread <-> (partner element, grid point) accumulates London terms in registers,
//   3. three-body: neighbour pairs are enumerated block by block (a block = one [z1,zc,z3] type),
//      their geometry (angle, two cosines, 1/r^3) is computed once into shared memory, then
//      thread <-> (block, grid point) accumulates the ATM terms over the block's pairs,
//   4. the spectrum (shared memory, laid out as the centre element's support) is written with
//      coalesced stores either into the dense (row, D) API layout or straight into the
//      element-packed operand the local-kernel GEMM consumes (+ |row|^2).
// Every output element has exactly one owner thread: no atomics, deterministic, and the
// accumulation order over pairs equals the reference's loop order (i ascending, k ascending).
#include "packing.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <cstdlib>

namespace aqml {

constexpr int MAXEL = 16;
constexpr int SL_THREADS = 128;
constexpr int PCH = 512;  // neighbour pairs per shared-memory chunk

constexpr int MAXBLK = MAXEL * (MAXEL + 1) / 2;
struct Block3 {
    int e1, e3, col, flip, p0;  // pair offset p0; pairs = n1*n3 or n1(n1-1)/2
    double c0;
};

struct SlatmDev {
    int n_el, nmb, nx2, nx3, D;
    double sigma2, sigma3, dgrid2, dgrid3, coeff2, coeff3, rcut, rpower;
    int el_label[MAXEL];
    double el_w[MAXEL];
    int pdim[MAXEL], pld[MAXEL];  // support size of each centre element in the current layout (+16-padded)
    const int *slot_kind, *slot_col, *slot_e0;  // per mbtype: 1|2|3, dense column, element index of a 1-body type
    const int *bop_slot;                        // [MAXEL*MAXEL] slot or -1 (symmetric)
    const int *bot_slot;                        // [MAXEL^3] (e1,ec,e3) -> 2*slot+flip or -1 (both orientations filled)
    const int *lay_col;                         // [n_el*nmb] column of slot in element e's layout or -1
    const int *owner;                           // [nmb] owning element of a slot in global mode
    const double *xs2, *xpow2, *xs3, *cosx3;    // grids (host libm): xs, xs**rpower, angles, cos(angles)
    const Block3 *blk;                          // [n_el*MAXBLK] three-body blocks (e1<=e3) per centre element
    int nblk[MAXEL];
    // Gaussian recurrence along the grid (exp(-(x+h-a)^2 c) = E*R, R' = R*q): 2 exp per (entry, 8-thread chunk)
    const double *inv2;                         // dgrid2 / xs2**rpower
    double h2, h3, q2, q3, cst2, cst3;          // grid steps, exp(-2 h^2 c), c = 1/(2 sigma^2)
};

struct SlatmJob {
    const double *coords;  // (A,3)
    const int *zs;         // (A) element labels
    const int *mol_off;    // (nmol+1)
    const int *atom_mol;   // (A)
    int nmol, A, local;
    // dense output
    double *out;
    long long ldo;
    // packed output (local only): per element destination
    const int *atom_prow;  // (A) packed row inside the element's buffer, -1 = skip
    double *px[MAXEL];
    double *pnorm[MAXEL];
    unsigned *pmask[MAXEL];  // [rows/64][maskw] non-zero k-tile bits of the packed rows (pre-zeroed) or null
    int maskw[MAXEL];
    // int8 digit planes of the packed rows (i8pair.cuh), written by the work-unit kernel right after the rows; null = not wanted
    int8_t *psl[MAXEL];
    int *pexpo[MAXEL];
};

// geometry with the oracle's operation order and no FMA contraction (decisions / angles are
// discontinuous or ill-conditioned, SURVEY 7 hard part 9)
__device__ __forceinline__ double dist2_rn(double ax, double ay, double az, double bx, double by, double bz) {
    double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
// cosine of the angle at b (utils.f90:82-103)
__device__ __forceinline__ double cos_angle_rn(const double *a, const double *b, const double *c) {
    double v1x = __dsub_rn(a[0], b[0]), v1y = __dsub_rn(a[1], b[1]), v1z = __dsub_rn(a[2], b[2]);
    double v2x = __dsub_rn(c[0], b[0]), v2y = __dsub_rn(c[1], b[1]), v2z = __dsub_rn(c[2], b[2]);
    double n1 = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(v1x, v1x), __dmul_rn(v1y, v1y)), __dmul_rn(v1z, v1z)));
    double n2 = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(v2x, v2x), __dmul_rn(v2y, v2y)), __dmul_rn(v2z, v2z)));
    v1x = __ddiv_rn(v1x, n1); v1y = __ddiv_rn(v1y, n1); v1z = __ddiv_rn(v1z, n1);
    v2x = __ddiv_rn(v2x, n2); v2y = __ddiv_rn(v2y, n2); v2z = __ddiv_rn(v2z, n2);
    return __dadd_rn(__dadd_rn(__dmul_rn(v1x, v2x), __dmul_rn(v1y, v2y)), __dmul_rn(v1z, v2z));
}

// dynamic shared memory carve-up (doubles first)
//   spec[pldmax] | mx,my,mz | nbx,nby,nbz,nbd,nbr2 | md,mr2 (each [na_max]) | pair[4*PCH] | ints: mel,nbi,nbe,nbf,mfl [na_max]
// General-purpose SLATM kernel: one exp per (entry, grid point), spectrum in shared memory, any grid size and
// molecule size.  Used for global SLATM of molecules with more than 36 atoms and for grids outside the
// recurrence's safe range; everything else runs slatm_units_kernel (slatm_units.cuh).
__global__ void __launch_bounds__(SL_THREADS) slatm_kernel(const SlatmDev T, const SlatmJob J, int pldmax, int na_max) {
    extern __shared__ __align__(16) double sm[];
    double *spec = sm;
    double *mx = spec + pldmax, *my = mx + na_max, *mz = my + na_max;
    double *nbx = mz + na_max, *nby = nbx + na_max, *nbz = nby + na_max, *nbd = nbz + na_max, *nbr2 = nbd + na_max;
    double *md = nbr2 + na_max, *mr2 = md + na_max;
    double *pair = mr2 + na_max;
    int *mel = reinterpret_cast<int *>(pair + 4 * PCH);
    int *nbi = mel + na_max, *nbe = nbi + na_max, *nbf = nbe + na_max, *mfl = nbf + na_max;

    __shared__ int s_cnt[MAXEL], s_start[MAXEL + 1];
    __shared__ Block3 s_blk[MAXBLK];
    __shared__ int s_nblk, s_ptot;
    __shared__ double s_red[SL_THREADS / 32];
    __shared__ unsigned s_mask[32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double eps = 2.220446049250313e-16;
    const double rcut = T.rcut, rcut2 = __dmul_rn(rcut, rcut);

    int m, ec, c_first, c_last;  // centre range inside the molecule
    if (J.local) {
        int a = blockIdx.x;
        m = J.atom_mol[a];
        c_first = a - J.mol_off[m];
        c_last = c_first + 1;
        ec = -2;  // from the atom
    } else {
        m = blockIdx.x / T.n_el;
        ec = blockIdx.x % T.n_el;
        c_first = 0;
        c_last = J.mol_off[m + 1] - J.mol_off[m];
    }
    const int o = J.mol_off[m], na = J.mol_off[m + 1] - o;

    for (int i = tid; i < na; i += SL_THREADS) {
        mx[i] = J.coords[3 * (long long)(o + i)];
        my[i] = J.coords[3 * (long long)(o + i) + 1];
        mz[i] = J.coords[3 * (long long)(o + i) + 2];
        int z = J.zs[o + i], e = -1;
        for (int q = 0; q < T.n_el; q++)
            if (T.el_label[q] == z) e = q;
        mel[i] = e;
    }
    __syncthreads();
    if (J.local) ec = mel[c_first];
    const int pld = (ec >= 0) ? T.pld[ec] : 0;
    for (int i = tid; i < pld; i += SL_THREADS) spec[i] = 0.0;

    if (ec >= 0) {
        const int *lay = T.lay_col + ec * T.nmb;
        for (int c = c_first; c < c_last; c++) {
            if (mel[c] != ec) continue;  // block-uniform
            __syncthreads();
            // ---- neighbour list, sorted by element, list order preserved ---------------------
            // pass A (all threads): distance / cut-off flags of every atom w.r.t. the centre
            for (int q = tid; q < na; q += SL_THREADS) {
                int fl = 0;
                double r2 = 0.0, d = 0.0;
                if (q != c && mel[q] >= 0) {
                    r2 = dist2_rn(mx[c], my[c], mz[c], mx[q], my[q], mz[q]);
                    d = __dsqrt_rn(r2);
                    fl = (int)(r2 < rcut2) | ((int)((d > eps) && (d <= rcut)) << 1);  // bit0: 2-body, bit1: 3-body
                }
                md[q] = d; mr2[q] = r2; mfl[q] = fl;
            }
            __syncthreads();
            // pass B (one warp per element): counts by ballot
            for (int e = warp; e < T.n_el; e += SL_THREADS / 32) {
                int n = 0;
                for (int q0 = 0; q0 < na; q0 += 32) {
                    int q = q0 + lane;
                    bool keep = (q < na) && mfl[q] && mel[q] == e;
                    n += __popc(__ballot_sync(0xffffffffu, keep));
                }
                if (lane == 0) s_cnt[e] = n;
            }
            __syncthreads();
            if (tid == 0) {
                int acc = 0;
                for (int e = 0; e < T.n_el; e++) { s_start[e] = acc; acc += s_cnt[e]; }
                s_start[T.n_el] = acc;
            }
            __syncthreads();
            // pass C: ordered compaction
            for (int e = warp; e < T.n_el; e += SL_THREADS / 32) {
                int w = s_start[e];
                for (int q0 = 0; q0 < na; q0 += 32) {
                    int q = q0 + lane;
                    bool keep = (q < na) && mfl[q] && mel[q] == e;
                    unsigned mk = __ballot_sync(0xffffffffu, keep);
                    if (keep) {
                        int pos = w + __popc(mk & ((1u << lane) - 1u));
                        nbx[pos] = mx[q]; nby[pos] = my[q]; nbz[pos] = mz[q]; nbd[pos] = md[q]; nbr2[pos] = mr2[q];
                        nbi[pos] = q; nbe[pos] = e; nbf[pos] = mfl[q];
                    }
                    w += __popc(mk);
                }
            }
            // ---- three-body block table: static per centre element (host built), pair offsets by scan ----
            {
                const int nb = T.nblk[ec];
                const Block3 *tab = T.blk + ec * MAXBLK;
                if (warp == 0) {
                    int carry = 0;
                    for (int b0 = 0; b0 < nb; b0 += 32) {
                        int b = b0 + lane, np = 0;
                        Block3 B;
                        if (b < nb) {
                            B = tab[b];
                            int n1 = s_cnt[B.e1], n3 = s_cnt[B.e3];
                            np = (B.e1 == B.e3) ? n1 * (n1 - 1) / 2 : n1 * n3;
                        }
                        int incl = np;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            int u = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += u;
                        }
                        if (b < nb) { B.p0 = carry + incl - np; s_blk[b] = B; }
                        carry += __shfl_sync(0xffffffffu, incl, 31);
                    }
                    if (lane == 0) { s_nblk = nb; s_ptot = carry; }
                }
            }
            __syncthreads();

            // ---- two-body: thread <-> (partner element, grid point) ---------------------------------
            {
                const double inv2 = -0.5 / (T.sigma2 * T.sigma2);
                const double pref = J.local ? 0.5 : 1.0;
                for (int oidx = tid; oidx < T.n_el * T.nx2; oidx += SL_THREADS) {
                    int e2 = oidx / T.nx2, x = oidx - e2 * T.nx2;
                    int sl = T.bop_slot[ec * MAXEL + e2];
                    if (sl < 0) continue;
                    if (!J.local && e2 < ec) continue;  // owned by the other element's block
                    int col = lay[sl];
                    if (col < 0) continue;
                    double c0 = pref * (T.el_w[ec] * T.el_w[e2]) * T.coeff2;
                    double w = c0 / T.xpow2[x] * T.dgrid2;
                    double xg = T.xs2[x];
                    double acc = 0.0;
                    for (int n = s_start[e2]; n < s_start[e2 + 1]; n++) {
                        if (!(nbf[n] & 1)) continue;
                        if (!J.local && e2 == ec && nbi[n] < c) continue;  // unordered pair counted once
                        double t = xg - sqrt(nbr2[n]);
                        acc += w * exp(inv2 * (t * t));
                    }
                    spec[col + x] += acc;
                }
            }

            // ---- three-body, chunked over the pair list ------------------------------------------
            const int ptot = s_ptot, nblk = s_nblk;
            const double inv3 = -1.0 / (2 * T.sigma3 * T.sigma3);
            const double cc[3] = {mx[c], my[c], mz[c]};
            for (int P0 = 0; P0 < max(ptot, 1); P0 += PCH) {  // runs once even without pairs (2-body items)
                const int P1 = min(ptot, P0 + PCH);
                __syncthreads();
                for (int p = P0 + tid; p < P1; p += SL_THREADS) {
                    int b = 0;
                    while (b + 1 < nblk && p >= s_blk[b + 1].p0) b++;
                    const Block3 B = s_blk[b];
                    int q = p - B.p0, i, k;
                    int n1 = s_cnt[B.e1], n3 = s_cnt[B.e3];
                    if (B.e1 == B.e3) {
                        i = 0;
                        while (q >= n1 - 1 - i) { q -= n1 - 1 - i; i++; }
                        k = i + 1 + q;
                    } else {
                        i = q / n3;
                        k = q - i * n3;
                    }
                    i += s_start[B.e1];
                    k += s_start[B.e3];
                    double ang = 0.0, cak = 0.0, cai = 0.0, rinv3 = 0.0;
                    if ((nbf[i] & 2) && (nbf[k] & 2)) {
                        // reference roles: "i" is the atom of the type's first element (flip swaps)
                        int I = B.flip ? k : i, K = B.flip ? i : k;
                        double pi[3] = {nbx[I], nby[I], nbz[I]}, pk[3] = {nbx[K], nby[K], nbz[K]};
                        double dik = __dsqrt_rn(dist2_rn(pi[0], pi[1], pi[2], pk[0], pk[1], pk[2]));
                        if (dik <= rcut) {
                            double ca = cos_angle_rn(pi, cc, pk);
                            ca = fmin(1.0, fmax(-1.0, ca));
                            ang = acos(ca);
                            cak = cos_angle_rn(pi, pk, cc);
                            cai = cos_angle_rn(pk, pi, cc);
                            double r = __dmul_rn(__dmul_rn(nbd[I], dik), nbd[K]);
                            rinv3 = 1.0 / (r * r * r);
                        }
                    }
                    double *pp = pair + 4 * (p - P0);
                    pp[0] = ang; pp[1] = cak; pp[2] = cai; pp[3] = rinv3;
                }
                __syncthreads();
                {
                for (int oidx = tid; oidx < nblk * T.nx3; oidx += SL_THREADS) {
                    int b = oidx / T.nx3, x = oidx - b * T.nx3;
                    const int q0 = max(s_blk[b].p0, P0);
                    const int q1 = min((b + 1 < nblk) ? s_blk[b + 1].p0 : ptot, P1);
                    if (q0 >= q1) continue;
                    const double c0 = s_blk[b].c0;
                    const double cosc = T.cosx3[x] * c0;  // cos_xs(i) = cos(xs(i)) * c0, fslatm.f90:338-340
                    const double xg = T.xs3[x];
                    double acc = 0.0;
                    for (int p = q0; p < q1; p++) {
                        const double *pp = pair + 4 * (p - P0);
                        double rinv3 = pp[3];
                        if (rinv3 == 0.0) continue;
                        double t = xg - pp[0];
                        acc += (c0 + cosc * pp[1] * pp[2]) * rinv3 * exp((t * t) * inv3);
                    }
                    spec[s_blk[b].col + x] += acc;
                }
                }
            }
        }
    }
    __syncthreads();

    // ---- output ---------------------------------------------------------------------------------
    if (J.local) {
        const int a = blockIdx.x;
        if (J.out) {
            double *row = J.out + (long long)a * J.ldo;
            const int z = J.zs[a];
            for (int s = 0; s < T.nmb; s++) {
                int kind = T.slot_kind[s], col = T.slot_col[s];
                if (kind == 1) {
                    if (tid == 0) row[col] = (ec >= 0 && T.slot_e0[s] == ec) ? (double)z : 0.0;
                } else {
                    int len = (kind == 2) ? T.nx2 : T.nx3;
                    int lc = (ec >= 0) ? T.lay_col[ec * T.nmb + s] : -1;
                    for (int x = tid; x < len; x += SL_THREADS) row[col + x] = (lc >= 0) ? spec[lc + x] : 0.0;
                }
            }
        }
        if (J.atom_prow && ec >= 0) {
            int pr = J.atom_prow[a];
            if (pr >= 0) {
                double *row = J.px[ec] + (long long)pr * pld;
                double ss = 0.0;
                const bool want_mask = J.pmask[ec] != nullptr;
                if (want_mask) {
                    if (tid < 32) s_mask[tid] = 0u;
                    __syncthreads();
                }
                for (int i = tid; i < pld; i += SL_THREADS) {
                    double v = spec[i];
                    row[i] = v;
                    ss += v * v;
                    if (want_mask && v != 0.0) atomicOr(&s_mask[i >> 9], 1u << ((i >> 4) & 31));
                }
                if (want_mask) {
                    __syncthreads();
                    if (tid < J.maskw[ec] && s_mask[tid])
                        atomicOr(J.pmask[ec] + (long long)(pr >> 6) * J.maskw[ec] + tid, s_mask[tid]);
                }
                ss = warp_sum(ss);
                if ((tid & 31) == 0) s_red[tid >> 5] = ss;
                __syncthreads();
                if (tid == 0) {
                    double t = 0.0;
                    for (int w = 0; w < SL_THREADS / 32; w++) t += s_red[w];
                    J.pnorm[ec][pr] = t;
                }
            }
        }
    } else {
        double *row = J.out + (long long)m * J.ldo;
        for (int s = 0; s < T.nmb; s++) {
            if (T.owner[s] != ec) continue;
            int kind = T.slot_kind[s], col = T.slot_col[s];
            if (kind == 1) {
                if (tid == 0) {
                    int cnt = 0;
                    for (int q = 0; q < na; q++) cnt += (mel[q] == ec);
                    row[col] = (double)(T.el_label[ec] * cnt);
                }
            } else {
                int len = (kind == 2) ? T.nx2 : T.nx3;
                int lc = T.lay_col[ec * T.nmb + s];
                for (int x = tid; x < len; x += SL_THREADS) row[col + x] = (lc >= 0) ? spec[lc + x] : 0.0;
            }
        }
    }
}

}  // namespace aqml
#include "i8digits.cuh"
#include "slatm_units.cuh"
#include "i8pair.cuh"
namespace aqml {

// K <- Kc + Kc^T (+ diag), in place, one 32x32 tile pair per block: Kc holds, for every unordered atom pair,
// the term once (at (mol_i, mol_j) in packed order), and nothing for i == j.
__global__ void __launch_bounds__(256) symmetrize_kernel(double *K, int n, int nsig, const double *diag) {
    __shared__ double ta[32][33], tb[32][33];
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (tj < ti) return;
    double *Ks = K + (long long)blockIdx.z * n * n;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        int a = ti * 32 + r, b = tj * 32 + tx;
        ta[r][tx] = (a < n && b < n) ? Ks[(long long)a * n + b] : 0.0;
        int a2 = tj * 32 + r, b2 = ti * 32 + tx;
        tb[r][tx] = (a2 < n && b2 < n) ? Ks[(long long)a2 * n + b2] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int a = ti * 32 + r, b = tj * 32 + tx;
        if (a < n && b < n) {
            double v = ta[r][tx] + tb[tx][r];
            if (a == b) v = 2.0 * ta[r][tx] + diag[a];
            Ks[(long long)a * n + b] = v;
        }
        if (ti != tj) {
            int a2 = tj * 32 + r, b2 = ti * 32 + tx;
            if (a2 < n && b2 < n) Ks[(long long)a2 * n + b2] = tb[r][tx] + ta[tx][r];
        }
    }
}

__global__ void zero_rows_kernel(double *x, long long ld, const int *rows, int nrows, double *norm) {
    int r = rows[blockIdx.x];
    for (long long i = threadIdx.x; i < ld; i += blockDim.x) x[(long long)r * ld + i] = 0.0;
    if (threadIdx.x == 0 && norm) norm[r] = 0.0;
}

// digit planes of rows that no aSLATM row is written to (padding at the end of an element, the unused half of the last
// 128-row block): all k-chunks zero, exponent 0
__global__ void i8_zero_rows_kernel(int8_t *sl, int KT, const int *rows, int nrows, int *expo, int nr) {
    const int r = rows[blockIdx.x], kt = blockIdx.y * 128 + threadIdx.x;
    if (kt < KT) {
        int8_t *dst = sl + (((long long)(r >> 7) * (KT + 1) + kt) * I8_S) * 2048 + (r & 127) * 16;
#pragma unroll
        for (int p = 0; p < I8_S; p++) *reinterpret_cast<uint4 *>(dst + p * 2048) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (kt == 0 && r < nr) expo[r] = 0;
}

// =============================================================================================
// host: tables
// =============================================================================================
static double zstar_host(int z) {
    static const int zs[20] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 35, 53};
    static const float ze[20] = {1.0f, 1.688f, 1.279f, 1.912f, 2.421f, 3.136f, 3.834f, 4.453f, 5.100f, 5.758f,
                                 2.507f, 3.308f, 4.066f, 4.285f, 4.886f, 5.482f, 6.116f, 6.764f, 9.028f, 11.612f};
    for (int j = 0; j < 20; j++) if (zs[j] == z) return (double)ze[j];  // atomdb.f90:97-106 (single-precision literals)
    fail(AQML_EINVAL, "izeff: no effective nuclear charge tabulated for Z=%d (atomdb.f90:94-96)", z);
}

struct SlatmTables {
    SlatmDev dev;
    std::vector<int> labels;  // element labels, ascending
    int pldmax = 0;
    int nmb = 0;
    std::vector<int> owner_h;
};

static void check_params(const aqml_slatm_params *p) {
    AQML_REQUIRE(p && p->mbtypes && p->nmb >= 1, "missing many-body types");
    AQML_REQUIRE(p->nx2 >= 2 && p->nx3 >= 2, "grids need at least 2 points");
    AQML_REQUIRE(p->sigma2 > 0 && p->sigma3 > 0 && p->rcut > 0, "sigma/rcut must be positive");
}

// mode_local: layout of element e = every 2-body block containing e + every 3-body block centred on e
// global: 2-body block owned by the smaller element index, 3-body by its centre
// device = false: host-side results only (row pitches, labels) -- a layout-only representation uploads no table
static SlatmTables build_tables(const aqml_slatm_params *p, bool mode_local, Workspace &ws, bool device = true) {
    auto up = [&](const auto &v) -> decltype(ws.upload(v)) { return device ? ws.upload(v) : nullptr; };
    check_params(p);
    SlatmTables t;
    t.nmb = p->nmb;
    std::vector<int> &labels = t.labels;
    for (int s = 0; s < p->nmb; s++)
        for (int q = 0; q < 3; q++) {
            int z = p->mbtypes[3 * s + q];
            if (z >= 0) labels.push_back(z);
            else AQML_REQUIRE(q > 0, "mbtype %d is empty", s);
        }
    std::sort(labels.begin(), labels.end());
    labels.erase(std::unique(labels.begin(), labels.end()), labels.end());
    const int ne = (int)labels.size();
    if (ne > MAXEL) fail(AQML_ELIMIT, "%d distinct elements in mbtypes, limit is %d", ne, MAXEL);
    auto eidx = [&](int z) { return (int)(std::lower_bound(labels.begin(), labels.end(), z) - labels.begin()); };

    SlatmDev &d = t.dev;
    memset(&d, 0, sizeof(d));
    d.n_el = ne; d.nmb = p->nmb; d.nx2 = p->nx2; d.nx3 = p->nx3;
    d.sigma2 = p->sigma2; d.sigma3 = p->sigma3; d.dgrid2 = p->dgrid2; d.dgrid3 = p->dgrid3;
    d.coeff2 = p->coeff2; d.coeff3 = p->coeff3; d.rcut = p->rcut; d.rpower = p->rpower;
    for (int e = 0; e < ne; e++) {
        d.el_label[e] = labels[e];
        d.el_w[e] = p->izeff ? zstar_host(labels[e]) : (double)(labels[e] % 1000);
    }
    std::vector<int> kind(p->nmb), col(p->nmb), e0(p->nmb, -1), owner(p->nmb, -1);
    std::vector<int> bop(MAXEL * MAXEL, -1), bot(MAXEL * MAXEL * MAXEL, -1);
    std::vector<int> boa_seen(MAXEL, 0);
    int D = 0;
    for (int s = 0; s < p->nmb; s++) {
        const int *mb = p->mbtypes + 3 * s;
        col[s] = D;
        if (mb[1] < 0) {
            kind[s] = 1; D += 1; e0[s] = eidx(mb[0]); owner[s] = e0[s];
            if (boa_seen[e0[s]]++) fail(AQML_ELIMIT, "duplicate one-body type [%d]", mb[0]);
        } else if (mb[2] < 0) {
            kind[s] = 2; D += p->nx2;
            int a = eidx(mb[0]), b = eidx(mb[1]);
            if (bop[a * MAXEL + b] >= 0) fail(AQML_ELIMIT, "two-body type [%d,%d] listed twice (either order)", mb[0], mb[1]);
            bop[a * MAXEL + b] = bop[b * MAXEL + a] = s;
            owner[s] = std::min(a, b);
        } else {
            kind[s] = 3; D += p->nx3;
            int a = eidx(mb[0]), c = eidx(mb[1]), b = eidx(mb[2]);
            if (bot[(a * MAXEL + c) * MAXEL + b] >= 0)
                fail(AQML_ELIMIT, "three-body type [%d,%d,%d] listed twice (either orientation)", mb[0], mb[1], mb[2]);
            // kernel enumerates neighbour pairs with e1 <= e3; flip=1 when the listed type has z1 > z3
            bot[(a * MAXEL + c) * MAXEL + b] = 2 * s + (a > b ? 1 : 0);
            bot[(b * MAXEL + c) * MAXEL + a] = 2 * s + (a > b ? 1 : 0);
            owner[s] = c;
        }
    }
    d.D = D;
    std::vector<int> lay((size_t)ne * p->nmb, -1);
    for (int e = 0; e < ne; e++) {
        int w = 0;
        for (int pass = 2; pass <= 3; pass++)
            for (int s = 0; s < p->nmb; s++) {
                if (kind[s] != pass) continue;
                const int *mb = p->mbtypes + 3 * s;
                bool in;
                if (pass == 2) in = mode_local ? (eidx(mb[0]) == e || eidx(mb[1]) == e) : (owner[s] == e);
                else in = (eidx(mb[1]) == e);
                if (in) { lay[(size_t)e * p->nmb + s] = w; w += (pass == 2 ? p->nx2 : p->nx3); }
            }
        d.pdim[e] = w;
        d.pld[e] = (int)round_up(std::max(w, 1), BK);
        t.pldmax = std::max(t.pldmax, d.pld[e]);
    }
    // static three-body block table per centre element: (e1 <= e3) with a listed type
    std::vector<Block3> blk((size_t)ne * MAXBLK);
    for (int ec = 0; ec < ne; ec++) {
        int nb = 0;
        for (int e1 = 0; e1 < ne; e1++)
            for (int e3 = e1; e3 < ne; e3++) {
                int sl = bot[(e1 * MAXEL + ec) * MAXEL + e3];
                if (sl < 0) continue;
                Block3 b;
                b.e1 = e1; b.e3 = e3; b.col = lay[(size_t)ec * p->nmb + (sl >> 1)]; b.flip = sl & 1; b.p0 = 0;
                // c0 = prefactor * (Z1*Z2*Z3) * coeff * dgrid (fslatm.f90:324-332), factors in the listed order
                double w1 = d.el_w[b.flip ? e3 : e1], w3 = d.el_w[b.flip ? e1 : e3];
                b.c0 = (1.0 / 3.0) * (w1 * d.el_w[ec] * w3) * p->coeff3 * p->dgrid3;
                blk[(size_t)ec * MAXBLK + nb++] = b;
            }
        d.nblk[ec] = nb;
    }
    d.blk = up(blk);
    // grids on the host (same libm as the CPU reference): utils.f90:29-50, slatm.py:60-62,118-122
    std::vector<double> xs2(p->nx2), xp2(p->nx2), xs3(p->nx3), cx3(p->nx3);
    {
        double step = (p->rcut - 0.1) / (p->nx2 - 1);
        for (int i = 0; i < p->nx2; i++) { xs2[i] = 0.1 + i * step; xp2[i] = std::pow(xs2[i], p->rpower); }
        const double pi = 4.0 * std::atan(1.0), d2r = pi / 180.0, a0 = -20.0 * d2r, a1 = pi + 20.0 * d2r;
        double st3 = (a1 - a0) / (p->nx3 - 1);
        for (int i = 0; i < p->nx3; i++) { xs3[i] = a0 + i * st3; cx3[i] = std::cos(xs3[i]); }
    }
    d.slot_kind = up(kind); d.slot_col = up(col); d.slot_e0 = up(e0);
    d.bop_slot = up(bop); d.bot_slot = up(bot); d.lay_col = up(lay); d.owner = up(owner);
    d.xs2 = up(xs2); d.xpow2 = up(xp2); d.xs3 = up(xs3); d.cosx3 = up(cx3);
    {
        std::vector<double> inv2(p->nx2);
        for (int i = 0; i < p->nx2; i++) inv2[i] = p->dgrid2 / xp2[i];
        d.inv2 = up(inv2);
        const double pi = 4.0 * std::atan(1.0);
        d.h2 = (p->rcut - 0.1) / (p->nx2 - 1);
        d.h3 = (pi + 40.0 * pi / 180.0) / (p->nx3 - 1);
        d.cst2 = 0.5 / (p->sigma2 * p->sigma2);
        d.cst3 = 0.5 / (p->sigma3 * p->sigma3);
        d.q2 = std::exp(-2.0 * d.h2 * d.h2 * d.cst2);
        d.q3 = std::exp(-2.0 * d.h3 * d.h3 * d.cst3);
    }
    t.owner_h = owner;
    return t;
}

struct MolIndex {
    std::vector<int> off, atom_mol;
    int A = 0, na_max = 0;
};
static MolIndex index_molecules(const int *nas, int nmol) {
    MolIndex mi;
    mi.off.resize(nmol + 1, 0);
    for (int i = 0; i < nmol; i++) {
        AQML_REQUIRE(nas[i] >= 0, "negative atom count");
        mi.off[i + 1] = mi.off[i] + nas[i];
        mi.na_max = std::max(mi.na_max, nas[i]);
    }
    mi.A = mi.off[nmol];
    mi.atom_mol.reserve(mi.A);
    for (int i = 0; i < nmol; i++) mi.atom_mol.insert(mi.atom_mol.end(), nas[i], i);
    return mi;
}

static size_t slatm_smem_bytes(int pldmax, int na_max) {
    return sizeof(double) * ((size_t)pldmax + 10 * (size_t)na_max + 4 * PCH) + sizeof(int) * (5 * (size_t)na_max + 2);
}

static void launch_slatm_general(const SlatmTables &t, const SlatmJob &J, int nblocks, int na_max, cudaStream_t st) {
    size_t smem = slatm_smem_bytes(t.pldmax, na_max);
    if (smem > 200 * 1024)
        fail(AQML_ELIMIT, "molecule with %d atoms / support %d needs %zu B of shared memory (limit 200 KiB)", na_max, t.pldmax, smem);
    AQML_CUDA(cudaFuncSetAttribute(slatm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    slatm_kernel<<<nblocks, SL_THREADS, smem, st>>>(t.dev, J, t.pldmax, na_max);
    AQML_LAUNCH_CHECK();
}

template <int M2, int M3, int GS>
static bool launch_units_t(const SlatmTables &t, const SlatmJob &J, const MolIndex &mi, Workspace &ws) {
    // batches: contiguous centre atoms of one molecule; the neighbour tables of a batch take 11 B per
    // (centre, atom) pair of shared memory -- 35 B with the unit vectors
    std::vector<Batch> batches;
    int nbn_max = 1, ba_max = 1;
    for (int m = 0; m < J.nmol; m++) {
        int na = mi.off[m + 1] - mi.off[m];
        if (na == 0) continue;
        int ba = std::max(1, std::min(std::min(32, na), 45000 / (35 * na)));
        if (!J.local) {  // global SLATM: all centres of the molecule in one batch
            if (na > 36) return false;
            ba = na;
        }
        for (int a0 = 0; a0 < na; a0 += ba) {
            Batch b{m, a0, std::min(ba, na - a0), 0};
            batches.push_back(b);
            nbn_max = std::max(nbn_max, b.cnt * na);
            ba_max = std::max(ba_max, b.cnt);
        }
    }
    if (batches.empty()) return true;
    int nblk_max = 0;
    for (int e = 0; e < t.dev.n_el; e++) nblk_max = std::max(nblk_max, t.dev.nblk[e]);
    const int uc = t.dev.n_el + nblk_max;
    const int na_max = (int)round_up(std::max(mi.na_max, 1), 2);
    nbn_max = (int)round_up(nbn_max, 4);
    ba_max = std::max(ba_max, t.dev.n_el);
    size_t smem = sizeof(double) * (3 * (size_t)na_max + 4 * (size_t)nbn_max + 2 * (size_t)ba_max * uc + 2 * GS * M2 + 2 * GS * M3) +
                  sizeof(int) * (na_max + ba_max) + sizeof(short) * ((size_t)nbn_max + (size_t)ba_max * (MAXEL + 1) + 4) + nbn_max + 16;
    if (smem > 160 * 1024 || mi.na_max > 30000) return false;
    AQML_CUDA(cudaFuncSetAttribute(slatm_units_kernel<M2, M3, GS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    slatm_units_kernel<M2, M3, GS><<<(unsigned)batches.size(), SL_THREADS, smem, ws.stream()>>>(
        t.dev, J, ws.upload(batches), na_max, nbn_max, ba_max, uc);
    AQML_LAUNCH_CHECK();
    return true;
}

// true: the work-unit kernel ran (it also wrote the digit planes J.psl asked for)
static bool launch_slatm(const SlatmTables &t, const SlatmJob &J, const MolIndex &mi, Workspace &ws) {
    cudaStream_t st = ws.stream();
    int nblocks = J.local ? J.A : J.nmol * t.dev.n_el;
    if (nblocks <= 0) return true;
    int na_max = (int)round_up(std::max(mi.na_max, 1), 2);
    const SlatmDev &d = t.dev;
    const int m2 = (d.nx2 + 7) / 8, m3 = (d.nx3 + 7) / 8;
    // the recurrence walks at most M grid steps from an exactly evaluated anchor; keep the anchor's exponent
    // (M h)^2 / (2 sigma^2) far from the double range, otherwise use one exp per grid point
    auto safe = [&](int M2, int M3) {
        return m2 <= M2 && m3 <= M3 && (M2 * d.h2) * (M2 * d.h2) * d.cst2 < 300.0 && (M3 * d.h3) * (M3 * d.h3) * d.cst3 < 300.0;
    };
    static const bool force_legacy = getenv("AQML_SLATM_LEGACY") != nullptr;
    if (!force_legacy) {  // work-unit kernel (aSLATM; SLATM of molecules with <= 36 atoms)
        // 8 threads per unit measured best (4: 10.8e6 atoms/s, 8: 16.8e6, 16: 16.0e6 on the bench workload)
        if (safe(15, 12)) { if (launch_units_t<15, 12, 8>(t, J, mi, ws)) return true; }
        else if (safe(20, 16)) { if (launch_units_t<20, 16, 8>(t, J, mi, ws)) return true; }
        else if (safe(32, 32)) { if (launch_units_t<32, 32, 8>(t, J, mi, ws)) return true; }
    }
    // general kernel: global SLATM of large molecules, grids outside the recurrence's safe range
    launch_slatm_general(t, J, nblocks, na_max, st);
    return false;
}

static void slatm_batch_dev(const double *coords, const int *zs_dev, const int *nas, int nmol, const aqml_slatm_params *p,
                            int local, double *out, long long ldo, cudaStream_t st) {
    AQML_REQUIRE(coords && zs_dev && nas && out && nmol >= 0, "bad argument");
    require_device();
    if (nmol == 0) return;
    Workspace ws(st);
    SlatmTables t = build_tables(p, local != 0, ws);
    AQML_REQUIRE(ldo >= t.dev.D, "output pitch %lld smaller than D=%d", ldo, t.dev.D);
    MolIndex mi = index_molecules(nas, nmol);
    if (mi.A == 0 && local) return;
    SlatmJob J;
    memset(&J, 0, sizeof(J));
    J.coords = coords; J.zs = zs_dev; J.mol_off = ws.upload(mi.off); J.atom_mol = ws.upload(mi.atom_mol);
    J.nmol = nmol; J.A = mi.A; J.local = local ? 1 : 0; J.out = out; J.ldo = ldo;
    launch_slatm(t, J, mi, ws);
}

}  // namespace aqml

// =============================================================================================
// aqml_rep: element-packed aSLATM
// =============================================================================================
struct aqml_rep {
    int nmol = 0, ne = 0;
    cudaStream_t stream = 0;
    std::vector<void *> owned;
    int64_t bytes = 0;
    struct El {
        int label, n, tiles, ld;
        double *x, *norm;
        int *seg;
        unsigned *mask;
        int maskw;
        int8_t *sl = nullptr;  // int8 digit planes for the tcgen05 engine (i8pair.cuh), null = not built
        int *expo = nullptr;   // per packed row exponent
        int nr = 0;            // packed rows (multiple of 64)
    };
    bool i8_ready = false;     // every element has digit planes
    unsigned char *slab = nullptr;  // the allocation the pair kernel's operands live in (norms, molecule index, masks, planes, exponents)
    int64_t slab_bytes = 0;
    unsigned char *xslab = nullptr; // the allocation of the packed FP64 rows; null: built without them (AQML_REP_NO_ROWS)
    int64_t xslab_bytes = 0;
    mutable std::vector<std::pair<cudaStream_t, cudaEvent_t>> uses;  // last consumer launch per foreign stream (aqml_rep_destroy waits)
    std::vector<El> els;
    std::vector<int> nas;
    std::vector<std::vector<int>> cnt;  // [el][mol]
};

namespace aqml {

// AQML_PAIR_ENGINE=dmma keeps the FP64 DMMA engine for the fused local kernel; default: int8 digit planes on tcgen05
static bool i8_engine_enabled() {
    const char *e = getenv("AQML_PAIR_ENGINE");
    return !(e && (!strcmp(e, "dmma") || !strcmp(e, "fp64")));
}

// digit planes + row exponents of a packed operand (i8pair.cuh)
static void i8_fill_planes(const double *x, long long ld, int nr, cudaStream_t st, int8_t *sl, int *expo);
template <class Alloc>
static void i8_make_planes(const double *x, long long ld, int nr, Alloc &mem, cudaStream_t st, int8_t *&sl, int *&expo,
                           size_t *bytes_out = nullptr) {
    const int KT = (int)(ld / 16), rb = (nr + 127) / 128;
    const size_t bytes = (size_t)rb * (KT + 1) * I8_A_CHUNK;
    sl = mem.template alloc<int8_t>(bytes);
    expo = mem.template alloc<int>(nr);
    if (bytes_out) *bytes_out = bytes;
    i8_fill_planes(x, ld, nr, st, sl, expo);
}
static void i8_fill_planes(const double *x, long long ld, int nr, cudaStream_t st, int8_t *sl, int *expo) {
    const int KT = (int)(ld / 16), rb = (nr + 127) / 128;
    // only the trailing all-zero k-chunk of every row block needs clearing: the slice kernel writes everything else
    AQML_CUDA(cudaMemset2DAsync(sl + (size_t)KT * I8_A_CHUNK, (size_t)(KT + 1) * I8_A_CHUNK, 0, I8_A_CHUNK, rb, st));
    i8_rowexp_kernel<<<(nr + 7) / 8, 256, 0, st>>>(x, ld, nr, expo);
    AQML_LAUNCH_CHECK();
    const int split = std::max(1, std::min(KT, 2048 / std::max(rb, 1)));
    const int per = (KT + split - 1) / split;
    i8_slice_kernel<<<dim3(rb, (KT + per - 1) / per), 128, 0, st>>>(x, ld, nr, expo, KT, per, sl);
    AQML_LAUNCH_CHECK();
}

// the B operand of a group as a TMA tensor (I8Params::tmB)
static void i8_encode_b(CUtensorMap *tm, const I8Group &g) {
    typedef CUresult (*encode_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_t encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        AQML_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        AQML_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        return (encode_t)fn;
    }();
    const cuuint64_t blocks = (cuuint64_t)((g.nrB + 127) / 128) * (cuuint64_t)(g.KT + 1);
    const cuuint64_t dims[3] = {256, (cuuint64_t)I8_S, blocks};                 // 8-byte elements: 2 KB = one plane of a 128-row block
    const cuuint64_t strides[2] = {2048, (cuuint64_t)I8_A_CHUNK};               // bytes: plane, (row block, k-chunk)
    const cuuint32_t box[3] = {128, (cuuint32_t)I8_S, 1}, estr[3] = {1, 1, 1};  // 64 rows x 16 B of every plane: 6 KB
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void *)g.slB, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AQML_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
}

// 2^(-j/4096) as Q53 for the integer epilogue (i8exp.cuh), computed in extended precision on the host, once per device
static const unsigned long long *i8_exp2_table() {
    static const unsigned long long *tab[32] = {nullptr};
    int dev = 0;
    AQML_CUDA(cudaGetDevice(&dev));
    AQML_REQUIRE(dev < 32, "device index %d not supported", dev);
    if (!tab[dev]) {
        std::vector<unsigned long long> h(i8x::TAB_N);
        for (int j = 0; j < i8x::TAB_N; j++)
            h[j] = j ? (unsigned long long)llroundl(exp2l(-(long double)j / i8x::TAB_N) * 9007199254740992.0L) : (1ull << 53) - 1;  // Q53; 2^53 would wrap exp2neg's 32-bit top word
        unsigned long long *d = nullptr;
        AQML_CUDA(cudaMalloc((void **)&d, h.size() * 8));
        AQML_CUDA(cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
        tab[dev] = d;
    }
    return tab[dev];
}

static bool launch_i8(std::vector<I8Group> &gs, const PairParams &PP, Workspace &ws, int store = 0) {
    int total = 0;
    for (auto &g : gs) { g.tile_begin = total; total += g.tilesM * g.tilesN; }
    if (total == 0) return true;
    if ((int)gs.size() > I8_MAXGROUPS) return false;  // (cannot happen with <= 16 elements) -> FP64 engine
    I8Params P;
    memset(&P, 0, sizeof(P));
    for (size_t i = 0; i < gs.size(); i++) i8_encode_b(&P.tmB[i], gs[i]);
    P.groups = ws.upload(gs); P.ngroups = (int)gs.size();
    P.isig = PP.isig; P.nsig = PP.nsig; P.zstride = PP.zstride; P.out = PP.out; P.ldo = PP.ldo; P.slab = PP.slab;
    P.stats = pair_stats_buffer();
    P.chain = PP.chain;
    for (int i = 0; i < 16; i++) { P.order[i] = PP.order[i]; P.nsq[i] = PP.nsq[i]; }
    const int smem = I8_STAGES * I8_STAGE_BYTES + (int)sizeof(I8Smem);
    static_assert(I8_STAGES * I8_STAGE_BYTES + sizeof(I8Smem) <= 232448, "i8_pair_kernel: more than 227 KB of shared memory");
    // AQML_I8_PRODUCTS=26: every digit product with i + j <= 6 (10 x lower truncation error, ~15 % more tensor work)
    static const bool full6 = getenv("AQML_I8_PRODUCTS") && atoi(getenv("AQML_I8_PRODUCTS")) >= 26;
    // AQML_I8_EPILOGUE=fp64: the FP64 epilogue (table exp, squaring chain, FP64 sums); default: integers (i8exp.cuh)
    const bool fp64epi = getenv("AQML_I8_EPILOGUE") && !strcmp(getenv("AQML_I8_EPILOGUE"), "fp64");   // read per call (tests switch it)
    auto kern = full6 ? (fp64epi ? i8_pair_kernel<1, 0> : i8_pair_kernel<1, 1>) : (fp64epi ? i8_pair_kernel<0, 0> : i8_pair_kernel<0, 1>);
    P.exp2tab = i8_exp2_table();
    AQML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int dev = 0, sms = 0;
    AQML_CUDA(cudaGetDevice(&dev));
    AQML_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    P.debug = getenv("AQML_I8_DEBUG") ? atoi(getenv("AQML_I8_DEBUG")) : 0;
    P.store = store;
    P.counter = ws.zeros<int>(1);
    P.gm = getenv("AQML_I8_GM") ? std::max(1, atoi(getenv("AQML_I8_GM"))) : 32;  // 32: lowest DRAM traffic (16: +25 %, 64: +35 %, 96: 4 x; the time does not depend on it)
    if (P.debug) AQML_CUDA(cudaMemsetAsync(P.stats + 2, 0, 496, ws.stream()));
    kern<<<std::min(total, sms), I8_THREADS, smem, ws.stream()>>>(P, total);
    AQML_LAUNCH_CHECK();
    if (P.debug) {
        unsigned long long h[62];
        AQML_CUDA(cudaMemcpyAsync(h, P.stats + 2, 496, cudaMemcpyDeviceToHost, ws.stream()));
        AQML_CUDA(cudaStreamSynchronize(ws.stream()));
        const double n = std::min(total, sms), st = (double)std::max<unsigned long long>(h[4], 1);
        fprintf(stderr, "[i8] %d tiles, %.0f stages; MMA thread per CTA: total %.0f clk = %.0f clk/stage; waiting for metadata %.1f%%, TMEM %.1f%%, operands %.1f%%\n",
                total, st, h[0] / n, (double)h[0] / st, 100.0 * h[1] / h[0], 100.0 * h[2] / h[0], 100.0 * h[3] / h[0]);
        fprintf(stderr, "[i8] epilogue warp per tile: waiting for accumulators %.0f clk, TMEM drain %.0f clk, exp + sums %.0f clk\n", (double)h[5] / total,
                (double)h[6] / total, (double)h[7] / total);
        fprintf(stderr, "[i8] exp + sums per tile, epilogue warps 3..18 (k clk):");
        for (int w = 0; w < I8_EPI_WARPS; w++) fprintf(stderr, " %.1f", h[14 + w] / 1e3 / total);
        fprintf(stderr, "\n[i8] waiting for accumulators per tile (k clk):");
        for (int w = 0; w < I8_EPI_WARPS; w++) fprintf(stderr, " %.1f", h[30 + w] / 1e3 / total);
        fprintf(stderr, "\n");
    }
    return true;
}

// fused path: the representations carry their digit planes
static bool run_pairs_i8(const aqml_rep *r1, const aqml_rep *r2, int zmax, int row0, int nrows, bool sym, const PairParams &PP, Workspace &ws) {
    std::vector<I8Group> gs;
    for (const auto &a : r1->els)
        for (const auto &b : r2->els) {
            if (a.label != b.label || a.n == 0 || b.n == 0) continue;
            AQML_REQUIRE(a.ld == b.ld, "representations were built with different many-body types");
            const int z = a.label % 1000;
            AQML_REQUIRE(z >= 1 && z <= zmax, "element %d outside 1..zmax=%d", z, zmax);
            I8Group g;
            memset(&g, 0, sizeof(g));
            g.slA = a.sl; g.slB = b.sl; g.expA = a.expo; g.expB = b.expo; g.normA = a.norm; g.normB = b.norm;
            g.segA = a.seg; g.segB = b.seg; g.nrA = a.nr; g.nrB = b.nr;
            g.maskA = a.mask; g.maskB = b.mask; g.maskw = a.maskw;
            g.KT = a.ld / 16;
            g.tilesM = (a.nr + I8_BM - 1) / I8_BM; g.tilesN = (b.nr + I8_BN - 1) / I8_BN;
            g.zcol = z - 1; g.segA_off = row0; g.segA_n = nrows; g.sym = sym ? 1 : 0;
            gs.push_back(g);
        }
    return launch_i8(gs, PP, ws);
}

}  // namespace aqml

// packed FP64 operands from the dense API (kernels.cu): temporary digit planes
bool aqml::i8_local_pairs(const std::vector<I8Pair> &pairs, const PairParams &PP, Workspace &ws) {
    if (!i8_engine_enabled() || PP.nsig > I8_MAXSIG) return false;
    for (const auto &pr : pairs)
        if (pr.a.ld > I8_MAX_LD) return false;
    cudaStream_t st = ws.stream();
    std::vector<I8Group> gs;
    for (const auto &pr : pairs) {
        if (pr.a.nr == 0 || pr.b.nr == 0) continue;
        I8Group g;
        memset(&g, 0, sizeof(g));
        int8_t *sa, *sb;
        int *ea, *eb;
        i8_make_planes(pr.a.x, pr.a.ld, pr.a.nr, ws, st, sa, ea);
        if (pr.b.x == pr.a.x && pr.b.nr == pr.a.nr) { sb = sa; eb = ea; }
        else i8_make_planes(pr.b.x, pr.b.ld, pr.b.nr, ws, st, sb, eb);
        g.slA = sa; g.slB = sb; g.expA = ea; g.expB = eb; g.normA = pr.a.norm; g.normB = pr.b.norm;
        g.segA = pr.a.seg; g.segB = pr.b.seg; g.nrA = pr.a.nr; g.nrB = pr.b.nr;
        g.maskA = pr.a.mask; g.maskB = pr.b.mask; g.maskw = pr.a.maskw;
        g.KT = (int)(pr.a.ld / 16);
        g.tilesM = (pr.a.nr + I8_BM - 1) / I8_BM; g.tilesN = (pr.b.nr + I8_BN - 1) / I8_BN;
        g.zcol = pr.zcol; g.segA_off = 0; g.segA_n = 0x7fffffff; g.sym = 0;
        gs.push_back(g);
    }
    return launch_i8(gs, PP, ws);
}

bool aqml::i8_global_pair(const I8Side &a, int na, const I8Side &b, int nb, const PairParams &PP, Workspace &ws) {
    if (!i8_engine_enabled() || PP.nsig > I8_MAXSIG || na <= 0 || nb <= 0 || a.ld > I8_MAX_LD) return false;
    cudaStream_t st = ws.stream();
    I8Group g;
    memset(&g, 0, sizeof(g));
    int8_t *sa, *sb;
    int *ea, *eb;
    i8_make_planes(a.x, a.ld, a.nr, ws, st, sa, ea);
    const bool same = (b.x == a.x && b.nr == a.nr);
    if (same) { sb = sa; eb = ea; }
    else i8_make_planes(b.x, b.ld, b.nr, ws, st, sb, eb);
    auto ident = [&](int nr, int n) {
        std::vector<int> v(nr, -1);
        for (int i = 0; i < n && i < nr; i++) v[i] = i;
        return ws.upload(v);
    };
    g.slA = sa; g.slB = sb; g.expA = ea; g.expB = eb; g.normA = a.norm; g.normB = b.norm;
    g.segA = ident(a.nr, na); g.segB = ident(b.nr, nb); g.nrA = a.nr; g.nrB = b.nr;
    g.maskA = a.mask; g.maskB = b.mask; g.maskw = a.maskw;
    g.KT = (int)(a.ld / 16);
    g.tilesM = (a.nr + I8_BM - 1) / I8_BM; g.tilesN = (b.nr + I8_BN - 1) / I8_BN;
    g.zcol = 0; g.segA_off = 0; g.segA_n = 0x7fffffff; g.sym = 0;
    std::vector<I8Group> gs{g};
    return launch_i8(gs, PP, ws, 1);
}

namespace aqml {

// a kernel on `st` reads the rep: remember it if `st` is not the stream the rep's memory is freed on
static void mark_use(const aqml_rep *r, cudaStream_t st) {
    if (!r || st == r->stream) return;
    for (auto &u : r->uses)
        if (u.first == st) { cudaEventRecord(u.second, st); return; }
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return;
    cudaEventRecord(ev, st);
    r->uses.push_back({st, ev});
}

static bool use_i8(const aqml_rep *r1, const aqml_rep *r2, int nsigmas) {
    for (const auto &e : r1->els)
        if (e.ld > I8_MAX_LD) return false;
    return i8_engine_enabled() && r1->i8_ready && r2->i8_ready && nsigmas <= I8_MAXSIG;
}

// flags bit 0: layout only -- allocate the storage exactly as a computing call would, launch nothing: the contents
// arrive from the rank that computed them (aqml_rep_slab + a broadcast; the layout depends only on zs / nas / p)
static void rep_create(const double *coords, int on_dev, const int *zs, const int *nas, int nmol,
                       const aqml_slatm_params *p, int flags, cudaStream_t st, aqml_rep **out) {
    const bool compute = !(flags & 1), with_x = !(flags & 2);
    AQML_REQUIRE((coords || !compute) && zs && nas && out && nmol >= 1, "bad argument");
    AQML_REQUIRE(with_x || !compute, "AQML_REP_NO_ROWS needs AQML_REP_LAYOUT_ONLY");
    require_device();
    Workspace ws(st);     // temporaries
    Workspace keep(st);   // becomes the rep's storage: |row|^2, molecule index, masks, digit planes, exponents -- what the pair kernel reads
    Workspace keepx(st);  // ... and the packed FP64 rows (width pass, FP64 engine, source of the digit planes)
    SlatmTables t = build_tables(p, true, ws, compute);
    MolIndex mi = index_molecules(nas, nmol);
    AQML_REQUIRE(mi.A >= 1, "no atoms");
    const int ne = t.dev.n_el;
    auto eidx = [&](int z) {
        auto it = std::lower_bound(t.labels.begin(), t.labels.end(), z);
        return (it != t.labels.end() && *it == z) ? (int)(it - t.labels.begin()) : -1;
    };
    std::unique_ptr<aqml_rep> rep(new aqml_rep());
    rep->nmol = nmol; rep->stream = st; rep->nas.assign(nas, nas + nmol);
    std::vector<std::vector<int>> rows(ne), mols(ne);
    rep->cnt.assign(ne, std::vector<int>(nmol, 0));
    // molecules in composition-class order: blocks of packed rows share their zero many-body blocks
    for (int m : class_sorted_molecules(zs, nas, nmol))
        for (int a = mi.off[m]; a < mi.off[m + 1]; a++) {
            int e = eidx(zs[a]);
            if (e < 0) continue;  // element without any many-body type: contributes nothing
            rows[e].push_back(a);
            mols[e].push_back(m);
            rep->cnt[e][m]++;
        }
    const bool planes = i8_engine_enabled();
    // pass 1: packed row maps and the size of the storage (one slab, so that a rep can travel as one buffer)
    std::vector<std::vector<int>> rms(ne), sgs(ne);
    size_t total = 0, totalx = 0;
    auto add = [&](size_t &tot, size_t n, size_t elem) { tot = Workspace::slab_aligned(tot) + (n ? n : 1) * elem; };  // = Workspace::alloc
    for (int e = 0; e < ne; e++) {
        build_tiles(rows[e], mols[e], rms[e], sgs[e]);
        const size_t nr = rms[e].size(), ld = (size_t)t.dev.pld[e];
        const int mw = mask_words((long long)ld);
        add(totalx, nr * ld, 8);
        add(total, nr, 8); add(total, nr, 4);
        if (mw <= 32 && nr) add(total, nr / 64 * (size_t)mw, 4);
        if (planes && !rows[e].empty()) { add(total, (size_t)((nr + 127) / 128) * (ld / 16 + 1) * I8_A_CHUNK, 1); add(total, nr, 4); }
    }
    keep.reserve(total);
    if (with_x) keepx.reserve(totalx);
    // Digit planes: by default two small split kernels after the aSLATM kernel (a second pass over the packed operand).
    // AQML_FUSE_PLANES=1 lets the aSLATM kernel write them itself (slatm_units.cuh, phase 4).  Measured (10 k molecules,
    // profiles/r02_slatm_units_kernel_ncu.txt): fused 20.5 ms, separate 18.1 ms -- with ~590 CTAs x 290 KB of rows in flight
    // (172 MB > the 126 MB L2) the re-read of phase 4 misses L2 half of the time (4.2 GB of DRAM reads), so the fusion
    // saves neither traffic nor time until the work units are handed out atom-major (rows would complete one by one).
    const bool fuse = getenv("AQML_FUSE_PLANES") && atoi(getenv("AQML_FUSE_PLANES")) != 0;
    // pass 2: carve (same order)
    std::vector<int> atom_prow(mi.A, -1);
    SlatmJob J;
    memset(&J, 0, sizeof(J));
    for (int e = 0; e < ne; e++) {
        const std::vector<int> &rm = rms[e], &sg = sgs[e];
        aqml_rep::El el;
        el.label = t.labels[e]; el.n = (int)rows[e].size(); el.tiles = (int)rm.size() / BM; el.ld = t.dev.pld[e];
        size_t nr = rm.size();
        el.x = with_x ? keepx.alloc<double>(nr * (size_t)el.ld) : nullptr;
        el.norm = keep.alloc<double>(nr);
        el.seg = keep.alloc<int>(nr);
        if (compute && nr) AQML_CUDA(cudaMemcpyAsync(el.seg, sg.data(), nr * sizeof(int), cudaMemcpyHostToDevice, st));
        el.maskw = mask_words(el.ld);
        el.mask = nullptr;
        if (el.maskw <= 32 && nr) {
            el.mask = keep.alloc<unsigned>(nr / 64 * (size_t)el.maskw);
            if (compute) AQML_CUDA(cudaMemsetAsync(el.mask, 0, nr / 64 * (size_t)el.maskw * sizeof(unsigned), st));
        }
        rep->bytes += (int64_t)(nr * (size_t)el.ld * 8 + nr * 12);
        std::vector<int> pad;
        for (size_t r = 0; r < nr; r++) {
            if (rm[r] >= 0) atom_prow[rm[r]] = (int)r;
            else pad.push_back((int)r);
        }
        if (compute && !pad.empty()) {
            zero_rows_kernel<<<(unsigned)pad.size(), 128, 0, st>>>(el.x, el.ld, ws.upload(pad), (int)pad.size(), el.norm);
            AQML_LAUNCH_CHECK();
        }
        el.nr = (int)nr;
        if (planes && el.n) {  // digit planes + row exponents (6 bytes per element next to the 8 of the FP64 copy)
            const size_t bytes = (size_t)((nr + 127) / 128) * (el.ld / 16 + 1) * I8_A_CHUNK;
            el.sl = keep.alloc<int8_t>(bytes);
            el.expo = keep.alloc<int>(nr);
            rep->bytes += (int64_t)bytes + 4 * el.nr;
        }
        J.px[e] = el.x; J.pnorm[e] = el.norm; J.pmask[e] = el.mask; J.maskw[e] = el.maskw;
        J.psl[e] = fuse ? el.sl : nullptr; J.pexpo[e] = fuse ? el.expo : nullptr;
        if (compute && el.sl && fuse) {
            // planes nobody else writes: the all-zero k-chunk KT of every row block, and the padding rows at the element's end
            const int KT = el.ld / 16, rb = (int)((nr + 127) / 128);
            AQML_CUDA(cudaMemset2DAsync(el.sl + (size_t)KT * I8_A_CHUNK, (size_t)(KT + 1) * I8_A_CHUNK, 0, I8_A_CHUNK, rb, st));
            std::vector<int> empty_rows = pad;
            for (int r = (int)nr; r < rb * 128; r++) empty_rows.push_back(r);
            if (!empty_rows.empty()) {
                i8_zero_rows_kernel<<<dim3((unsigned)empty_rows.size(), (KT + 127) / 128), 128, 0, st>>>(el.sl, KT, ws.upload(empty_rows),
                                                                                                   (int)empty_rows.size(), el.expo, (int)nr);
                AQML_LAUNCH_CHECK();
            }
        }
        rep->els.push_back(el);
    }
    rep->ne = ne;
    if (compute) {
        const double *dcoords = on_dev ? coords : ws.upload(coords, (size_t)mi.A * 3);
        J.coords = dcoords; J.zs = ws.upload(zs, mi.A); J.mol_off = ws.upload(mi.off); J.atom_mol = ws.upload(mi.atom_mol);
        J.nmol = nmol; J.A = mi.A; J.local = 1; J.out = nullptr; J.ldo = 0; J.atom_prow = ws.upload(atom_prow);
        const bool fused_planes = launch_slatm(t, J, mi, ws);
        if (planes && !(fused_planes && fuse))  // the general SLATM kernel does not write digit planes: split the finished rows
            for (auto &el : rep->els)
                if (el.n) i8_fill_planes(el.x, el.ld, el.nr, st, el.sl, el.expo);
    }
    rep->i8_ready = planes;
    rep->slab = keep.slab(); rep->slab_bytes = (int64_t)keep.slab_bytes();
    rep->xslab = with_x ? keepx.slab() : nullptr; rep->xslab_bytes = with_x ? (int64_t)keepx.slab_bytes() : 0;
    rep->owned = keep.pointers();
    for (void *q : keepx.pointers()) rep->owned.push_back(q);
    keep.release_all();
    keepx.release_all();
    *out = rep.release();
}

static void rep_groups(const aqml_rep *r1, const aqml_rep *r2, int zmax, std::vector<PairGroup> &groups,
                       std::vector<int> &labels) {
    for (const auto &a : r1->els)
        for (const auto &b : r2->els) {
            if (a.label != b.label || a.n == 0 || b.n == 0) continue;
            AQML_REQUIRE(a.ld == b.ld, "representations were built with different many-body types");
            int z = a.label % 1000;
            AQML_REQUIRE(z >= 1 && z <= zmax, "element %d outside 1..zmax=%d", z, zmax);
            PairGroup g;
            memset(&g, 0, sizeof(g));
            g.A = a.x; g.B = b.x; g.normA = a.norm; g.normB = b.norm; g.segA = a.seg; g.segB = b.seg;  // null x: checked by the caller if used
            g.ld = a.ld; g.tilesM = a.tiles; g.tilesN = b.tiles * (BM / BN); g.zcol = z - 1; g.segA_off = 0; g.segA_n = 0x7fffffff;
            g.maskA = a.mask; g.maskB = b.mask; g.maskw = a.maskw;
            groups.push_back(g);
            labels.push_back(a.label);
        }
}

}  // namespace aqml

using namespace aqml;

extern "C" {

int aqml_slatm_dim(const aqml_slatm_params *p) {
    if (!p || !p->mbtypes) return AQML_EINVAL;
    int D = 0;
    for (int s = 0; s < p->nmb; s++) {
        const int *mb = p->mbtypes + 3 * s;
        D += (mb[1] < 0) ? 1 : (mb[2] < 0 ? p->nx2 : p->nx3);
    }
    return D;
}

int aqml_generate_slatm_batch_dev(const double *coords, const int *zs_dev, const int *nas, int nmol,
                                  const aqml_slatm_params *p, int local, double *out, int64_t ldo, void *stream) {
    return guarded([&] { slatm_batch_dev(coords, zs_dev, nas, nmol, p, local, out, ldo, (cudaStream_t)stream); });
}

int aqml_generate_slatm_batch(const double *coords, const int *zs, const int *nas, int nmol, const aqml_slatm_params *p,
                              int local, double *out) {
    return guarded([&] {
        AQML_REQUIRE(coords && zs && nas && out && nmol >= 0, "bad argument");
        require_device();
        if (nmol == 0) return;
        check_params(p);
        int D = aqml_slatm_dim(p);
        MolIndex mi = index_molecules(nas, nmol);
        size_t rows = local ? (size_t)mi.A : (size_t)nmol;
        if (rows == 0) return;
        cudaStream_t st = 0;
        Workspace ws(st);
        double *dc = ws.upload(coords, (size_t)mi.A * 3);
        int *dz = ws.upload(zs, (size_t)mi.A);
        double *dout = ws.alloc<double>(rows * (size_t)D);
        slatm_batch_dev(dc, dz, nas, nmol, p, local, dout, D, st);
        AQML_CUDA(cudaMemcpyAsync(out, dout, sizeof(double) * rows * (size_t)D, cudaMemcpyDeviceToHost, st));
        AQML_CUDA(cudaStreamSynchronize(st));
    });
}

static int single_type(const double *coords, const int *zs, int na, int ia, int izeff, const int *mb, double rcut, int nx,
                       double dgrid, double sigma, double coeff, double rpower, double *ys) {
    return guarded([&] {
        AQML_REQUIRE(coords && zs && ys && na >= 1 && nx >= 2, "bad argument");
        AQML_REQUIRE(ia >= -1 && ia < na, "centre atom %d outside the molecule", ia);
        aqml_slatm_params p;
        memset(&p, 0, sizeof(p));
        p.nmb = 1; p.mbtypes = mb; p.izeff = izeff; p.nx2 = nx; p.nx3 = nx;
        p.sigma2 = p.sigma3 = sigma; p.dgrid2 = p.dgrid3 = dgrid; p.coeff2 = p.coeff3 = coeff;
        p.rcut = rcut; p.rpower = rpower;
        size_t rows = (ia >= 0) ? (size_t)na : 1;
        std::vector<double> tmp(rows * (size_t)nx);
        int rc = aqml_generate_slatm_batch(coords, zs, &na, 1, &p, ia >= 0, tmp.data());
        if (rc != AQML_OK) fail(rc, "%s", aqml_last_error());
        memcpy(ys, tmp.data() + (ia >= 0 ? (size_t)ia * nx : 0), sizeof(double) * nx);
    });
}

int aqml_calc_sbop(const double *coords, const int *zs, int na, int ia, int izeff, int z1, int z2, double rcut, int nx,
                   double dgrid, double sigma, double coeff, double rpower, double *ys) {
    int mb[3] = {z1, z2, -1};
    return single_type(coords, zs, na, ia, izeff, mb, rcut, nx, dgrid, sigma, coeff, rpower, ys);
}

int aqml_calc_sbot(const double *coords, const int *zs, int na, int ia, int izeff, int z1, int z2, int z3, double rcut,
                   int nx, double dgrid, double sigma, double coeff, double *ys) {
    int mb[3] = {z1, z2, z3};
    return single_type(coords, zs, na, ia, izeff, mb, rcut, nx, dgrid, sigma, coeff, 6.0, ys);
}

int aqml_measure_i8_peak(double *ops_per_s, void *stream) {
    return guarded([&] {
        AQML_REQUIRE(ops_per_s, "bad argument");
        require_device();
        cudaStream_t st = (cudaStream_t)stream;
        int dev = 0, sms = 0;
        AQML_CUDA(cudaGetDevice(&dev));
        AQML_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        cudaEvent_t e0, e1;
        AQML_CUDA(cudaEventCreate(&e0));
        AQML_CUDA(cudaEventCreate(&e1));
        const int rounds = 20000;  // x 2 MMAs x 128 clk = 2.6 ms at 1.965 GHz
        double best = 0.0;
        for (int rep = 0; rep < 4; rep++) {  // first pass warms up
            AQML_CUDA(cudaEventRecord(e0, st));
            i8_peak_kernel<<<sms, 128, 0, st>>>(rounds);
            AQML_LAUNCH_CHECK();
            AQML_CUDA(cudaEventRecord(e1, st));
            AQML_CUDA(cudaEventSynchronize(e1));
            float ms = 0.f;
            AQML_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0) best = std::max(best, 2.0 * 128 * 256 * 32 * 2.0 * rounds * sms / (ms * 1e-3));
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *ops_per_s = best;
    });
}

int aqml_rep_create(const double *coords, int coords_on_device, const int *zs, const int *nas, int nmol,
                    const aqml_slatm_params *p, void *stream, aqml_rep **out) {
    return guarded([&] { rep_create(coords, coords_on_device, zs, nas, nmol, p, 0, (cudaStream_t)stream, out); });
}

int aqml_rep_create_ex(const double *coords, int coords_on_device, const int *zs, const int *nas, int nmol,
                       const aqml_slatm_params *p, int flags, void *stream, aqml_rep **out) {
    return guarded([&] { rep_create(coords, coords_on_device, zs, nas, nmol, p, flags, (cudaStream_t)stream, out); });
}

int aqml_rep_slab(const aqml_rep *r, int which, void **ptr, int64_t *bytes) {
    return guarded([&] {
        AQML_REQUIRE(r && ptr && bytes && (which == 0 || which == 1), "bad argument");
        *ptr = which ? r->xslab : r->slab;
        *bytes = which ? r->xslab_bytes : r->slab_bytes;
    });
}

int aqml_rep_elements(const aqml_rep *r) { return r ? (int)r->els.size() : 0; }

int aqml_rep_element(const aqml_rep *r, int e, int *label, int *n, int *ld, const double **x) {
    return guarded([&] {
        AQML_REQUIRE(r && e >= 0 && e < (int)r->els.size(), "bad argument");
        const auto &el = r->els[e];
        if (label) *label = el.label;
        if (n) *n = el.n;
        if (ld) *ld = el.ld;
        if (x) *x = el.x;
    });
}

void aqml_rep_destroy(aqml_rep *r) {
    if (!r) return;
    // kernels that read the rep on another stream than the one it was created on must finish before the memory goes
    // back to the pool
    for (auto &u : r->uses) {
        cudaStreamWaitEvent(r->stream, u.second, 0);
        cudaEventDestroy(u.second);
    }
    for (void *p : r->owned) cudaFreeAsync(p, r->stream);
    delete r;
}

int64_t aqml_rep_bytes(const aqml_rep *r) { return r ? r->bytes : 0; }

int aqml_rep_local_gaussian(const aqml_rep *r1, const aqml_rep *r2, const double *sigmas, int nsigmas, int zmax, int row0,
                            int nrows, double *K, void *stream) {
    return guarded([&] {
        AQML_REQUIRE(r1 && r2 && sigmas && K && nsigmas >= 1 && zmax >= 1, "bad argument");
        AQML_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= r1->nmol, "row block outside r1");
        require_device();
        cudaStream_t st = (cudaStream_t)stream;
        for (int i = 0; i < nsigmas * zmax; i++) AQML_REQUIRE(sigmas[i] != 0.0, "sigma is zero");
        AQML_CUDA(cudaMemsetAsync(K, 0, sizeof(double) * (size_t)nsigmas * nrows * r2->nmol, st));
        if (nrows == 0) return;
        Workspace ws(st);
        std::vector<PairGroup> groups;
        std::vector<int> labels;
        rep_groups(r1, r2, zmax, groups, labels);
        for (auto &g : groups) { g.segA_off = row0; g.segA_n = nrows; }
        std::vector<double> isig((size_t)nsigmas * zmax);
        for (size_t i = 0; i < isig.size(); i++) isig[i] = -0.5 / (sigmas[i] * sigmas[i]);
        // element-mismatched pairs carry d2 = 1e9 in the reference (fkernels.f90:73): exactly 0 after exp
        // unless sigma is enormous; refuse instead of silently dropping them
        for (const auto &a : r1->els)
            for (int k = 0; k < nsigmas; k++) {
                int z = a.label % 1000;
                if (a.n && z >= 1 && z <= zmax)
                    AQML_REQUIRE(std::exp((double)1.e9f * isig[(size_t)k * zmax + z - 1]) == 0.0,
                                 "sigma %g for Z=%d is so wide that element-mismatched pairs do not vanish; use aqml_calc_lk_gaussian",
                                 sigmas[(size_t)k * zmax + z - 1], z);
            }
        PairParams P{};
        P.isig = ws.upload(isig); P.nsig = nsigmas; P.zstride = zmax;
        plan_sigma_chain(P, isig);
        P.out = K; P.ldo = r2->nmol; P.slab = (long long)nrows * r2->nmol;
        if (!(use_i8(r1, r2, nsigmas) && run_pairs_i8(r1, r2, zmax, row0, nrows, false, P, ws))) {
            AQML_REQUIRE(r1->xslab && r2->xslab, "the FP64 engine needs the FP64 rows (representation built with AQML_REP_NO_ROWS)");
            run_pairs_public(MODE_DOT, EPI_LOCAL, groups, P, ws);
        }
        mark_use(r1, st);
        mark_use(r2, st);
    });
}

int aqml_rep_local_gaussian_sym(const aqml_rep *r, const double *sigmas, int nsigmas, int zmax, double *K, void *stream) {
    return guarded([&] {
        AQML_REQUIRE(r && sigmas && K && nsigmas >= 1 && zmax >= 1, "bad argument");
        require_device();
        cudaStream_t st = (cudaStream_t)stream;
        const int n = r->nmol;
        for (int i = 0; i < nsigmas * zmax; i++) AQML_REQUIRE(sigmas[i] != 0.0, "sigma is zero");
        AQML_CUDA(cudaMemsetAsync(K, 0, sizeof(double) * (size_t)nsigmas * n * n, st));
        Workspace ws(st);
        std::vector<PairGroup> groups;
        std::vector<int> labels;
        rep_groups(r, r, zmax, groups, labels);
        for (auto &g : groups) g.sym = 1;
        std::vector<double> isig((size_t)nsigmas * zmax);
        for (size_t i = 0; i < isig.size(); i++) isig[i] = -0.5 / (sigmas[i] * sigmas[i]);
        // i == j terms: exp(0) = 1 for every atom whose element has a group (exactly what the reference adds)
        std::vector<double> diag(n, 0.0);
        for (size_t e = 0; e < r->els.size(); e++) {
            const auto &el = r->els[e];
            int z = el.label % 1000;
            if (!el.n) continue;
            for (int k = 0; k < nsigmas; k++)
                AQML_REQUIRE(std::exp((double)1.e9f * isig[(size_t)k * zmax + z - 1]) == 0.0,
                             "sigma %g for Z=%d is so wide that element-mismatched pairs do not vanish; use aqml_calc_lk_gaussian",
                             sigmas[(size_t)k * zmax + z - 1], z);
            for (int m = 0; m < n; m++) diag[m] += r->cnt[e][m];
        }
        PairParams P{};
        P.isig = ws.upload(isig); P.nsig = nsigmas; P.zstride = zmax;
        plan_sigma_chain(P, isig);
        P.out = K; P.ldo = n; P.slab = (long long)n * n;
        if (!(use_i8(r, r, nsigmas) && run_pairs_i8(r, r, zmax, 0, n, true, P, ws))) {
            AQML_REQUIRE(r->xslab, "the FP64 engine needs the FP64 rows (representation built with AQML_REP_NO_ROWS)");
            run_pairs_public(MODE_DOT, EPI_LOCAL, groups, P, ws);
        }
        dim3 grid((n + 31) / 32, (n + 31) / 32, nsigmas);
        symmetrize_kernel<<<grid, 256, 0, st>>>(K, n, nsigmas, ws.upload(diag));
        AQML_LAUNCH_CHECK();
        mark_use(r, st);
    });
}

int aqml_rep_kernel_width_multi(const aqml_rep *const *reps1, int n1, const aqml_rep *const *reps2, int n2, int zmax, double *dso,
                                void *stream) {
    return guarded([&] {
        AQML_REQUIRE(reps1 && reps2 && n1 >= 1 && n2 >= 1 && dso && zmax >= 1, "bad argument");
        require_device();
        cudaStream_t st = (cudaStream_t)stream;
        Workspace ws(st);
        bool same = (n1 == n2);
        for (int i = 0; same && i < n1; i++) same = (reps1[i] == reps2[i]);
        // element label -> row segments of every representation that holds atoms of it
        std::map<int, std::pair<std::vector<RowSeg>, std::vector<RowSeg>>> segs;
        std::map<int, long long> lds;
        auto collect = [&](const aqml_rep *const *reps, int n, bool second) {
            for (int i = 0; i < n; i++) {
                AQML_REQUIRE(reps[i], "null representation");
                for (const auto &e : reps[i]->els) {
                    if (e.n == 0) continue;
                    int z = e.label % 1000;
                    AQML_REQUIRE(z >= 1 && z <= zmax, "element %d outside 1..zmax=%d", z, zmax);
                    auto it = lds.find(e.label);
                    if (it == lds.end()) lds[e.label] = e.ld;
                    else AQML_REQUIRE(it->second == e.ld, "representations were built with different many-body types");
                    AQML_REQUIRE(e.x, "this representation was built without its FP64 rows (AQML_REP_NO_ROWS)");
                    (second ? segs[e.label].second : segs[e.label].first).push_back(RowSeg{e.x, e.n});
                }
            }
        };
        collect(reps1, n1, false);
        collect(reps2, n2, true);
        for (int z = 0; z < zmax; z++) dso[z] = 1.e-9;
        for (auto &kv : segs) {
            if (kv.second.first.empty() || kv.second.second.empty()) continue;
            dso[kv.first % 1000 - 1] = std::sqrt(pruned_max_segs_public(2, kv.second.first, kv.second.second, lds[kv.first], same, ws));
        }
        for (int i = 0; i < n1; i++) mark_use(reps1[i], st);
        for (int i = 0; i < n2; i++) mark_use(reps2[i], st);
    });
}

int aqml_rep_kernel_width(const aqml_rep *r1, const aqml_rep *r2, int zmax, double *dso, void *stream) {
    return aqml_rep_kernel_width_multi(&r1, 1, &r2, 1, zmax, dso, stream);
}

}  // extern "C"
