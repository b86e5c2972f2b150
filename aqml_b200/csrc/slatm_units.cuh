// aSLATM, "work unit" formulation (local representation; the default path).
//
// One CTA per batch of centre atoms of one molecule (a QM9-sized molecule is one batch).
//   phase 1  one warp per centre: distances / cut-off flags and the neighbour list sorted by element
//            (ballot compaction, list order preserved) -> shared memory.
//   phase 2  work units (centre atom, many-body block) are handed out dynamically to 8-thread groups.
//            A unit owns one output block (118 or 96 grid points): each thread keeps M consecutive grid
//            points in registers, walks the unit's neighbours (2-body) or neighbour pairs (3-body;
//            geometry of 8 pairs at a time is computed one pair per thread and shared by shuffles), and
//            evaluates the Gaussians by the two-term recurrence  E(x+h) = E(x) R(x), R(x+h) = R(x) q
//            from an exactly evaluated first point of its chunk: 1 exp per (entry, thread) instead of M
//            (R at the grid start is one more exp per entry, shared).  A chunk that starts > 38 sigma
//            left of a peak starts from an underflowed E: terms below 1e-190 of the block are lost.
//            The finished block goes straight to global memory (dense row or element-packed operand).
//   phase 3  |row|^2 of the packed rows, summed over the row's blocks in fixed order.
// No spectrum / pair table / partial-sum slabs in shared memory (high occupancy), two block barriers per
// batch, every output element has one owner and a fixed summation order (deterministic).
#pragma once

namespace aqml {

struct Batch {
    int mol, a0, cnt, pad;
};

__device__ __forceinline__ double gshfl(double v, int src, unsigned gmask, int gbase) {
    return __shfl_sync(gmask, v, gbase + src);
}

// GS = threads per work unit (4 or 8); M2/M3 = grid points per thread (GS*M >= nx)
template <int M2, int M3, int GS>
__global__ void __launch_bounds__(SL_THREADS, (GS == 16) ? 5 : (GS == 8) ? 4 : 3)
slatm_units_kernel(const SlatmDev T, const SlatmJob J, const Batch *batches, int na_max, int nbn_max, int ba_max, int uc) {
    extern __shared__ __align__(16) double sm[];
    double *mx = sm, *my = mx + na_max, *mz = my + na_max;
    double *dd = mz + na_max;                          // [ba][na] distance centre -> atom q
    double *ux = dd + nbn_max, *uy = ux + nbn_max, *uz = uy + nbn_max;  // [ba][na] unit vector centre -> q
    double *s_norm = uz + nbn_max;                     // [ba][uc] sum of squares of each written block
    double *sx2 = s_norm + ba_max * uc, *sinv2 = sx2 + GS * M2, *sx3 = sinv2 + GS * M2, *scos3 = sx3 + GS * M3;
    int *mel = reinterpret_cast<int *>(scos3 + GS * M3);
    short *nbidx = reinterpret_cast<short *>(mel + na_max);   // [ba][na] neighbours sorted by element
    short *sstart = nbidx + nbn_max;                           // [ba][MAXEL+1]
    unsigned char *fls = reinterpret_cast<unsigned char *>(sstart + ba_max * (MAXEL + 1));  // [ba][na]
    __shared__ int s_next;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double eps = 2.220446049250313e-16;
    const double rcut = T.rcut, rcut2 = __dmul_rn(rcut, rcut);
    const Batch bt = batches[blockIdx.x];
    const int o = J.mol_off[bt.mol], na = J.mol_off[bt.mol + 1] - o;

    for (int i = tid; i < GS * M2; i += SL_THREADS) {
        int q = min(i, T.nx2 - 1);
        sx2[i] = T.xs2[q];
        sinv2[i] = T.inv2[q];
    }
    for (int i = tid; i < GS * M3; i += SL_THREADS) {
        int q = min(i, T.nx3 - 1);
        sx3[i] = T.xs3[q];
        scos3[i] = T.cosx3[q];
    }
    for (int i = tid; i < na; i += SL_THREADS) {
        mx[i] = J.coords[3 * (long long)(o + i)];
        my[i] = J.coords[3 * (long long)(o + i) + 1];
        mz[i] = J.coords[3 * (long long)(o + i) + 2];
        int z = J.zs[o + i], e = -1;
        for (int q = 0; q < T.n_el; q++)
            if (T.el_label[q] == z) e = q;
        mel[i] = e;
    }
    if (tid == 0) s_next = 0;
    // dense output: clear the rows, the units overwrite their blocks after the barrier
    if (J.out) {
        const int nrows = J.local ? bt.cnt : 1;
        for (int ai = 0; ai < nrows; ai++) {
            double *row = J.out + (long long)(J.local ? (o + bt.a0 + ai) : bt.mol) * J.ldo;
            for (int x = tid; x < T.D; x += SL_THREADS) row[x] = 0.0;
        }
    }
    __syncthreads();

    // ---- phase 1: neighbour lists ------------------------------------------------------------------
    for (int ai = warp; ai < bt.cnt; ai += SL_THREADS / 32) {
        const int c = bt.a0 + ai;
        double *dr = dd + ai * na;
        unsigned char *fr = fls + ai * na;
        for (int q0 = 0; q0 < na; q0 += 32) {
            int q = q0 + lane;
            if (q < na) {
                int fl = 0;
                double d = 0.0;
                if (q != c && mel[q] >= 0) {
                    double r2 = dist2_rn(mx[c], my[c], mz[c], mx[q], my[q], mz[q]);
                    d = __dsqrt_rn(r2);
                    fl = (int)(r2 < rcut2) | ((int)((d > eps) && (d <= rcut)) << 1);  // bit0: 2-body, bit1: 3-body
                    // (q - c)/|q - c|: v1 / norm2(v1) of calc_cos_angle (utils.f90:96-97), shared by all pairs
                    ux[ai * na + q] = __ddiv_rn(__dsub_rn(mx[q], mx[c]), d);
                    uy[ai * na + q] = __ddiv_rn(__dsub_rn(my[q], my[c]), d);
                    uz[ai * na + q] = __ddiv_rn(__dsub_rn(mz[q], mz[c]), d);
                }
                dr[q] = d;
                fr[q] = (unsigned char)fl;
            }
        }
        __syncwarp();
        int w = 0;
        for (int e = 0; e < T.n_el; e++) {
            if (lane == 0) sstart[ai * (MAXEL + 1) + e] = (short)w;
            for (int q0 = 0; q0 < na; q0 += 32) {
                int q = q0 + lane;
                bool keep = (q < na) && fr[q] && mel[q] == e;
                unsigned mk = __ballot_sync(0xffffffffu, keep);
                if (keep) nbidx[ai * na + w + __popc(mk & ((1u << lane) - 1u))] = (short)q;
                w += __popc(mk);
            }
        }
        if (lane == 0) sstart[ai * (MAXEL + 1) + T.n_el] = (short)w;
    }
    __syncthreads();

    // ---- phase 2: work units -----------------------------------------------------------------------
    const int ch = tid & (GS - 1), gbase = lane & ~(GS - 1);
    const unsigned gmask = ((1u << GS) - 1u) << gbase;
    // R(x_j0) = Rg * exp(-2 j0 h^2 c): Rg (first grid point) is computed once per entry and shared
    const double qm2 = exp(-2.0 * (ch * M2) * T.h2 * T.h2 * T.cst2);
    const double qm3 = exp(-2.0 * (ch * M3) * T.h3 * T.h3 * T.cst3);
    const bool loc = J.local != 0;
    const int nunits = (loc ? bt.cnt : T.n_el) * uc;  // global SLATM: unit = (element, block), summed over its centres
    for (;;) {
        int u = 0;
        if (ch == 0) u = atomicAdd(&s_next, 1);
        u = __shfl_sync(gmask, u, gbase);
        if (u >= nunits) break;
        const int ui = u / uc, ci = u - ui * uc;
        const int ec = loc ? mel[bt.a0 + ui] : ui;
        const int c_lo = loc ? ui : 0, c_hi = loc ? ui + 1 : bt.cnt;   // centres (batch-relative) of this unit
        const long long orow = loc ? (long long)(o + bt.a0 + ui) : (long long)bt.mol;
        double ssq = 0.0;  // this thread's share of the block's sum of squares
        if (ec >= 0) {
            const int *lay = T.lay_col + ec * T.nmb;
            const int pr = (loc && J.atom_prow) ? J.atom_prow[o + bt.a0 + ui] : -1;
            int col = -1, dcol = -1, nx = 0;
            if (ci < T.n_el) {
                const int sl = T.bop_slot[ec * MAXEL + ci];
                if (sl >= 0) { col = lay[sl]; dcol = T.slot_col[sl]; nx = T.nx2; }
                if (col >= 0) {
                    const int e2 = ci, j0 = ch * M2;
                    double a2[M2];
#pragma unroll
                    for (int j = 0; j < M2; j++) a2[j] = 0.0;
                    const double xa = sx2[j0];
                    for (int ai = c_lo; ai < c_hi; ai++) {
                    const int c = bt.a0 + ai;
                    if (mel[c] != ec) continue;
                    const short *st = sstart + ai * (MAXEL + 1);
                    const short *nb = nbidx + ai * na;
                    const double *dr = dd + ai * na;
                    const unsigned char *fr = fls + ai * na;
                    for (int n0 = st[e2]; n0 < st[e2 + 1]; n0 += GS) {
                        // my neighbour of this group-sized batch: distance and the grid-start ratio, shared by shuffles
                        double g_r = -1.0, g_R = 0.0;
                        if (n0 + ch < st[e2 + 1]) {
                            const int q = nb[n0 + ch];
                            // global SLATM counts an unordered same-element pair once (fslatm.f90:501-502)
                            if ((fr[q] & 1) && (loc || e2 != ec || q > c)) {
                                g_r = dr[q];
                                g_R = exp(-(2.0 * (sx2[0] - g_r) * T.h2 + T.h2 * T.h2) * T.cst2);
                            }
                        }
                        const int nt = min(GS, st[e2 + 1] - n0);
                        for (int t = 0; t < nt; t++) {
                            const double r = gshfl(g_r, t, gmask, gbase);
                            if (r < 0.0) continue;  // outside the 2-body cut-off (group-uniform)
                            const double d = xa - r;
                            double E = exp(-(d * d) * T.cst2), R = gshfl(g_R, t, gmask, gbase) * qm2;
#pragma unroll
                            for (int j = 0; j < M2; j++) { a2[j] += E; E *= R; R *= T.q2; }
                        }
                    }
                    }
                    const double c0 = (loc ? 0.5 : 1.0) * (T.el_w[ec] * T.el_w[e2]) * T.coeff2;  // fslatm.f90:610-617 / :487-491
                    unsigned bits0 = 0, bits1 = 0;
                    const int kt0 = (col + j0) >> 4;
#pragma unroll
                    for (int j = 0; j < M2; j++) {
                        const double v = a2[j] * (c0 * sinv2[j0 + j]);
                        if (j0 + j < nx) {
                            if (J.out) J.out[orow * J.ldo + dcol + j0 + j] = v;
                            if (pr >= 0) J.px[ec][(long long)pr * T.pld[ec] + col + j0 + j] = v;
                            ssq += v * v;
                            if (v != 0.0) {
                                int kt = ((col + j0 + j) >> 4) - kt0;
                                if (kt == 0) bits0 = 1; else if (kt == 1) bits1 = 1; else bits1 |= 2;
                            }
                        }
                    }
                    if (pr >= 0 && J.pmask[ec]) {
                        unsigned *mw = J.pmask[ec] + (long long)(pr >> 6) * J.maskw[ec];
                        if (bits0) atomicOr(mw + (kt0 >> 5), 1u << (kt0 & 31));
                        if (bits1 & 1) atomicOr(mw + ((kt0 + 1) >> 5), 1u << ((kt0 + 1) & 31));
                        if (bits1 & 2) atomicOr(mw + ((kt0 + 2) >> 5), 1u << ((kt0 + 2) & 31));
                    }
                }
            } else if (ci - T.n_el < T.nblk[ec]) {
                const Block3 B = T.blk[ec * MAXBLK + (ci - T.n_el)];
                col = B.col; nx = T.nx3;
                // dense column of this block's type: find it through the slot table
                dcol = T.slot_col[T.bot_slot[(B.e1 * MAXEL + ec) * MAXEL + B.e3] >> 1];
                if (col >= 0) {
                    const int j0 = ch * M3;
                    const double xa = sx3[j0];
                    double a3[M3];
#pragma unroll
                    for (int j = 0; j < M3; j++) a3[j] = 0.0;
                    for (int ai = c_lo; ai < c_hi; ai++) {
                    if (mel[bt.a0 + ai] != ec) continue;
                    const short *st = sstart + ai * (MAXEL + 1);
                    const short *nb = nbidx + ai * na;
                    const double *dr = dd + ai * na;
                    const unsigned char *fr = fls + ai * na;
                    const int n1 = st[B.e1 + 1] - st[B.e1], n3 = st[B.e3 + 1] - st[B.e3];
                    const int np = (B.e1 == B.e3) ? n1 * (n1 - 1) / 2 : n1 * n3;
                    for (int p0 = 0; p0 < np; p0 += GS) {
                        // my pair of this group-sized batch: geometry once, shared with the group by shuffles
                        double g_ang = 0.0, g_A = 0.0, g_B = 0.0, g_R = 0.0;
                        int p = p0 + ch;
                        if (p < np) {
                            int i, k;
                            if (B.e1 == B.e3) {
                                i = 0;
                                while (p >= n1 - 1 - i) { p -= n1 - 1 - i; i++; }
                                k = i + 1 + p;
                            } else {
                                i = p / n3;
                                k = p - i * n3;
                            }
                            const int qi = nb[st[B.e1] + i], qk = nb[st[B.e3] + k];
                            if ((fr[qi] & 2) && (fr[qk] & 2)) {
                                // reference roles: "i" is the atom of the type's first element (flip swaps)
                                const int I = B.flip ? qk : qi, K = B.flip ? qi : qk;
                                const double wx = __dsub_rn(mx[I], mx[K]), wy = __dsub_rn(my[I], my[K]), wz = __dsub_rn(mz[I], mz[K]);
                                const double dik = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(wx, wx), __dmul_rn(wy, wy)), __dmul_rn(wz, wz)));
                                if (dik <= rcut) {
                                    const double *uI = ux + ai * na, *vI = uy + ai * na, *zI = uz + ai * na;
                                    // angle at the centre: u_I . u_K  (calc_angle(i, c, k), utils.f90:52-80)
                                    double ca = __dadd_rn(__dadd_rn(__dmul_rn(uI[I], uI[K]), __dmul_rn(vI[I], vI[K])), __dmul_rn(zI[I], zI[K]));
                                    ca = fmin(1.0, fmax(-1.0, ca));
                                    g_ang = acos(ca);
                                    // w = (I-K)/|I-K|.  cak = cos at K between I and c = w . (-u_K); cai = cos at I between K and c
                                    // = (-w) . (-u_I): negations are exact, so these equal calc_cos_angle's values bit for bit
                                    const double nx_ = __ddiv_rn(wx, dik), ny_ = __ddiv_rn(wy, dik), nz_ = __ddiv_rn(wz, dik);
                                    const double cak = -__dadd_rn(__dadd_rn(__dmul_rn(nx_, uI[K]), __dmul_rn(ny_, vI[K])), __dmul_rn(nz_, zI[K]));
                                    const double cai = __dadd_rn(__dadd_rn(__dmul_rn(nx_, uI[I]), __dmul_rn(ny_, vI[I])), __dmul_rn(nz_, zI[I]));
                                    const double r = __dmul_rn(__dmul_rn(dr[I], dik), dr[K]);
                                    const double rinv3 = 1.0 / (r * r * r);
                                    g_A = B.c0 * rinv3;                  // (c0 + cos(x) c0 cak cai) / r^3
                                    g_B = (B.c0 * cak * cai) * rinv3;
                                    if (g_A == 0.0) g_A = 4.9e-324;      // keep "valid" distinguishable (never hit in practice)
                                    g_R = exp(-(2.0 * (sx3[0] - g_ang) * T.h3 + T.h3 * T.h3) * T.cst3);
                                }
                            }
                        }
                        const int nt = min(GS, np - p0);
                        for (int t = 0; t < nt; t++) {
                            const double A0 = gshfl(g_A, t, gmask, gbase);
                            if (A0 == 0.0) continue;  // pair failed a cut-off test (group-uniform)
                            const double ang = gshfl(g_ang, t, gmask, gbase), B0 = gshfl(g_B, t, gmask, gbase);
                            const double d = xa - ang;
                            double E = exp(-(d * d) * T.cst3), R = gshfl(g_R, t, gmask, gbase) * qm3;
#pragma unroll
                            for (int j = 0; j < M3; j++) { a3[j] += fma(scos3[j0 + j], B0, A0) * E; E *= R; R *= T.q3; }
                        }
                    }
                    }
                    unsigned bits0 = 0, bits1 = 0;
                    const int kt0 = (col + j0) >> 4;
#pragma unroll
                    for (int j = 0; j < M3; j++) {
                        const double v = a3[j];
                        if (j0 + j < nx) {
                            if (J.out) J.out[orow * J.ldo + dcol + j0 + j] = v;
                            if (pr >= 0) J.px[ec][(long long)pr * T.pld[ec] + col + j0 + j] = v;
                            ssq += v * v;
                            if (v != 0.0) {
                                int kt = ((col + j0 + j) >> 4) - kt0;
                                if (kt == 0) bits0 = 1; else if (kt == 1) bits1 = 1; else bits1 |= 2;
                            }
                        }
                    }
                    if (pr >= 0 && J.pmask[ec]) {
                        unsigned *mw = J.pmask[ec] + (long long)(pr >> 6) * J.maskw[ec];
                        if (bits0) atomicOr(mw + (kt0 >> 5), 1u << (kt0 & 31));
                        if (bits1 & 1) atomicOr(mw + ((kt0 + 1) >> 5), 1u << ((kt0 + 1) & 31));
                        if (bits1 & 2) atomicOr(mw + ((kt0 + 2) >> 5), 1u << ((kt0 + 2) & 31));
                    }
                }
            }
        }
        // block's sum of squares: fixed-order tree over the 8 threads of the group
        ssq += __shfl_xor_sync(gmask, ssq, 1);
        ssq += __shfl_xor_sync(gmask, ssq, 2);
        if (GS >= 8) ssq += __shfl_xor_sync(gmask, ssq, 4);
        if (GS >= 16) ssq += __shfl_xor_sync(gmask, ssq, 8);
        if (ch == 0) s_norm[ui * uc + ci] = ssq;
    }
    __syncthreads();

    // ---- phase 3: one-body slots (dense), padding columns and |row|^2 (packed) --------------------------
    if (!loc) {  // global SLATM one-body slots: Z * count (slatm.py:17-18)
        if (J.out)
            for (int s = tid; s < T.nmb; s += SL_THREADS)
                if (T.slot_kind[s] == 1) {
                    int cnt = 0;
                    for (int q = 0; q < na; q++) cnt += (mel[q] == T.slot_e0[s]);
                    J.out[(long long)bt.mol * J.ldo + T.slot_col[s]] = (double)(T.el_label[T.slot_e0[s]] * cnt);
                }
        return;
    }
    for (int ai = tid; ai < bt.cnt; ai += SL_THREADS) {
        const int c = bt.a0 + ai, ec = mel[c];
        if (J.out && ec >= 0) {
            double *row = J.out + (long long)(o + c) * J.ldo;
            for (int s = 0; s < T.nmb; s++)
                if (T.slot_kind[s] == 1 && T.slot_e0[s] == ec) row[T.slot_col[s]] = (double)J.zs[o + c];
        }
        if (J.atom_prow && ec >= 0) {
            const int pr = J.atom_prow[o + c];
            if (pr >= 0) {
                double t = 0.0;
                for (int q = 0; q < uc; q++) t += s_norm[ai * uc + q];
                J.pnorm[ec][pr] = t;
                for (int x = T.pdim[ec]; x < T.pld[ec]; x++) J.px[ec][(long long)pr * T.pld[ec] + x] = 0.0;
            }
        }
    }
}

}  // namespace aqml
