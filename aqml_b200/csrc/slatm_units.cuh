// aSLATM, "work unit" formulation (local representation; the default path).
//
// One CTA per batch of centre atoms of one molecule (a QM9-sized molecule is one batch).
//   phase 1  one warp per centre: distances / cut-off flags and the neighbour list sorted by element
//            (ballot compaction, list order preserved) -> shared memory.
//   phase 2  work units (centre atom, many-body block) are handed out dynamically to 8-thread groups.
//            A unit owns one output block (118 or 96 grid points): each thread keeps M consecutive grid
//            points in registers, walks the unit's neighbours (2-body) or neighbour pairs (3-body;
//            geometry of 8 pairs at a time is computed one pair per thread and shared through a
//            shared-memory mailbox), and
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

// exp(x) for x <= 0 in the inner loops: k = rint(x log2 e), r = x - k ln 2 (two-term), degree-13 Taylor
// polynomial on |r| <= 0.347 (truncation 4e-18), scaled by 2^k built from the exponent bits.  Relative error
// <= 2 ulp; results below 2^-1021 are flushed to 0 (the library exp keeps denormals).  About half the
// instructions of the library routine, whose constant set-up and special-case path dominate here.
__constant__ double c_expc[12] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0,
                                  1.0 / 362880.0,     1.0 / 40320.0,     1.0 / 5040.0,     1.0 / 720.0,
                                  1.0 / 120.0,        1.0 / 24.0,        1.0 / 6.0,        0.5};
__device__ __forceinline__ double exp_nonpos(double x) {
    if (x < -707.0) return 0.0;
    const double magic = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to nearest integer
    double t = fma(x, 1.4426950408889634, magic);
    const int k = __double2loint(t);
    t -= magic;
    double r = fma(t, -6.93147180369123816490e-01, x);
    r = fma(t, -1.90821492927058770002e-10, r);
    double p = c_expc[0];
#pragma unroll
    for (int i = 1; i < 12; i++) p = fma(p, r, c_expc[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return p * __hiloint2double((k + 1023) << 20, 0);
}

// GS = threads per work unit (4 or 8); M2/M3 = grid points per thread (GS*M >= nx)
// exp(x), x <= 0 (or -inf; NaN in, NaN out), from two 128-entry tables in shared memory: x log2(e) = k + j1/128 + j2/16384 +
// rho; exp(x) = 2^k T[j1] U[j2] exp(r), |r| <= 2.2e-5, cubic polynomial.  9 FP64 instructions (exp_nonpos: 18), 5e-16.
// tab[0..127] = 2^(j/128), tab[128..255] = 2^(j/16384)  (fill_exp_table).
__device__ __forceinline__ double exp_tab(double x, const double *tab) {
    const double xc = fmax(x, -708.0);
    const double magic = 6755399441055744.0;
    double t = fma(xc, 23637.115549924776 /* 2^14 log2(e) */, magic);
    const int n = __double2loint(t);
    t -= magic;
    double r = fma(t, -4.2306346131226746e-05 /* ln2 / 2^14, upper 27 bits: t * hi is exact */, xc);
    r = fma(t, -3.384964780863074e-13, r);
    double p = fma(r, 1.0 / 6.0, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double w = tab[(n >> 7) & 127] * tab[128 + (n & 127)];
    const double v = w * p;
    const double sc = __hiloint2double(((n >> 14) + 1023) << 20, 0);
    return (x < -707.0) ? 0.0 : ((x != x) ? x : v * sc);
}
__device__ __forceinline__ void fill_exp_table(double *tab, int tid, int nthreads) {
    for (int j = tid; j < 256; j += nthreads) tab[j] = (j < 128) ? exp2((double)j / 128.0) : exp2((double)(j - 128) / 16384.0);
}

__device__ __forceinline__ int mel_of(const SlatmJob &J, const SlatmDev &T, long long atom) {
    const int z = J.zs[atom];
    int e = -1;
    for (int q = 0; q < T.n_el; q++)
        if (T.el_label[q] == z) e = q;
    return e;
}

template <int M2, int M3, int GS>
__global__ void __launch_bounds__(SL_THREADS, (GS == 16) ? 5 : (GS == 8) ? 4 : 3)
slatm_units_kernel(const SlatmDev T, const SlatmJob J, const Batch *batches, int na_max, int nbn_max, int ba_max, int uc) {
    extern __shared__ __align__(16) double sm[];
    double *mx = sm, *my = mx + na_max, *mz = my + na_max;
    double *dd = mz + na_max;                          // [ba][na] distance centre -> atom q
    double *ux = dd + nbn_max, *uy = ux + nbn_max, *uz = uy + nbn_max;  // [ba][na] unit vector centre -> q
    double *s_norm = uz + nbn_max;                     // [ba][uc] sum of squares of each written block
    double *s_max = s_norm + ba_max * uc;              // [ba][uc] largest finite |value| of each written block
    double *sx2 = s_max + ba_max * uc, *sinv2 = sx2 + GS * M2, *sx3 = sinv2 + GS * M2, *scos3 = sx3 + GS * M3;
    int *mel = reinterpret_cast<int *>(scos3 + GS * M3);
    int *s_exp = mel + na_max;                         // [ba] digit-plane exponent of each centre's packed row
    short *nbidx = reinterpret_cast<short *>(s_exp + ba_max);   // [ba][na] neighbours sorted by element
    short *sstart = nbidx + nbn_max;                           // [ba][MAXEL+1]
    unsigned char *fls = reinterpret_cast<unsigned char *>(sstart + ba_max * (MAXEL + 1));  // [ba][na]
    __shared__ double4 s_mbox[SL_THREADS];  // per thread: the geometry of "its" entry, read by its whole group
    __shared__ int s_next, s_next3, s_fetch[SL_THREADS / GS];
    __shared__ double s_exptab[256];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double eps = 2.220446049250313e-16;
    const double rcut = T.rcut, rcut2 = __dmul_rn(rcut, rcut);
    const Batch bt = batches[blockIdx.x];
    const int o = J.mol_off[bt.mol], na = J.mol_off[bt.mol + 1] - o;

    fill_exp_table(s_exptab, tid, SL_THREADS);
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
    if (tid == 0) { s_next = 0; s_next3 = 0; }
    if (J.local) for (int i = tid; i < bt.cnt * uc; i += SL_THREADS) { s_norm[i] = 0.0; s_max[i] = 0.0; }
    // packed output: clear the rows too -- units without a single entry are then never visited
    if (J.local && J.atom_prow) {
        for (int ai = 0; ai < bt.cnt; ai++) {
            const int e = mel_of(J, T, o + bt.a0 + ai);
            const int pr = J.atom_prow[o + bt.a0 + ai];
            if (e >= 0 && pr >= 0) {
                double *row = J.px[e] + (long long)pr * T.pld[e];
                for (int x = tid; x < T.pld[e]; x += SL_THREADS) row[x] = 0.0;
            }
        }
    }
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

    // ---- phase 2: work units, executed warp-lockstep ---------------------------------------------------
    // Every 8-thread group owns one unit at a time, but the four groups of a warp walk through the same
    // code together: per pass each group takes the next batch of GS entries of ITS unit (geometry by one
    // thread per entry -> mailbox), then all groups run the GS-entry accumulation loop.  Divergence is
    // confined to unit boundaries (write-out + fetching the next non-empty unit); with independently
    // looping groups only 46 % of the lanes were active per issued instruction.
    constexpr unsigned FULL = 0xffffffffu;
    const int ch = tid & (GS - 1), gbase = lane & ~(GS - 1), gtid0 = tid & ~(GS - 1), gid = tid / GS;
    const unsigned gmask = ((1u << GS) - 1u) << gbase;
    // R(x_j0) = Rg * exp(-2 j0 h^2 c): Rg (first grid point) is computed once per entry and shared
    const double qm2 = exp(-2.0 * (ch * M2) * T.h2 * T.h2 * T.cst2);
    const double qm3 = exp(-2.0 * (ch * M3) * T.h3 * T.h3 * T.cst3);
    const bool loc = J.local != 0;
    const int nown = loc ? bt.cnt : T.n_el;  // unit owners: centre atoms (aSLATM) or elements (SLATM, summed over centres)

    auto group_fetch = [&](int *counter) -> int {  // one atomic per group, broadcast through shared memory
        if (ch == 0) s_fetch[gid] = atomicAdd(counter, 1);
        __syncwarp(gmask);
        const int u = s_fetch[gid];
        __syncwarp(gmask);
        return u;
    };
    // store one finished block: dense row and/or packed operand row, sum of squares, k-tile mask bits
    auto write_block = [&](const double *acc, int m, int nx, int ui, int cidx, int ec, int col, int dcol) {
        const long long orow = loc ? (long long)(o + bt.a0 + ui) : (long long)bt.mol;
        const int pr = (loc && J.atom_prow) ? J.atom_prow[o + bt.a0 + ui] : -1;
        const int j0 = ch * m, kt0 = (col + j0) >> 4;
        unsigned bits = 0;
        double ssq = 0.0, vmax = 0.0;
        for (int j = 0; j < m; j++) {
            const double v = acc[j];
            if (j0 + j < nx) {
                if (J.out) J.out[orow * J.ldo + dcol + j0 + j] = v;
                if (pr >= 0) J.px[ec][(long long)pr * T.pld[ec] + col + j0 + j] = v;
                ssq += v * v;
                if (fabs(v) <= 1.7976931348623157e308) vmax = fmax(vmax, fabs(v));
                if (v != 0.0) bits |= 1u << min(((col + j0 + j) >> 4) - kt0, 2);
            }
        }
        if (pr >= 0 && J.pmask[ec]) {
            unsigned *mw = J.pmask[ec] + (long long)(pr >> 6) * J.maskw[ec];
            for (int q = 0; q < 3; q++)
                if (bits & (1u << q)) atomicOr(mw + ((kt0 + q) >> 5), 1u << ((kt0 + q) & 31));
        }
#pragma unroll
        for (int o = 1; o < GS; o <<= 1) {
            ssq += __shfl_xor_sync(gmask, ssq, o);
            vmax = fmax(vmax, __shfl_xor_sync(gmask, vmax, o));
        }
        if (ch == 0 && loc) { s_norm[ui * uc + cidx] = ssq; s_max[ui * uc + cidx] = vmax; }
    };

    // ---------- two-body units: (owner, partner element) ----------
    {
        int ui = 0, e2 = 0, ec = -1, col = -1, dcol = -1, ai = 0, aiend = 0, n0 = -1, nend = 0;
        bool active = true;
        double a2[M2];
        const int j0 = ch * M2;
        const double xa = sx2[j0];
        auto seek = [&]() -> bool {  // next centre of this unit with entries left
            while (ai < aiend) {
                if (mel[bt.a0 + ai] == ec) {
                    if (n0 < 0) { n0 = sstart[ai * (MAXEL + 1) + e2]; nend = sstart[ai * (MAXEL + 1) + e2 + 1]; }
                    if (n0 < nend) return true;
                }
                ai++;
                n0 = -1;
            }
            return false;
        };
        auto fetch = [&]() {
            for (;;) {
                const int u = group_fetch(&s_next);
                if (u >= nown * T.n_el) { active = false; return; }
                ui = u / T.n_el;
                e2 = u - ui * T.n_el;
                ec = loc ? mel[bt.a0 + ui] : ui;
                if (ec < 0) continue;
                const int sl = T.bop_slot[ec * MAXEL + e2];
                if (sl < 0) continue;
                col = T.lay_col[ec * T.nmb + sl];  // global mode: -1 unless this element owns the block
                if (col < 0) continue;
                dcol = T.slot_col[sl];
                ai = loc ? ui : 0;
                aiend = loc ? ui + 1 : bt.cnt;
                n0 = -1;
                if (seek()) {
#pragma unroll
                    for (int j = 0; j < M2; j++) a2[j] = 0.0;
                    return;
                }
            }
        };
        fetch();
        while (__any_sync(FULL, active)) {
            double4 g = make_double4(-1.0, 0.0, 0.0, 0.0);
            if (active && n0 + ch < nend) {
                const int q = nbidx[ai * na + n0 + ch];
                // global SLATM counts an unordered same-element pair once (fslatm.f90:501-502)
                if ((fls[ai * na + q] & 1) && (loc || e2 != ec || q > bt.a0 + ai)) {
                    g.x = dd[ai * na + q];
                    g.y = exp(-(2.0 * (sx2[0] - g.x) * T.h2 + T.h2 * T.h2) * T.cst2);
                }
            }
            s_mbox[tid] = g;
            __syncwarp();
            if (active) {
#pragma unroll 1
                for (int t = 0; t < GS; t++) {
                    const double4 mb = s_mbox[gtid0 + t];
                    if (mb.x < 0.0) continue;  // no entry / outside the 2-body cut-off (group-uniform)
                    const double d = xa - mb.x;
                    double E = exp_tab(-(d * d) * T.cst2, s_exptab), R = mb.y * qm2;
#pragma unroll
                    for (int j = 0; j < M2; j++) { a2[j] += E; E *= R; R *= T.q2; }
                }
            }
            __syncwarp();
            if (active) {
                n0 += GS;
                if (!seek()) {
                    const double c0 = (loc ? 0.5 : 1.0) * (T.el_w[ec] * T.el_w[e2]) * T.coeff2;  // fslatm.f90:610-617 / :487-491
#pragma unroll
                    for (int j = 0; j < M2; j++) a2[j] *= c0 * sinv2[j0 + j];
                    write_block(a2, M2, T.nx2, ui, e2, ec, col, dcol);
                    fetch();
                }
            }
        }
    }

    // ---------- three-body units: (owner, [z1, zc, z3] block) ----------
    {
        int ui = 0, bi = 0, ec = -1, ai = 0, aiend = 0, p0 = -1, np = 0, n1 = 0, n3 = 0;
        Block3 B;
        B.col = -1;
        bool active = true;
        double a3[M3];
        const int j0 = ch * M3;
        const double xa = sx3[j0];
        const int nbmax = uc - T.n_el;
        auto seek = [&]() -> bool {
            while (ai < aiend) {
                if (mel[bt.a0 + ai] == ec) {
                    if (p0 < 0) {
                        const short *st = sstart + ai * (MAXEL + 1);
                        n1 = st[B.e1 + 1] - st[B.e1];
                        n3 = st[B.e3 + 1] - st[B.e3];
                        np = (B.e1 == B.e3) ? n1 * (n1 - 1) / 2 : n1 * n3;
                        p0 = 0;
                    }
                    if (p0 < np) return true;
                }
                ai++;
                p0 = -1;
            }
            return false;
        };
        auto fetch = [&]() {
            for (;;) {
                const int u = group_fetch(&s_next3);
                if (u >= nown * nbmax) { active = false; return; }
                ui = u / nbmax;
                bi = u - ui * nbmax;
                ec = loc ? mel[bt.a0 + ui] : ui;
                if (ec < 0 || bi >= T.nblk[ec]) continue;
                B = T.blk[ec * MAXBLK + bi];
                if (B.col < 0) continue;
                ai = loc ? ui : 0;
                aiend = loc ? ui + 1 : bt.cnt;
                p0 = -1;
                if (seek()) {
#pragma unroll
                    for (int j = 0; j < M3; j++) a3[j] = 0.0;
                    return;
                }
            }
        };
        fetch();
        while (__any_sync(FULL, active)) {
            double4 g = make_double4(0.0, 0.0, 0.0, 0.0);  // (A0, angle, B0, Rg); A0 == 0 marks "no entry"
            if (active && p0 + ch < np) {
                const short *st = sstart + ai * (MAXEL + 1);
                const short *nb = nbidx + ai * na;
                const double *dr = dd + ai * na;
                const unsigned char *fr = fls + ai * na;
                int p = p0 + ch, i, k;
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
                        g.y = acos(ca);
                        // w = (I-K)/|I-K|.  cak = cos at K between I and c = w . (-u_K); cai = cos at I between K and c
                        // = (-w) . (-u_I): negations are exact, so these equal calc_cos_angle's values bit for bit
                        const double nx_ = __ddiv_rn(wx, dik), ny_ = __ddiv_rn(wy, dik), nz_ = __ddiv_rn(wz, dik);
                        const double cak = -__dadd_rn(__dadd_rn(__dmul_rn(nx_, uI[K]), __dmul_rn(ny_, vI[K])), __dmul_rn(nz_, zI[K]));
                        const double cai = __dadd_rn(__dadd_rn(__dmul_rn(nx_, uI[I]), __dmul_rn(ny_, vI[I])), __dmul_rn(nz_, zI[I]));
                        const double r = __dmul_rn(__dmul_rn(dr[I], dik), dr[K]);
                        const double rinv3 = 1.0 / (r * r * r);
                        g.x = B.c0 * rinv3;                  // (c0 + cos(x) c0 cak cai) / r^3
                        g.z = (B.c0 * cak * cai) * rinv3;
                        if (g.x == 0.0) g.x = 4.9e-324;      // keep "valid" distinguishable (never hit in practice)
                        g.w = exp(-(2.0 * (sx3[0] - g.y) * T.h3 + T.h3 * T.h3) * T.cst3);
                    }
                }
            }
            s_mbox[tid] = g;
            __syncwarp();
            if (active) {
#pragma unroll 1
                for (int t = 0; t < GS; t++) {
                    const double4 mb = s_mbox[gtid0 + t];
                    if (mb.x == 0.0) continue;  // no entry / pair failed a cut-off test (group-uniform)
                    const double d = xa - mb.y;
                    double E = exp_tab(-(d * d) * T.cst3, s_exptab), R = mb.w * qm3;
#pragma unroll
                    for (int j = 0; j < M3; j++) { a3[j] += fma(scos3[j0 + j], mb.z, mb.x) * E; E *= R; R *= T.q3; }
                }
            }
            __syncwarp();
            if (active) {
                p0 += GS;
                if (!seek()) {
                    const int dcol = T.slot_col[T.bot_slot[(B.e1 * MAXEL + ec) * MAXEL + B.e3] >> 1];
                    write_block(a3, M3, T.nx3, ui, T.n_el + bi, ec, B.col, dcol);
                    fetch();
                }
            }
        }
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
                double t = 0.0, vmax = 0.0;
                for (int q = 0; q < uc; q++) { t += s_norm[ai * uc + q]; vmax = fmax(vmax, s_max[ai * uc + q]); }
                J.pnorm[ec][pr] = t;
                if (J.pexpo[ec]) {
                    const int e = i8_row_exponent(vmax);
                    J.pexpo[ec][pr] = e;
                    s_exp[ai] = e;
                }
            }
        }
    }
    // ---- phase 4: the rows' int8 digit planes (i8pair.cuh), straight from the rows this CTA has just written (they are
    // still in L2): no separate pass over the packed operand, no second kernel.  One item = 16 columns of one row.
    if (!J.atom_prow) return;
    int ktmax = 0;
    for (int q = 0; q < T.n_el; q++)
        if (J.psl[q]) ktmax = max(ktmax, T.pld[q] >> 4);
    if (ktmax == 0) return;
    __syncthreads();   // rows (global memory, written by other threads of this CTA) and s_exp are complete
    for (int item = tid; item < bt.cnt * ktmax; item += SL_THREADS) {
        const int ai = item % bt.cnt, kt = item / bt.cnt, c = bt.a0 + ai, ec = mel[c];
        if (ec < 0 || !J.psl[ec]) continue;
        const int pr = J.atom_prow[o + c], KT = T.pld[ec] >> 4;
        if (pr < 0 || kt >= KT) continue;
        const double sc = i8::pow2(48 - s_exp[ai]);   // |e| <= 1000: exact power of two
        const double *xr = J.px[ec] + (long long)pr * T.pld[ec] + kt * 16;
        double t[16];
#pragma unroll
        for (int q = 0; q < 16; q += 2) {
            const double2 v = __ldcg(reinterpret_cast<const double2 *>(xr + q));
            t[q] = v.x;
            t[q + 1] = v.y;
        }
        i8_split16(t, sc, J.psl[ec] + (((long long)(pr >> 7) * (KT + 1) + kt) * I8_S) * 2048 + (pr & 127) * 16);
    }
}

}  // namespace aqml
