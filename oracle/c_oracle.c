/*
 * C + OpenMP restatement of the reference's Fortran hot loops.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY -- never linked into the product library.
 *
 * Restates (binghuang2018/aqml, paths relative to the reference tree):
 *   coreml/cml/representation/fslatm.f90   calc_sbot:4  calc_sbot_local:206
 *                                          calc_sbop:413 calc_sbop_local:531
 *   coreml/cml/representation/utils.f90    linspace:29 calc_angle:52 calc_cos_angle:82
 *   coreml/cml/atomdb.f90                  zstar:90
 *   coreml/cml/fkernels.f90                calc_lk_gaussian:23 calc_lk_laplacian:97
 *                                          fgaussian_kernel:327 flaplacian_kernel:361
 *   coreml/cml/fdistance.f90               fl1_distance:3 fl2_distance:25
 *
 * Same loop structure and OpenMP placement as the Fortran: direct-difference
 * distances, no BLAS, one pass per sigma for the global kernels.  The reference
 * itself cannot be compiled in this image (no gfortran), so this file is the
 * "port" CPU baseline; it is validated against oracle/np_oracle.py, which in turn
 * reproduces the reference's demo known-answer curve (demo/example/README.md:56-63).
 *
 * Arrays are C-order: x(natoms, D) rows == the Fortran x(D, natoms) columns.
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -march=native -ffp-contract=off).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifdef _OPENMP
#include <omp.h>
#endif

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* atomdb.f90:90-116 (single-precision literals promoted to real*8) */
static double zstar(int z) {
    static const int zs[20] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 35, 53};
    static const float ze[20] = {1.0f, 1.688f, 1.279f, 1.912f, 2.421f, 3.136f, 3.834f, 4.453f, 5.100f, 5.758f,
                                 2.507f, 3.308f, 4.066f, 4.285f, 4.886f, 5.482f, 6.116f, 6.764f, 9.028f, 11.612f};
    double r = 0.0;
    for (int j = 0; j < 20; j++) if (z == zs[j]) r = (double)ze[j];
    return r;
}

static double weight(int z, int izeff) { return izeff ? zstar(z) : (double)(z % 1000); }

/* utils.f90:29-50 */
static void linspace(double x0, double x1, int nx, double *xs) {
    double step = (x1 - x0) / (nx - 1);
    for (int i = 0; i < nx; i++) xs[i] = x0 + i * step;
}

/* utils.f90:82-103 : cosine of the angle at b */
static double cos_angle(const double *a, const double *b, const double *c) {
    double v1[3], v2[3];
    for (int d = 0; d < 3; d++) { v1[d] = a[d] - b[d]; v2[d] = c[d] - b[d]; }
    double n1 = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
    double n2 = sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
    for (int d = 0; d < 3; d++) { v1[d] /= n1; v2[d] /= n2; }
    return v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
}

/* utils.f90:52-80 */
static double angle(const double *a, const double *b, const double *c) {
    double ca = cos_angle(a, b, c);
    if (ca > 1.0) ca = 1.0;
    if (ca < -1.0) ca = -1.0;
    return acos(ca);
}

/* fslatm.f90:413-527 (ia < 0, global) and :531-655 (ia >= 0, local, 0-based) */
void oracle_calc_sbop(const double *coords, const int *zs, int na, int ia, int izeff, int z1, int z2,
                      double rcut, int nx, double dgrid, double sigma, double coeff, double rpower, double *ys) {
    int *ias1 = (int *)malloc(sizeof(int) * (na + 1)), *ias2 = (int *)malloc(sizeof(int) * (na + 1));
    int n1 = 0, n2 = 0;
    for (int i = 0; i < na; i++) if (zs[i] == z1) ias1[n1++] = i;
    for (int i = 0; i < na; i++) if (zs[i] == z2) ias2[n2++] = i;
    double *xs = (double *)malloc(sizeof(double) * nx), *xs0 = (double *)malloc(sizeof(double) * nx);
    linspace(0.1, rcut, nx, xs);
    memset(ys, 0, sizeof(double) * nx);
    double pref = (ia >= 0) ? 0.5 : 1.0;
    double c0 = pref * (weight(z1, izeff) * weight(z2, izeff)) * coeff;
    double inv_sigma = -0.5 / (sigma * sigma);
    for (int i = 0; i < nx; i++) xs0[i] = c0 / pow(xs[i], rpower) * dgrid;
    double rcut2 = rcut * rcut;
    int same = (z1 == z2);
    for (int p1 = 0; p1 < n1; p1++) {
        for (int p2 = same ? p1 + 1 : 0; p2 < n2; p2++) {
            int i = ias1[p1], j = ias2[p2];
            if (ia >= 0 && i != ia && j != ia) continue;
            double dx = coords[3 * i] - coords[3 * j], dy = coords[3 * i + 1] - coords[3 * j + 1],
                   dz = coords[3 * i + 2] - coords[3 * j + 2];
            double r = dx * dx + dy * dy + dz * dz;
            if (r < rcut2) {
                double sr = sqrt(r);
                for (int g = 0; g < nx; g++) {
                    double t = xs[g] - sr;
                    ys[g] += xs0[g] * exp(inv_sigma * (t * t));
                }
            }
        }
    }
    free(ias1); free(ias2); free(xs); free(xs0);
}

/* fslatm.f90:4-202 (ia < 0) and :206-409 (ia >= 0); intended index semantics for the
 * global z1 != z3 branch (Trash/fslatm_0.f90:263), see DESIGN.md. */
void oracle_calc_sbot(const double *coords, const int *zs, int na, int ia, int izeff, int z1, int z2, int z3,
                      double rcut, int nx, double dgrid, double sigma, double coeff, double *ys) {
    const double eps = DBL_EPSILON;
    const double pi = 4.0 * atan(1.0);
    memset(ys, 0, sizeof(double) * nx);
    double *dm = (double *)calloc((size_t)na * na, sizeof(double));
    for (int i = 0; i < na; i++)
        for (int j = i + 1; j < na; j++) {
            double dx = coords[3 * j] - coords[3 * i], dy = coords[3 * j + 1] - coords[3 * i + 1],
                   dz = coords[3 * j + 2] - coords[3 * i + 2];
            double n = sqrt(dx * dx + dy * dy + dz * dz);
            dm[(size_t)i * na + j] = n; dm[(size_t)j * na + i] = n;
        }
    int *ias1 = (int *)malloc(sizeof(int) * (na + 1)), *ias2 = (int *)malloc(sizeof(int) * (na + 1)),
        *ias3 = (int *)malloc(sizeof(int) * (na + 1));
    int n1 = 0, n2 = 0, n3 = 0;
    for (int i = 0; i < na; i++) if (zs[i] == z1) ias1[n1++] = i;
    for (int i = 0; i < na; i++) if (zs[i] == z2) ias2[n2++] = i;
    for (int i = 0; i < na; i++) if (zs[i] == z3) ias3[n3++] = i;
    if (ia >= 0) {
        int found = 0;
        for (int p = 0; p < n2; p++) if (ias2[p] == ia) found = 1;
        if (!found) { free(dm); free(ias1); free(ias2); free(ias3); return; }
        ias2[0] = ia; n2 = 1;
    }
    double d2r = pi / 180.0, a0 = -20.0 * d2r, a1 = pi + 20.0 * d2r;
    double *xs = (double *)malloc(sizeof(double) * nx), *cos_xs = (double *)malloc(sizeof(double) * nx);
    linspace(a0, a1, nx, xs);
    double prefactor = 1.0 / 3.0;
    double c0 = prefactor * (weight(z1, izeff) * weight(z2, izeff) * weight(z3, izeff)) * coeff * dgrid;
    double inv_sigma = -1.0 / (2 * sigma * sigma);
    for (int g = 0; g < nx; g++) cos_xs[g] = cos(xs[g]) * c0;
    int same = (z1 == z3);
    for (int p1 = 0; p1 < n1; p1++) {
        for (int p2 = 0; p2 < n2; p2++) {
            int i = ias1[p1], j = ias2[p2];
            double dij = dm[(size_t)i * na + j];
            if (!(dij > eps && dij <= rcut)) continue;
            for (int p3 = same ? p1 + 1 : 0; p3 < n3; p3++) {
                int k = ias3[p3];
                if (dm[(size_t)i * na + k] > rcut) continue;
                double dkj = dm[(size_t)j * na + k];
                if (!(dkj > eps && dkj <= rcut)) continue;
                double ang = angle(coords + 3 * i, coords + 3 * j, coords + 3 * k);
                double cak = cos_angle(coords + 3 * i, coords + 3 * k, coords + 3 * j);
                double cai = cos_angle(coords + 3 * k, coords + 3 * i, coords + 3 * j);
                double r = dm[(size_t)i * na + j] * dm[(size_t)i * na + k] * dm[(size_t)k * na + j];
                double r3 = r * r * r;
                for (int g = 0; g < nx; g++) {
                    double t = xs[g] - ang;
                    ys[g] += (c0 + cos_xs[g] * cak * cai) / r3 * exp((t * t) * inv_sigma);
                }
            }
        }
    }
    free(dm); free(ias1); free(ias2); free(ias3); free(xs); free(cos_xs);
}

/*
 * Batch driver: representation/core.py:430-554 for many molecules.  mbtypes is (nmb,3)
 * with -1 padding.  local: out is (A_total, D); global: out is (nmol, D).
 * OpenMP over atoms/molecules (the reference parallelises inside each Fortran call and
 * loops atoms x mbtypes in Python; here every host thread is kept busy instead).
 */
void oracle_generate_slatm_batch(const double *coords, const int *zs, const int *mol_off, int nmol,
                                 const int *mbtypes, int nmb, int local, int izeff,
                                 double sigma2, double sigma3, double dgrid2, double dgrid3, double rcut,
                                 double rpower, int nx2, int nx3, int D, double *out) {
    const double pi = 4.0 * atan(1.0);
    double coeff2 = 1.0 / sqrt(2 * sigma2 * sigma2 * pi), coeff3 = 1.0 / sqrt(2 * sigma3 * sigma3 * pi);
    if (local) {
        int A = mol_off[nmol];
        int *atom_mol = (int *)malloc(sizeof(int) * (A + 1));
        for (int m = 0; m < nmol; m++) for (int a = mol_off[m]; a < mol_off[m + 1]; a++) atom_mol[a] = m;
#pragma omp parallel for schedule(dynamic, 4)
        for (int a = 0; a < A; a++) {
            int m = atom_mol[a], o = mol_off[m], na = mol_off[m + 1] - o, ia = a - o;
            double *row = out + (size_t)a * D;
            int col = 0;
            for (int t = 0; t < nmb; t++) {
                const int *mb = mbtypes + 3 * t;
                if (mb[1] < 0) { row[col] = (zs[a] == mb[0]) ? (double)mb[0] : 0.0; col += 1; }
                else if (mb[2] < 0) {
                    oracle_calc_sbop(coords + 3 * (size_t)o, zs + o, na, ia, izeff, mb[0], mb[1], rcut, nx2, dgrid2,
                                     sigma2, coeff2, rpower, row + col);
                    col += nx2;
                } else {
                    oracle_calc_sbot(coords + 3 * (size_t)o, zs + o, na, ia, izeff, mb[0], mb[1], mb[2], rcut, nx3,
                                     dgrid3, sigma3, coeff3, row + col);
                    col += nx3;
                }
            }
        }
        free(atom_mol);
    } else {
#pragma omp parallel for schedule(dynamic, 1)
        for (int m = 0; m < nmol; m++) {
            int o = mol_off[m], na = mol_off[m + 1] - o;
            double *row = out + (size_t)m * D;
            int col = 0;
            for (int t = 0; t < nmb; t++) {
                const int *mb = mbtypes + 3 * t;
                if (mb[1] < 0) {
                    int cnt = 0;
                    for (int a = 0; a < na; a++) cnt += (zs[o + a] == mb[0]);
                    row[col] = (double)(mb[0] * cnt); col += 1;
                } else if (mb[2] < 0) {
                    oracle_calc_sbop(coords + 3 * (size_t)o, zs + o, na, -1, izeff, mb[0], mb[1], rcut, nx2, dgrid2,
                                     sigma2, coeff2, rpower, row + col);
                    col += nx2;
                } else {
                    oracle_calc_sbot(coords + 3 * (size_t)o, zs + o, na, -1, izeff, mb[0], mb[1], mb[2], rcut, nx3,
                                     dgrid3, sigma3, coeff3, row + col);
                    col += nx3;
                }
            }
        }
    }
}

/* fkernels.f90:327-359.  a (na,D), b (nb,D); k Fortran-order (na,nb): k[j + na*i]. */
void oracle_fgaussian_kernel(const double *a, int na, const double *b, int nb, int D, double *k, double sigma) {
    double inv_sigma = -0.5 / (sigma * sigma);
#pragma omp parallel for
    for (int i = 0; i < nb; i++)
        for (int j = 0; j < na; j++) {
            const double *aj = a + (size_t)j * D, *bi = b + (size_t)i * D;
            double s = 0.0;
            for (int f = 0; f < D; f++) { double t = aj[f] - bi[f]; s += t * t; }
            k[(size_t)j + (size_t)na * i] = exp(inv_sigma * s);
        }
}

/* fkernels.f90:361-387 */
void oracle_flaplacian_kernel(const double *a, int na, const double *b, int nb, int D, double *k, double sigma) {
    double inv_sigma = -1.0 / sigma;
#pragma omp parallel for
    for (int i = 0; i < nb; i++)
        for (int j = 0; j < na; j++) {
            const double *aj = a + (size_t)j * D, *bi = b + (size_t)i * D;
            double s = 0.0;
            for (int f = 0; f < D; f++) s += fabs(aj[f] - bi[f]);
            k[(size_t)j + (size_t)na * i] = exp(inv_sigma * s);
        }
}

/* fdistance.f90:25-44 (p=2) and :3-22 (p=1).  D Fortran-order (n1,n2). */
void oracle_fl2_distance(const double *A, const double *B, int n1, int n2, int nc, double *Dm) {
#pragma omp parallel for
    for (int i = 0; i < n2; i++)
        for (int j = 0; j < n1; j++) {
            const double *aj = A + (size_t)j * nc, *bi = B + (size_t)i * nc;
            double s = 0.0;
            for (int f = 0; f < nc; f++) { double t = aj[f] - bi[f]; s += t * t; }
            Dm[(size_t)j + (size_t)n1 * i] = sqrt(s);
        }
}

void oracle_fl1_distance(const double *A, const double *B, int n1, int n2, int nc, double *Dm) {
#pragma omp parallel for
    for (int i = 0; i < n2; i++)
        for (int j = 0; j < n1; j++) {
            const double *aj = A + (size_t)j * nc, *bi = B + (size_t)i * nc;
            double s = 0.0;
            for (int f = 0; f < nc; f++) s += fabs(aj[f] - bi[f]);
            Dm[(size_t)j + (size_t)n1 * i] = s;
        }
}

/* fkernels.f90:23-93.  sigmas C-order (nsigmas, zmax); kernels C-order (nsigmas, nm1, nm2). */
void oracle_calc_lk_gaussian(const double *x1, const double *x2, const int *zs1, const int *zs2,
                             const int *nas1, const int *nas2, int zmax, int iab, const double *sigmas,
                             int nm1, int nm2, int nsigmas, int D, double *kernels) {
    int *ias = (int *)malloc(sizeof(int) * (nm1 + 1)), *jas = (int *)malloc(sizeof(int) * (nm2 + 1));
    ias[0] = 0; for (int i = 0; i < nm1; i++) ias[i + 1] = ias[i] + nas1[i];
    jas[0] = 0; for (int j = 0; j < nm2; j++) jas[j + 1] = jas[j] + nas2[j];
    double *inv_sigma2 = (double *)malloc(sizeof(double) * nsigmas * zmax);
    for (int q = 0; q < nsigmas * zmax; q++) inv_sigma2[q] = -0.5 / (sigmas[q] * sigmas[q]);
    memset(kernels, 0, sizeof(double) * (size_t)nsigmas * nm1 * nm2);
    int mx1 = 0, mx2 = 0;
    for (int i = 0; i < nm1; i++) if (nas1[i] > mx1) mx1 = nas1[i];
    for (int j = 0; j < nm2; j++) if (nas2[j] > mx2) mx2 = nas2[j];
#pragma omp parallel
    {
        double *dsij = (double *)malloc(sizeof(double) * (size_t)(mx1 > 0 ? mx1 : 1) * (mx2 > 0 ? mx2 : 1));
#pragma omp for schedule(dynamic, 1)
        for (int a = 0; a < nm1; a++) {
            int ni = nas1[a];
            for (int b = 0; b < nm2; b++) {
                int nj = nas2[b];
                for (int i = 0; i < ni; i++)
                    for (int j = 0; j < nj; j++) {
                        int ia = i + ias[a], ja = j + jas[b];
                        double d = (double)1.e9f;
                        if (iab || zs1[ia] == zs2[ja]) {
                            const double *p = x1 + (size_t)ia * D, *q = x2 + (size_t)ja * D;
                            double s = 0.0;
                            for (int f = 0; f < D; f++) { double t = p[f] - q[f]; s += t * t; }
                            d = s;
                        }
                        dsij[(size_t)i * nj + j] = d;
                    }
                for (int k = 0; k < nsigmas; k++) {
                    double s = 0.0;
                    for (int j = 0; j < nj; j++)      /* Fortran array sum runs column-major: i fastest */
                        for (int i = 0; i < ni; i++)
                            s += exp(dsij[(size_t)i * nj + j] * inv_sigma2[k * zmax + zs1[i + ias[a]] - 1]);
                    kernels[((size_t)k * nm1 + a) * nm2 + b] = s;
                }
            }
        }
        free(dsij);
    }
    free(ias); free(jas); free(inv_sigma2);
}

/* fkernels.f90:97-183.  sigmas (nsigmas,); kernels C-order (nsigmas, nm1, nm2). */
void oracle_calc_lk_laplacian(const double *x1, const double *x2, const int *nas1, const int *nas2,
                              const double *sigmas, int nm1, int nm2, int nsigmas, int D, double *kernels) {
    int *ias = (int *)malloc(sizeof(int) * (nm1 + 1)), *jas = (int *)malloc(sizeof(int) * (nm2 + 1));
    ias[0] = 0; for (int i = 0; i < nm1; i++) ias[i + 1] = ias[i] + nas1[i];
    jas[0] = 0; for (int j = 0; j < nm2; j++) jas[j + 1] = jas[j] + nas2[j];
    memset(kernels, 0, sizeof(double) * (size_t)nsigmas * nm1 * nm2);
    int mx1 = 0, mx2 = 0;
    for (int i = 0; i < nm1; i++) if (nas1[i] > mx1) mx1 = nas1[i];
    for (int j = 0; j < nm2; j++) if (nas2[j] > mx2) mx2 = nas2[j];
#pragma omp parallel
    {
        double *dsij = (double *)malloc(sizeof(double) * (size_t)(mx1 > 0 ? mx1 : 1) * (mx2 > 0 ? mx2 : 1));
#pragma omp for schedule(dynamic, 1)
        for (int a = 0; a < nm1; a++) {
            int ni = nas1[a];
            for (int b = 0; b < nm2; b++) {
                int nj = nas2[b];
                for (int i = 0; i < ni; i++)
                    for (int j = 0; j < nj; j++) {
                        const double *p = x1 + (size_t)(i + ias[a]) * D, *q = x2 + (size_t)(j + jas[b]) * D;
                        double s = 0.0;
                        for (int f = 0; f < D; f++) s += fabs(p[f] - q[f]);
                        dsij[(size_t)i * nj + j] = s;
                    }
                for (int k = 0; k < nsigmas; k++) {
                    double inv = -1.0 / sigmas[k], s = 0.0;
                    for (int j = 0; j < nj; j++)
                        for (int i = 0; i < ni; i++) s += exp(dsij[(size_t)i * nj + j] * inv);
                    kernels[((size_t)k * nm1 + a) * nm2 + b] = s;
                }
            }
        }
        free(dsij);
    }
    free(ias); free(jas);
}
