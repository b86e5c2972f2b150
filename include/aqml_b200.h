/*
 * aqml_b200 -- C ABI of the B200-native AQML hot path (SLATM/aSLATM generation and
 * global / local Gaussian + Laplacian kernel-matrix assembly).
 *
 * This header is the drop-in boundary.  Every entry point names the reference interface it
 * replaces (binghuang2018/aqml, paths relative to the reference tree).  The reference binds
 * its Fortran through f2py (coreml/setup.py:48-64); the f2py-level routines are mirrored
 * one-to-one below with plain pointers and sizes (no torch / numpy types).
 *
 * Conventions
 *   - all reals are IEEE double, all integers int32 (f2py: INTEGER*4), row-major ("C order")
 *     unless a parameter says column-major.  A row-major (n, D) matrix is bit-for-bit the
 *     Fortran x(D, n) the reference passes (kernels.py:35,64,234: "Transposed for Fortran").
 *   - functions WITHOUT a `_dev` suffix take HOST pointers, copy to the current CUDA device,
 *     run, copy back and return when the result is complete (like an f2py call).
 *   - functions WITH a `_dev` suffix take DEVICE pointers for the bulk arrays (stated per
 *     parameter), enqueue on `stream` (a cudaStream_t passed as void*; NULL = default stream)
 *     and return without waiting for the kernels unless stated.  Small metadata arrays
 *     (element numbers, atom counts, sigmas, many-body types) are always HOST pointers.
 *   - return value: 0 on success, negative AQML_E* on failure; aqml_last_error() gives the
 *     message (thread-local).  Nothing aborts the process (the Fortran `stop`s,
 *     fslatm.f90:52-57).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     AQML_ECUDA.
 */
#ifndef AQML_B200_H
#define AQML_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AQML_OK        0
#define AQML_EINVAL   -1   /* bad argument */
#define AQML_ECUDA    -2   /* CUDA runtime error (no device, launch failure, out of memory) */
#define AQML_ELIMIT   -3   /* outside an implementation limit */

#define AQML_KERNEL_GAUSSIAN  0
#define AQML_KERNEL_LAPLACIAN 1

int aqml_version(void);
const char *aqml_last_error(void);
/* number of visible CUDA devices (0 when none / no driver); never fails */
int aqml_device_count(void);

/* ------------------------------------------------------------------------------------------
 * Global kernels and distances
 * ---------------------------------------------------------------------------------------- */

/* fgaussian_kernel(a, na, b, nb, k, sigma)   coreml/cml/fkernels.f90:327-359
 *   (Python: kernels.gaussian_kernel, coreml/cml/kernels.py:40-67)
 * a (na, D), b (nb, D) row-major; k column-major (na, nb) caller-allocated, overwritten:
 * k[j + na*i] = exp(-0.5 * |a_j - b_i|^2 / sigma^2). */
int aqml_fgaussian_kernel(const double *a, int na, const double *b, int nb, int D, double *k, double sigma);

/* flaplacian_kernel(a, na, b, nb, k, sigma)  coreml/cml/fkernels.f90:361-387
 *   (Python: kernels.laplacian_kernel, kernels.py:11-38)   k = exp(-|a_j - b_i|_1 / sigma) */
int aqml_flaplacian_kernel(const double *a, int na, const double *b, int nb, int D, double *k, double sigma);

/* fl2_distance(A, B) -> D(n1, n2)            coreml/cml/fdistance.f90:25-44
 *   (Python: distance.l2_distance, distance.py:39-68)   Dm column-major (n1, n2). */
int aqml_fl2_distance(const double *A, const double *B, int n1, int n2, int nc, double *Dm);
/* fl1_distance                                coreml/cml/fdistance.f90:3-22 */
int aqml_fl1_distance(const double *A, const double *B, int n1, int n2, int nc, double *Dm);

/* flinear_kernel(a, na, b, nb, k)              coreml/cml/fkernels.f90:390-412  k[j + na*i] = a_j . b_i */
int aqml_flinear_kernel(const double *a, int na, const double *b, int nb, int D, double *k);
/* distance / Gram matrix on the device: p = 2 -> |a_i - b_j|_2, p = 1 -> |a_i - b_j|_1, p = 0 -> a_i . b_j.
 * a, b, out DEVICE; out row-major (na, nb).  Building block of the Matern / Sargan kernels
 * (fkernels.f90:414-518), whose scalar post-processing of d is not on the hot path. */
int aqml_distance_dev(int p, const double *a, int na, int64_t lda, const double *b, int nb, int64_t ldb, int D, double *out,
                      void *stream);

/* All sigmas in one pass (the reference re-runs the whole distance loop per sigma:
 * representation/xa.py:223, slatm_x.py:565).  kind = AQML_KERNEL_*.
 * a (na, lda>=D), b (nb, ldb>=D) row-major DEVICE; sigmas HOST (nsig); K DEVICE row-major
 * (nsig, na, nb): K[s][i][j] = k(a_i, b_j; sigma_s).  Row-block sharding over GPUs = calling
 * this with a pointer into a's row block.  center != 0 subtracts the column mean of b from
 * both operands before the Gaussian contraction (distance-preserving, conditions the
 * |x|^2+|y|^2-2xy expansion). */
int aqml_global_kernel_dev(int kind, const double *a, int na, int64_t lda, const double *b, int nb, int64_t ldb,
                           int D, const double *sigmas, int nsig, int center, double *K, void *stream);

/* Building blocks of the max-distance pass when the rows are spread over several GPUs (aqml_b200/sharded.py: the
 * reference's get_kernel_width, aqml.py:204-255, over rows that never meet on one device): column sums of n rows ADDED to
 * colsum (DEVICE, ncols), and the distance (p = 2: Euclidean, 1: Manhattan) of every row to one point c (DEVICE) -> r (DEVICE, n).
 * Asynchronous on `stream`. */
int aqml_rows_colsum_dev(const double *x, int64_t ld, int n, int ncols, double *colsum, void *stream);
int aqml_rows_dist_dev(int p, const double *x, int64_t ld, int n, const double *c, int ncols, double *r, void *stream);

/* max_{i,j} |a_i - b_j|_p (p = 2 or 1) without materialising the distance matrix: the
 * reduction get_kernel_width needs (algo/aqml.py:204-255: np.max(l2_distance(x_z, x_z))).
 * a, b DEVICE; result written to HOST *dmax (synchronises the stream). */
int aqml_max_distance_dev(int p, const double *a, int na, int64_t lda, const double *b, int nb, int64_t ldb, int D,
                          double *dmax, void *stream);

/* ------------------------------------------------------------------------------------------
 * Local (atomic) kernels
 * ---------------------------------------------------------------------------------------- */

/* calc_lk_gaussian(x1,x2,zs1,zs2,nas1,nas2,zmax,iab,sigmas,nm1,nm2,nsigmas) -> kernels
 *   coreml/cml/fkernels.f90:23-93   (Python: kernels.get_local_kernels_gaussian, kernels.py:193-235)
 * x1 (sum nas1, D), x2 (sum nas2, D) row-major; zs* element number per atom (1..zmax);
 * sigmas row-major (nsigmas, zmax), column Z-1; kernels row-major (nsigmas, nm1, nm2):
 * K[k][a][b] = sum_{i in a} sum_{j in b} exp(-0.5 d2_ij / sigmas[k][Z1_i - 1]^2) over pairs with
 * Z1_i == Z2_j (all pairs when iab != 0).  Only same-element pairs are contracted, and only on
 * the columns where that element's atoms are non-zero (detected from the data, exact). */
int aqml_calc_lk_gaussian(const double *x1, const double *x2, const int *zs1, const int *zs2, const int *nas1,
                          const int *nas2, int zmax, int iab, const double *sigmas, int nm1, int nm2, int nsigmas,
                          int D, double *kernels);
/* same; x1, x2, kernels DEVICE (ldx* = row pitch in doubles), everything else HOST */
int aqml_calc_lk_gaussian_dev(const double *x1, int64_t ldx1, const double *x2, int64_t ldx2, const int *zs1,
                              const int *zs2, const int *nas1, const int *nas2, int zmax, int iab,
                              const double *sigmas, int nm1, int nm2, int nsigmas, int D, int center,
                              double *kernels, void *stream);

/* calc_lk_laplacian(x1,x2,nas1,nas2,sigmas,nm1,nm2,nsigmas) -> kernels
 *   coreml/cml/fkernels.f90:97-183  (Python: kernels.get_local_kernels_laplacian, kernels.py:237-276)
 * no element matching, sigmas (nsigmas). */
int aqml_calc_lk_laplacian(const double *x1, const double *x2, const int *nas1, const int *nas2,
                           const double *sigmas, int nm1, int nm2, int nsigmas, int D, double *kernels);
int aqml_calc_lk_laplacian_dev(const double *x1, int64_t ldx1, const double *x2, int64_t ldx2, const int *nas1,
                               const int *nas2, const double *sigmas, int nm1, int nm2, int nsigmas, int D,
                               double *kernels, void *stream);

/* Per-element max distance, the kernel-width heuristic's reduction
 *   algo/aqml.py:204-255 get_kernel_width (cab=False): dso[z-1] = max_{i,j in z} |x_i - x_j|_p,
 *   1e-9 for absent elements.  x DEVICE (A, ldx), zs HOST (A), dso HOST (zmax). p = 2 | 1. */
int aqml_kernel_width_dev(int p, const double *x, int64_t ldx, const int *zs, int A, int D, int zmax, double *dso,
                          void *stream);

/* ------------------------------------------------------------------------------------------
 * SLATM / aSLATM
 * ---------------------------------------------------------------------------------------- */

/* Parameters of generate_slatm (representation/core.py:382-384) after the Python wrappers
 * get_sbop / get_sbot (representation/slatm.py:22-130) resolved the grids:
 *   nx2 = int((rcut-0.1)/dgrid2)+1, nx3 = int((pi+40deg)/dgrid3)+1, coeff = 1/sqrt(2 sigma^2 pi). */
typedef struct {
    int nmb;             /* number of many-body types */
    const int *mbtypes;  /* HOST (nmb, 3), -1 padded: [z], [z1,z2], [z1,z2,z3]; column order of X */
    int izeff;           /* use the Z* table, atomdb.f90:90-116 */
    int nx2, nx3;
    double sigma2, sigma3, dgrid2, dgrid3, coeff2, coeff3;
    double rcut, rpower;
} aqml_slatm_params;

/* feature length D = sum over types of 1 | nx2 | nx3 */
int aqml_slatm_dim(const aqml_slatm_params *p);

/* calc_sbop / calc_sbop_local   coreml/cml/representation/fslatm.f90:413-527 / :531-655
 * calc_sbot / calc_sbot_local   coreml/cml/representation/fslatm.f90:4-202   / :206-409
 * coords row-major (na,3) [the f2py wrapper copies to Fortran order itself], zs (na),
 * ia = 0-based centre atom (ia_python) or -1 for the global routine; ys (nx) out.
 * cg is not a parameter: the reference always passes all-ones (slatm.py:53,107). */
int aqml_calc_sbop(const double *coords, const int *zs, int na, int ia, int izeff, int z1, int z2, double rcut,
                   int nx, double dgrid, double sigma, double coeff, double rpower, double *ys);
int aqml_calc_sbot(const double *coords, const int *zs, int na, int ia, int izeff, int z1, int z2, int z3,
                   double rcut, int nx, double dgrid, double sigma, double coeff, double *ys);

/* generate_slatm for a batch of molecules (representation/core.py:382-556, one call instead of
 * natoms x nmbtypes f2py crossings per molecule).  coords (A,3), zs (A), nas (nmol).
 * local != 0: out (A, D) one aSLATM row per atom; else out (nmol, D) one SLATM row per molecule. */
int aqml_generate_slatm_batch(const double *coords, const int *zs, const int *nas, int nmol,
                              const aqml_slatm_params *p, int local, double *out);
/* coords DEVICE (A,3), zs_dev DEVICE (A) int32, out DEVICE with row pitch ldo >= D; nas HOST. */
int aqml_generate_slatm_batch_dev(const double *coords, const int *zs_dev, const int *nas, int nmol,
                                  const aqml_slatm_params *p, int local, double *out, int64_t ldo, void *stream);

/* ------------------------------------------------------------------------------------------
 * Fused path: molecules -> element-packed aSLATM -> local Gaussian kernel, all device resident.
 * No reference equivalent (the reference materialises dense x1/x2 on the host between
 * generate_slatm and get_local_kernels_gaussian, algo/aqml.py:711-789).
 * ---------------------------------------------------------------------------------------- */
typedef struct aqml_rep aqml_rep; /* aSLATM rows grouped by element, columns restricted to the element's support */

/* Contraction engines of the Gaussian kernels (calc_lk_gaussian, fgaussian_kernel and the fused path below).  Default:
 * FP64-accurate a.b from exact int8 digit-plane products on the tcgen05 tensor cores (six 8-bit digit planes per row, 22
 * exact int32 plane products in TMEM issued as 8 stacked-N UTCIMMA per 32 columns, i8pair.cuh; AQML_I8_PRODUCTS=26 adds
 * the rest of the seventh product level); the representation then also holds 6 bytes of digit planes per packed element.
 * Environment AQML_PAIR_ENGINE=dmma selects the FP64 DMMA engine (no digit planes are built).  More than 8 sigmas or rows
 * longer than 21840 columns use the FP64 engine automatically.  Both agree with the reference to <= 1e-10 per element. */

/* coords: DEVICE if coords_on_device else HOST; zs, nas HOST.  Asynchronous on `stream`. */
int aqml_rep_create(const double *coords, int coords_on_device, const int *zs, const int *nas, int nmol,
                    const aqml_slatm_params *p, void *stream, aqml_rep **out);
/* Same with flags.  Bit 0 (AQML_REP_LAYOUT_ONLY): allocate the representation's storage exactly as a computing call
 * would but launch nothing (coords may be NULL) -- the contents are then received from the GPU that computed them.
 * The storage of a representation is ONE device allocation whose layout depends only on (zs, nas, p), so
 * multi-GPU code generates 1/N of the molecule blocks per rank and ships each block once over NCCL
 * (aqml_b200/parallel.py::share_reps) instead of recomputing it N times. */
#define AQML_REP_LAYOUT_ONLY 1
/* Bit 1 (AQML_REP_NO_ROWS, only with LAYOUT_ONLY): do not allocate the packed FP64 rows either -- the receiver of a block only
 * contracts it on the int8 tensor engine, which reads the digit planes; the width pass runs distributed on the owners' rows
 * (aqml_b200/sharded.py) and the FP64 engine refuses such a representation. */
#define AQML_REP_NO_ROWS 2
int aqml_rep_create_ex(const double *coords, int coords_on_device, const int *zs, const int *nas, int nmol,
                       const aqml_slatm_params *p, int flags, void *stream, aqml_rep **out);
/* the representation's device allocations: which = 0 the pair kernel's operands (|row|^2, molecule index, k-tile masks, digit
 * planes, exponents), which = 1 the packed FP64 rows (NULL / 0 with AQML_REP_NO_ROWS); pointer (DEVICE) and size in bytes */
int aqml_rep_slab(const aqml_rep *r, int which, void **ptr, int64_t *bytes);
/* element groups of a representation: count; label (Z), valid rows n (the first n packed rows), row pitch ld, DEVICE rows */
int aqml_rep_elements(const aqml_rep *r);
int aqml_rep_element(const aqml_rep *r, int e, int *label, int *n, int *ld, const double **x);
/* frees on the creation stream after every kernel that read the representation on another stream has finished */
void aqml_rep_destroy(aqml_rep *r);
int64_t aqml_rep_bytes(const aqml_rep *r);
/* K DEVICE row-major (nsigmas, r1.nmol, r2.nmol), element matched; sigmas HOST (nsigmas, zmax).
 * row0/nrows select a block of r1's molecules (row-block sharding); K has nrows rows. */
int aqml_rep_local_gaussian(const aqml_rep *r1, const aqml_rep *r2, const double *sigmas, int nsigmas, int zmax,
                            int row0, int nrows, double *K, void *stream);
/* K(r, r): symmetric, DEVICE row-major (nsigmas, r.nmol, r.nmol).  Only atom pairs above the diagonal of the
 * packed order are contracted (half the work of aqml_rep_local_gaussian(r, r, ...)), then K is mirrored and
 * the i == j terms (exactly 1 each, as in the reference) are added. */
int aqml_rep_local_gaussian_sym(const aqml_rep *r, const double *sigmas, int nsigmas, int zmax, double *K, void *stream);
/* per-element max L2 distance over r1 x r2 atoms (get_kernel_width on packed rows); dso HOST (zmax) */
int aqml_rep_kernel_width(const aqml_rep *r1, const aqml_rep *r2, int zmax, double *dso, void *stream);
/* same over the union of several representations on each side (a data set packed in blocks): the maximum over
 * ALL atom pairs of reps1[*] x reps2[*].  The pass is pruned by the triangle inequality (distances to a common
 * reference point + one attained pair distance as lower bound), so only the rows that can take part in the
 * maximum are contracted -- the reference forms the whole distance matrix (aqml.py:220-237). */
int aqml_rep_kernel_width_multi(const aqml_rep *const *reps1, int n1, const aqml_rep *const *reps2, int n2, int zmax, double *dso,
                                void *stream);

/* ---- pair-wise / bond aSLATM ("next" row 8f; coreml/cml/representation/fslatm_pairwise.f90) --------------------------
 * The file is outside the reference's build and cannot run as shipped (`na` unset in calc_sbop_local, :394); semantics =
 * its statements with na := size(zs) (oracle/np_pairwise.py, PARITY UNPINNED).
 * aqml_pairwise_dims: grid sizes nsx[3] = {two-body atoms, two-body bonds, three-body} and the row lengths
 *   n = n1 + n2 nsx[0] + n3 nsx[2] (atom rows), nu = n1 + n2 nsx[1] + n3 nsx[2] (bond rows)      (calc_mk1 :781-788, xb.py:110-124)
 * aqml_pairwise_local_spectrum_dev: replaces calc_all_local_spectrum / calc_local_spectrum (:525-673) with ia0 = ib0 = 0:
 *   coords (A,3), zs (A), nas (nmol), nbrs (A, nbmax; molecule-local indices, -1 padded; NULL with nbmax = 0: no bond rows),
 *   mbs1 (n1), mbs2 (n2,2), mbs3 (n3,3), rscut = {racut, rbcut}: all HOST.  ys DEVICE (A, n), ys2 DEVICE (A, nbmax, nu).
 *   Synchronises `stream`. */
int aqml_pairwise_dims(const double *rscut, double dgrid, int n1, int n2, int n3, int *nsx, int *n, int *nu);
int aqml_pairwise_local_spectrum_dev(const double *coords, const int *zs, const int *nas, int nmol, const int *nbrs, int nbmax,
                                     const int *mbs1, int n1, const int *mbs2, int n2, const int *mbs3, int n3, const double *rscut,
                                     double dgrid, double sigma, double w2, double rpower2, double w3, double rpower3, double *ys,
                                     double *ys2, void *stream);

/* counters for bench.py: number of device kernels this library launched since load */
int64_t aqml_launch_count(void);
/* k-tiles (16 columns x one 128x64 tile) the pair kernels executed vs. a dense sweep would have, summed
 * over all tiles since the last reset (all-zero k-tiles are skipped exactly);
 * synchronises the device. */
int aqml_pair_stats(int64_t *executed_ktiles, int64_t *dense_ktiles, int reset);
/* Peak rate of the int8 tensor pipe (tcgen05.mma kind::i8, M 128, N 256, K 32, operands resident in shared
 * memory), measured now on the current device: int8 multiply-add operations (2 per product) per second.
 * bench.py divides the pair kernel's executed int8 operations by this for `roofline.frac`.  Blocks ~10 ms. */
int aqml_measure_i8_peak(double *ops_per_s, void *stream);
/* CTA tile of the pair kernels: bm x bn atom pairs, bk columns per k-tile (2*bm*bn*bk flops each) */
void aqml_tile_shape(int *bm, int *bn, int *bk);

#ifdef __cplusplus
}
#endif
#endif /* AQML_B200_H */
