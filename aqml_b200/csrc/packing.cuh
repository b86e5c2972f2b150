// Declarations shared between kernels.cu and slatm.cu.
#pragma once
#include "common.cuh"
#include "pairtile.cuh"

namespace aqml {

struct PackJob {
    const double *src;   // (rows, lds)
    long long lds;
    const int *rowmap;   // packed row -> source row, -1 = zero row; null -> identity for r < nvalid
    const int *colmap;   // packed col -> source col; null -> identity
    const double *mean;  // per packed col, subtracted; may be null
    double *dst;         // (packed rows, ldd)
    long long ldd;
    double *norm;        // |row|^2 per packed row; may be null
    int ncols;           // valid packed columns (< = ldd)
    int nvalid;          // valid rows when rowmap is null
    unsigned *mask;      // [packed rows / 64][maskw] non-zero k-tile bits (pre-zeroed), or null
    int maskw;
    const double *const *srcrows;  // if set: source row pointer per packed row (null = zero row); src/rowmap unused
};

struct RowSeg {  // n packed rows at x (one operand may consist of several allocations)
    const double *x;
    int n;
};

// words of k-tile mask per 64-row block for a row pitch of ld columns
inline int mask_words(long long ld) { return (int)((ld / BK + 31) / 32); }

void launch_pack(const PackJob &J, int nrows_packed, cudaStream_t st);
void launch_colstats(const double *x, long long ldx, const int *rows, int nrows, int D, double *colsum, unsigned *colnz,
                     cudaStream_t st);
void build_tiles(const std::vector<int> &rows, const std::vector<int> &mol_of_row, std::vector<int> &rowmap,
                 std::vector<int> &seg);
std::vector<int> class_sorted_molecules(const int *zs, const int *nas, int nm);
void plan_sigma_chain(PairParams &P, const std::vector<double> &isig);
unsigned long long *pair_stats_buffer();

// int8 digit-plane engine (i8pair.cuh, compiled in slatm.cu) for element-matched local Gaussian groups given as packed
// FP64 operands.  Slices both sides into temporary digit planes, launches the tcgen05 kernel and returns true; returns
// false without launching anything when the engine is disabled (AQML_PAIR_ENGINE=dmma) or more than 8 sigmas are asked
// for.  NaN / Inf rows need no fallback: they reach K through |row|^2 as in the reference.
struct I8Side {
    const double *x, *norm;   // packed rows (pitch ld, multiple of 16), |row|^2
    const int *seg;           // molecule per packed row, -1 = padding
    const unsigned *mask;     // k-tile masks per 64-row block, or null
    int maskw, nr;            // mask words per block, packed rows (multiple of 64)
    long long ld;
};
struct I8Pair {
    I8Side a, b;
    int zcol;
};
bool i8_local_pairs(const std::vector<I8Pair> &pairs, const PairParams &P, Workspace &ws);
// global Gaussian kernel K[s][i][j] = exp(d2_ij * isig[s]) for packed operands a (na valid rows) and b (nb valid rows):
// seg is ignored (identity).  Same fallback rules as above.
bool i8_global_pair(const I8Side &a, int na, const I8Side &b, int nb, const PairParams &P, Workspace &ws);
void run_pairs_public(int mode, int epi, const std::vector<PairGroup> &groups, PairParams P, Workspace &ws);
double pruned_max_public(int p, const double *A, int nA, const double *B, int nB, long long ld, bool same, Workspace &ws);
double pruned_max_segs_public(int p, const std::vector<RowSeg> &SA, const std::vector<RowSeg> &SB, long long ld, bool same,
                              Workspace &ws);

}  // namespace aqml
