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
void run_pairs_public(int mode, int epi, const std::vector<PairGroup> &groups, PairParams P, Workspace &ws);

}  // namespace aqml
