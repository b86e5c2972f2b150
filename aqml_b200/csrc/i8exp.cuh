// exp and sums of the int8 engine's epilogue WITHOUT FP64 instructions (i8pair.cuh).
//
// Why: while UTCIMMA stream through the tensor pipe an FP64 instruction takes 750-900 clk instead of 59 (integer and FP32
// instructions are not affected: profiles/r02_micro_contention.txt), and the FP64 epilogue (exp by table + polynomial, squaring
// chain, sums, shuffle trees: ~24 FP64 instructions per atom pair) took longer than the tile's MMAs.  Here everything after
// T = d^2 / (2 sigma^2) * log2(e) (ONE FP64 multiply per pair; d^2 itself stays the FP64 expression of the reference,
// fkernels.f90:71-84) is 32/64-bit integer arithmetic:
//   * T >= 0 as Q16.48 fixed point F (bit pattern of the double, two shifts); the sigma ladder is a left shift of F -- no
//     squaring chain, so the error does not grow along the ladder;
//   * 2^-T = 2^-n * tab[j] * 2^-r:  n = integer part, j = the next 12 bits (table of 2^(-j/4096) as Q53 in shared memory),
//     r < 2^-12 the remaining 36 bits, 2^-r = 1 - y + y^2/2, y = r ln 2 < 1.7e-4, in Q44 with 32 x 32 -> high products
//     (omitted y^3/6 <= 8e-13 relative);
//   * a term is (M, n): value = M 2^-(53+n), M in [2^52, 2^53] -- M + ((1021 - n) << 52) IS the bit pattern of the double;
//   * the terms of one (thread, molecule-pair run) are added as 64-bit integers aligned to the smallest n seen so far
//     (truncation <= 2^-52 of the largest term per addition); the lanes of one molecule agree on their smallest n, add the
//     aligned integers by a segmented shuffle reduction, and the total becomes a double with one conversion.
// Relative error of a term: ~2^-44 (r truncated to 32 bits) + 8e-13 (cubic) + 2^-53 T (the FP64 multiply) -- measured <= 1e-12
// against exp() in extended precision (tools/micro/i8exp_check.cpp); stored terms and sums below 2^-1022 are flushed to zero
// (the FP64 epilogue flushed below exp(-708)).  T = +inf (pairs that do not count) and NaN give 0.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AQML_HD __host__ __device__ __forceinline__
#else
#define AQML_HD inline
#endif

namespace aqml {
namespace i8x {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int TAB_BITS = 12, TAB_N = 1 << TAB_BITS;
constexpr u32 LN2_Q32 = 2977044472u;  // round(ln 2 * 2^32)
constexpr int N_ZERO = 1022;               // 2^-T < 2^-1022: flushed to zero

AQML_HD u32 mulhi(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (u32)(((u64)a * b) >> 32);
#endif
}

AQML_HD u32 funnel_l(u32 lo, u32 hi, int k) {  // high word of (hi:lo) << k, k = 0 .. 31
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, k);
#else
    return k ? ((hi << k) | (lo >> (32 - k))) : hi;
#endif
}

// T >= 0 given as the two words of a double -> Q16.48.  No branches, no selects: T >= 2^16 (also +inf, NaN) leaves the top bit
// set (integer part >= 32768 -> the term is zero), T < 2^-15 becomes 0 or 1 (an error of 2^-48 in T)
AQML_HD u64 fix_from_bits(u32 hi, u32 lo) {
    int sh = 1038 - (int)(hi >> 20);                       // F = (m << 11) >> sh,  m = 53-bit significand
    sh = sh < 0 ? 0 : (sh > 63 ? 63 : sh);
    const u64 m11 = ((u64)(funnel_l(lo, hi, 11) | 0x80000000u) << 32) | (u64)(lo << 11);
    return m11 >> sh;
}

// F << k for the next sigma of the ladder, k = 0 .. 8: clamping the high word first keeps "integer part >= 1022" (a zero term)
// true instead of wrapping around; a term that is not zero has F < 1022 * 2^48 and cannot overflow for k <= 6
AQML_HD u64 fix_shl(u64 F, int k) {
    u32 hi = (u32)(F >> 32);
    const u32 lo = (u32)F, cap = 0xffffffffu >> k;
    hi = hi < cap ? hi : cap;
    return ((u64)funnel_l(lo, hi, k) << 32) | (u64)(lo << k);
}

// clamp(a, 0, 63)
AQML_HD int clamp63(int a) {
#if defined(__CUDA_ARCH__)
    return __vimin_s32_relu(a, 63);
#else
    return a < 0 ? 0 : (a > 63 ? 63 : a);
#endif
}

// one term 2^-T, T = F 2^-48:  value = M 2^-(53 + n), M in [2^52 - few, 2^53]; n >= N_ZERO: below the normal doubles.
// tab[j] = 2^(-j/4096) 2^53 (Q53, in (2^52, 2^53]);  M = tab[j] (1 - E 2^-44) with the product's low 22 bits dropped (2^-43).
// Two halves so that a caller can issue the table look-up of the next term before the arithmetic of this one.
struct Lookup { u64 t; u32 R; int n; };
AQML_HD Lookup exp2neg_lookup(u64 F, const u64 *tab) {
    const u32 hi = (u32)(F >> 32), lo = (u32)F;
    Lookup L;
    L.n = (int)(hi >> 16);
    L.R = funnel_l(lo, hi, 28);                           // r 2^44,  r < 2^-12 (its top 32 bits)
    L.t = tab[(hi >> 4) & (TAB_N - 1)];
    return L;
}
AQML_HD u64 exp2neg_finish(const Lookup &L) {
    const u32 Y = mulhi(L.R, LN2_Q32);                    // y 2^44,  y = r ln 2
    const u32 E = Y - (mulhi(Y, Y) >> 13);                // (y - y^2 / 2) 2^44
    const u32 p = mulhi(funnel_l((u32)L.t, (u32)(L.t >> 32), 11), E);   // top 32 bits of the table value times E, >> 32
    return L.t - ((u64)p << 9);
}
AQML_HD void exp2neg(u64 F, const u64 *tab, u64 &M, int &n) {
    const Lookup L = exp2neg_lookup(F, tab);
    n = L.n;
    M = exp2neg_finish(L);
}

// bit pattern of the double M 2^-(53 + n)  (M = 0 -> 0.0)
AQML_HD u64 term_bits(u64 M, int n) { return n < N_ZERO ? M + ((u64)(1021 - n) << 52) : 0ull; }

// acc 2^-(53 + nacc) += M 2^-(53 + n), keeping nacc = the smallest n so far (start: acc = 0, nacc = 4095)
AQML_HD void add_term(u64 &acc, int &nacc, u64 M, int n) {
    const int sa = clamp63(nacc - n), sm = clamp63(n - nacc);
    acc = (acc >> sa) + (M >> sm);
    nacc = n < nacc ? n : nacc;
}

}  // namespace i8x
}  // namespace aqml
