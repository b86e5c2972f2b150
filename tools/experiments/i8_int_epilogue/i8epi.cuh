// FP64-free arithmetic for the epilogue of the int8 pair kernel (i8pair.cuh).
//
// Why: on B200 an FP64 instruction that is issued while UTCIMMA stream through the tensor pipe waits ~800 clk
// (tools/micro/i8_merged.cu, "contention": a dependent DFMA step 59 -> 750..900 clk, FFMA / IMAD unaffected), so the
// round-1 epilogue -- ~130 dependent FP64 steps per tile (exp, squaring chain, running sums, shuffle reductions) -- took
// as long as the tile's MMAs.  Everything between the exact integer accumulators and the RED into K is therefore done
// with 64-bit integers and FP32 pairs ("double-single"), which overlap with the tensor pipe:
//
//   d^2        D = NA + NB - (V >> s)            64-bit fixed point, unit 2^(EN-58), EN: per-tile exponent of the largest
//                                                |a|^2 + |b|^2; V = exact 2^64 a.b 2^-(ea+eb) from the accumulators.  The
//                                                cancellation is exact to 2^-58 of the norms (FP64: 2^-53)
//   -y log2 e  Z = (D x mant) >> (64 + t)        y = d^2 / (2 sigma^2); 44 fractional bits; one 64 x 64 -> 128 multiply
//   exp        2^-Z = 2^-k T1[f1] T2[f2] T3[f3] (1 - r),   three 512-entry tables of 2^-(j / 512^n) as FP32 pairs, r < 2^-27.5
//              result: mantissa pair (h, l) in (1/8, 1] and the integer k       (relative error ~2^-46)
//   sigmas     chain: square the pair, double k, renormalise                    (~1 bit per squaring)
//   sums       per molecule pair: terms aligned to the smallest k of the segment, converted to 64-bit integers with 53
//              fractional bits, added exactly; lanes aligned to the smallest k of the run, shuffled and added exactly; the
//              head lane builds the double from the integer (shift, round, exponent) -- no FP64 instruction at all
//
// Host-callable (HD) so that tests/ can check the arithmetic against long double on the CPU.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define I8E_HD __host__ __device__ __forceinline__
#else
#define I8E_HD inline
#endif

namespace aqml {
namespace i8e {

constexpr int KBIG = 1 << 26;            // "this pair contributes nothing" (exp underflow, padding, masked)
constexpr int TAB = 512;                 // entries per table
constexpr int FRAC = 44;                 // fractional bits of Z
constexpr int UNIT = 58;                 // d^2 fixed point: unit 2^(EN - UNIT)
constexpr int SUMBITS = 53;              // fractional bits of the integer sums

struct F2 { float h, l; };               // value h + l, |l| <= ulp(h) / 2

I8E_HD F2 mul22(F2 a, F2 b) {
    const float ph = a.h * b.h;
    const float pe = fmaf(a.h, b.h, -ph);
    const float pl = fmaf(a.h, b.l, fmaf(a.l, b.h, pe));
    const float s = ph + pl;
    return F2{s, pl - (s - ph)};
}
I8E_HD F2 sqr2(F2 a) {
    const float ph = a.h * a.h;
    const float pe = fmaf(a.h, a.h, -ph);
    const float pl = fmaf(a.h + a.h, a.l, pe);
    const float s = ph + pl;
    return F2{s, pl - (s - ph)};
}

I8E_HD uint64_t umulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
I8E_HD int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}
I8E_HD long long f2ll(float v) {
#if defined(__CUDA_ARCH__)
    return __float2ll_rn(v);
#else
    return llrintf(v);
#endif
}
I8E_HD float pow2f(int e) {  // 2^e, -126 <= e <= 127
    union { unsigned u; float f; } c;
    c.u = (unsigned)(e + 127) << 23;
    return c.f;
}

// d^2 in units of 2^(EN - UNIT): nab = NA + NB (fixed point), V = hi 2^32 + lo, s = EN + 5 - ea - eb in [0, 127]
I8E_HD uint64_t dfix(long long hi, long long lo, int s, long long nab) {
    const __int128 V = ((__int128)hi << 32) + (__int128)lo;
    const long long d = nab - (long long)(V >> s);
    return d < 0 ? 0ull : (uint64_t)d;   // rounding noise of a.b ~ |a||b|: clamp like the reference's max(., 0) never needs to
}

// Z = d^2 / (2 sigma^2) log2(e) with FRAC fractional bits; mant = 2^64 frac(c), t = 14 - EN - exponent(c).  Saturates.
I8E_HD uint64_t zfix(uint64_t D, uint64_t mant, int t) {
    const uint64_t H = umulhi64(D, mant);
    if (t >= 0) return t > 63 ? 0ull : H >> t;
    const int u = -t;
    return (u > 40 || (H >> (63 - u))) ? ~0ull : H << u;
}

// 2^(-Z 2^-FRAC) = (m.h + m.l) 2^-k; tab = T1 | T2 | T3 (3 x TAB pairs).  Straight-line code (no branch): the 16 pairs
// of a thread are independent and the compiler interleaves them -- with a branch per pair every latency is exposed.
template <class TabPtr>
I8E_HD void exp2neg(uint64_t Z, TabPtr tab, F2 &m, int &k) {
    const unsigned zh = (unsigned)(Z >> 32), zl = (unsigned)Z;
    const unsigned f1 = (zh >> 3) & 511u, f2 = ((zh << 6) | (zl >> 26)) & 511u, f3 = (zl >> 17) & 511u;
    const float r = (float)(zl & 0x1ffffu) * 3.9400921295e-14f;  // ln 2 * 2^-44
    const F2 a = tab[f1], b = tab[TAB + f2], c = tab[2 * TAB + f3];
    const F2 p = mul22(mul22(a, b), c);
    const float l = p.l - p.h * r;       // (1 - r): r < 2^-27.5, r^2 / 2 < 2^-56
    const float s = p.h + l;
    m = F2{s, l - (s - p.h)};
    k = (Z >= ((uint64_t)1100 << FRAC)) ? KBIG : (int)(zh >> (FRAC - 32));
}

// (m, k) <- (m, k)^2, mantissa back into [1/2, 1)
I8E_HD void square(F2 &m, int &k) {
    if (k >= KBIG / 2) { k = KBIG; return; }
    F2 q = sqr2(m);
    union { float f; unsigned u; } c;
    c.f = q.h;
    const int e = (int)((c.u >> 23) & 255u) - 126;      // q.h = x 2^e, x in [1/2, 1); e in [-7, 0] here
    const float sc = pow2f(-e);
    m = F2{q.h * sc, q.l * sc};
    k = 2 * k - e;
}

// term (m, k), 0 <= m <= 1, as an integer with SUMBITS fractional bits relative to 2^-kref (k >= kref); straight-line:
// a term more than 2^-63 below the reference rounds to 0 by itself
I8E_HD long long term_fix(F2 m, int k, int kref) {
    const int d = (k - kref) > 63 ? 63 : (k - kref);
    const float sc = pow2f(SUMBITS - d);
    return f2ll(m.h * sc) + f2ll(m.l * sc);
}

// I 2^e2 as a double, round to nearest (ties away), results below 2^-1022 flushed to zero, no FP64 instruction
I8E_HD double fix_to_double(uint64_t I, int e2) {
    if (I == 0) return 0.0;
    int lz = clz64(I);
    uint64_t m = I << lz;                 // bit 63 set
    uint64_t mr = m + (1ull << 10);
    int E = 63 - lz + e2;                 // value = 1.xxx 2^E
    if (mr < m) { mr = 1ull << 63; E += 1; }
    const long long field = (long long)E + 1023;
    uint64_t bits;
    if (field <= 0) bits = 0;
    else if (field >= 2047) bits = 0x7ff0000000000000ull;
    else bits = ((uint64_t)field << 52) | ((mr >> 11) & 0xfffffffffffffull);
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)bits);
#else
    union { uint64_t u; double d; } c;
    c.u = bits;
    return c.d;
#endif
}

// tables (host): T_n[j] = 2^-(j / 512^n), n = 1..3, as FP32 pairs
inline void fill_tables(float *tab /* 3 * TAB * 2 floats */) {
    for (int n = 0; n < 3; n++)
        for (int j = 0; j < TAB; j++) {
            const double v = exp2(-(double)j / pow(512.0, n + 1));
            const float h = (float)v;
            tab[2 * (n * TAB + j)] = h;
            tab[2 * (n * TAB + j) + 1] = (float)(v - (double)h);
        }
}

// per (sigma, element) constant c = -isig log2(e) = frac 2^fe: mant = 2^64 frac, fe
inline void sigma_const(double isig, uint64_t &mant, int &fe) {
    const double c = -isig * 1.4426950408889634;
    if (!(c > 0.0) || !std::isfinite(c)) { mant = 0; fe = 0; return; }
    const double fm = frexp(c, &fe);
    mant = (uint64_t)ldexp(fm, 63) << 1;  // fm 2^64 (fm 2^63 fits a uint64 exactly)
}

}  // namespace i8e
}  // namespace aqml
