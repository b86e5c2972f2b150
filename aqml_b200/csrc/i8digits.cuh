// Digit planes of the int8 tensor engine (i8pair.cuh): the rule that turns an FP64 row into six signed 8-bit digits of a
// per-row fixed point number, shared by the stand-alone split kernels and by the aSLATM kernel, which writes the planes of a
// row right after the row itself (slatm_units.cuh, phase 3).
#pragma once

namespace aqml {

constexpr int I8_S = 6;                          // digit planes (8 bits each)

namespace i8 {
__device__ __forceinline__ double pow2(int e) {  // 2^e, clamped to the normal range
    e = max(-1022, min(1023, e));
    return __hiloint2double((e + 1023) << 20, 0);
}
}  // namespace i8

// exponent of a row with largest finite |x| = m:  |x| 2^-e < 0.498 (the first digit floor(256 t + 128/255) must not exceed
// 127).  Non-finite or absurdly large (> 1e301) values need no special path: such a row has |row|^2 = NaN / Inf, which reaches
// K through d^2 = |a|^2 + |b|^2 - 2 a.b exactly as in the reference (NaN stays NaN, Inf gives exp(-Inf) = 0); whatever digits
// the row gets are irrelevant.
__device__ __forceinline__ int i8_row_exponent(double m) {
    int e = 0;
    if (m > 0.0) {
        const double f = frexp(m, &e);  // m = f 2^e, f in [1/2, 1)
        e += (f < 0.996) ? 1 : 2;
    }
    return max(min(e, 1000), -990);  // tinier rows keep |x| 2^-e < 0.498 and only lose (irrelevant) digits
}

// 8-bit digits of t, |t| < 0.498.  Rule: q = floor(256 t + 128/255) in [-128, 127], remainder in [-128/255, 127/255) -- the
// one interval that maps onto itself, so every digit fits an int8 -- applied six times.  Done on integers: I = rint(t 2^48)
// (ONE FP64 multiply + conversion per element), J = I + 0x808080808080 (adds 128 to every base-256 digit, carries included:
// 0 < J < 2^48 because |t| < 0.498 < 128/255), digit_k = byte_k(J) - 128 = byte_k(J) ^ 0x80.  The six bytes of 16 columns
// are regrouped into the six planes with byte permutes.  ~150 integer instructions per 16 columns instead of ~600 FP64 ones
// (floor / min / max / subtract per digit).
//
// 16 consecutive columns of one row, scale = 2^(48 - e)  ->  6 planes x 16 B at dst, dst + 2048, ...
__device__ __forceinline__ void i8_split16(const double (&x)[16], double scale, int8_t *dst) {
    unsigned lo[16], hi[16];   // J = hi 2^32 + lo, hi < 2^16
#pragma unroll
    for (int c = 0; c < 16; c++) {
        long long I = __double2ll_rn(x[c] * scale);
        if (I >= (1ll << 47) || I <= -(1ll << 47)) I = 0;          // NaN / Inf / > 1e301 (see i8_row_exponent): digits irrelevant
        const unsigned long long J = (unsigned long long)(I + 0x808080808080ll);
        lo[c] = (unsigned)J;
        hi[c] = (unsigned)(J >> 32);
    }
#pragma unroll
    for (int p = 0; p < I8_S; p++) {   // plane p = digit of weight 256^-(p+1) = byte 5 - p of J
        unsigned w[4];
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const unsigned *src = (p < 2) ? hi : lo;
            const int b = (p < 2) ? (1 - p) : (5 - p);             // byte inside the 32-bit word
            const unsigned a01 = __byte_perm(src[4 * g], src[4 * g + 1], b | ((4 + b) << 4));        // bytes of columns 0, 1
            const unsigned a23 = __byte_perm(src[4 * g + 2], src[4 * g + 3], b | ((4 + b) << 4));    // bytes of columns 2, 3
            w[g] = __byte_perm(a01, a23, 0x5410) ^ 0x80808080u;
        }
        *reinterpret_cast<uint4 *>(dst + p * 2048) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// largest finite |x| of every row -> exponent
__global__ void __launch_bounds__(256) i8_rowexp_kernel(const double *x, long long ld, int nrows, int *expo) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= nrows) return;
    const double *xr = x + (long long)row * ld;
    double m = 0.0;
    for (int k = lane; k < ld; k += 32) {
        const double v = fabs(xr[k]);
        if (v <= 1.7976931348623157e308) m = fmax(m, v);
    }
    m = warp_max(m);
    if (lane == 0) expo[row] = i8_row_exponent(m);
}

}  // namespace aqml
