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

// 8-bit digits of t, |t| < 0.498: q = floor(256 t + 128/255) in [-128, 127], remainder in [-128/255, 127/255) -- the one
// interval that maps onto itself, so every digit fits an int8 (the clamp only acts when the rounding of u + c crosses
// the interval's end; the remainder stays exact and the following digits absorb it)
__device__ __forceinline__ int i8_digit(double &t) {
    const double u = t * 256.0;                                       // exact
    const double q = fmin(fmax(floor(u + 128.0 / 255.0), -128.0), 127.0);
    t = u - q;                                                        // exact
    return (int)q;
}

// 16 consecutive columns of one row (already scaled by 2^-e) -> 6 planes x 16 B at dst, dst + 2048, ...
__device__ __forceinline__ void i8_split16(double (&t)[16], int8_t *dst) {
#pragma unroll
    for (int c = 0; c < 16; c++)
        if (!(fabs(t[c]) <= 0.5)) t[c] = 0.0;          // NaN / Inf / > 1e301: see i8_row_exponent
#pragma unroll
    for (int s = 0; s < I8_S; s++) {
        unsigned w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int c = 0; c < 16; c++) w[c >> 2] |= ((unsigned)i8_digit(t[c]) & 0xffu) << ((c & 3) * 8);
        *reinterpret_cast<uint4 *>(dst + s * 2048) = make_uint4(w[0], w[1], w[2], w[3]);
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
