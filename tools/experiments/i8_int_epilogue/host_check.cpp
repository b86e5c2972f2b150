#include "i8epi.cuh"
#include <cstdio>
#include <random>
#include <vector>
#include <algorithm>
using namespace aqml::i8e;
int main() {
    std::vector<float> tabf(3 * TAB * 2);
    fill_tables(tabf.data());
    const F2 *tab = reinterpret_cast<const F2 *>(tabf.data());
    std::mt19937_64 rng(1);
    // 1. exp2neg + squaring chain
    double worst = 0, worst6 = 0;
    for (int it = 0; it < 2000000; it++) {
        uint64_t Z = rng() % ((uint64_t)(it % 3 == 0 ? 1000 : 40) << FRAC);
        F2 m; int k;
        exp2neg(Z, tab, m, k);
        long double ex = exp2l(-(long double)Z / (long double)(1ull << FRAC));
        long double got = ((long double)m.h + (long double)m.l) * exp2l(-(long double)k);
        worst = std::max(worst, (double)fabsl(got / ex - 1));
        if (it % 7 == 0 && Z < ((uint64_t)15 << FRAC)) {
            for (int q = 0; q < 6; q++) square(m, k);
            long double ex6 = exp2l(-(long double)Z * 64 / (long double)(1ull << FRAC));
            long double g6 = ((long double)m.h + (long double)m.l) * exp2l(-(long double)k);
            worst6 = std::max(worst6, (double)fabsl(g6 / ex6 - 1));
        }
    }
    printf("exp2neg max rel err %.3g (2^%.1f); after 6 squarings %.3g\n", worst, log2(worst), worst6);
    // 2. d2 path
    std::uniform_real_distribution<double> U(0, 1);
    double worstd = 0, worsty = 0;
    for (int it = 0; it < 1000000; it++) {
        double scale = exp2(floor(U(rng) * 40 - 20));
        double na = scale * (0.1 + U(rng)) * 100, nb = scale * (0.1 + U(rng)) * 100;
        double cosv = 2 * U(rng) - 1; if (it % 5 == 0) cosv = 1 - 1e-9 * U(rng);
        long double ab = (long double)sqrtl((long double)na * nb) * cosv;
        if (it % 5 == 0) { nb = na * (1 + 1e-7 * U(rng)); ab = sqrtl((long double)na * nb) * (1 - 1e-10L * U(rng)); }
        // exponents: 2^ea in [4 max|a_k|, ...]; take max|a_k| between |a|/sqrt(2000) and |a|
        int ea, eb; frexp(sqrt(na) * (0.03 + 0.97 * U(rng)), &ea); ea += 1 + (rng() & 1);
        frexp(sqrt(nb) * (0.03 + 0.97 * U(rng)), &eb); eb += 1 + (rng() & 1);
        __int128 V = (__int128)roundl(ab * exp2l(64 - ea - eb));
        long long lo = (long long)(V & 0xffffffffll), hi = (long long)((V - lo) >> 32);
        // lo may also be large/negative in the kernel; emulate: move some of hi into lo
        long long mv = (long long)(rng() % 2000000) - 1000000; hi -= mv; lo += mv << 32;
        int en; frexp(na + nb, &en);  // na + nb < 2^en
        const int EN = en;
        long long NA = llrint(ldexp(na, UNIT - EN)), NB = llrint(ldexp(nb, UNIT - EN));
        int s = EN + 5 - ea - eb; if (s < 0) { printf("s<0!\n"); return 1; } if (s > 127) s = 127;
        uint64_t D = dfix(hi, lo, s, NA + NB);
        long double dex = (long double)na + nb - 2 * ab; if (dex < 0) dex = 0;
        long double dg = (long double)D * exp2l(EN - UNIT);
        worstd = std::max(worstd, (double)(fabsl(dg - dex) / ((long double)na + nb)));
        // sigma: y = d2 c
        double sigma = sqrt(na + nb) * (0.02 + U(rng)); double isig = -0.5 / (sigma * sigma);
        uint64_t mant; int fe; aqml::i8e::sigma_const(isig, mant, fe);
        uint64_t Z = zfix(D, mant, 14 - EN - fe);
        long double yex = dex * (long double)(-isig) * 1.4426950408889634073599246810019L;
        if (Z != ~0ull && yex < 1090) worsty = std::max(worsty, (double)fabsl((long double)Z / (long double)(1ull << FRAC) - yex));
        if (Z == ~0ull && yex < 1090) { printf("saturated wrongly y=%Lg\n", yex); return 1; }
    }
    printf("d2 fixed point: max |err| / (na+nb) = %.3g (2^%.1f); max abs err of y = %.3g\n", worstd, log2(worstd), worsty);
    // 3. sums + conversion
    double worsts = 0;
    for (int it = 0; it < 200000; it++) {
        int n = 1 + rng() % 16; int kbase = rng() % 900;
        F2 ms[16]; int ks[16]; long double ex = 0; int kref = KBIG;
        for (int j = 0; j < n; j++) {
            uint64_t Z = ((uint64_t)kbase << FRAC) + rng() % ((uint64_t)30 << FRAC);
            exp2neg(Z, tab, ms[j], ks[j]);
            ex += ((long double)ms[j].h + ms[j].l) * exp2l(-(long double)ks[j]);
            kref = std::min(kref, ks[j]);
        }
        long long acc = 0;
        for (int j = 0; j < n; j++) acc += term_fix(ms[j], ks[j], kref);
        double got = fix_to_double((uint64_t)acc, -SUMBITS - kref);
        worsts = std::max(worsts, (double)fabsl(got / ex - 1));
    }
    printf("sum + conversion max rel err %.3g\n", worsts);
    printf("fix_to_double(1,0)=%g (3,-1)=%g ((1<<53)+1, 0)=%.17g tiny=%g\n", fix_to_double(1, 0), fix_to_double(3, -1), fix_to_double((1ull << 53) + 1, 0), fix_to_double(5, -1080));
    return 0;
}
