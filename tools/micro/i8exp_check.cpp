// Host check of aqml_b200/csrc/i8exp.cuh (the integer exp / sums of the int8 engine's epilogue) against extended precision.
//   g++ -O2 -o /tmp/i8exp_check tools/micro/i8exp_check.cpp && /tmp/i8exp_check
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
#include "../../aqml_b200/csrc/i8exp.cuh"
using namespace aqml::i8x;

int main() {
    std::vector<u64> tab(TAB_N);
    for (int j = 0; j < TAB_N; j++) {
        long double v = exp2l(-(long double)j / TAB_N) * 9007199254740992.0L;  // 2^53
        tab[j] = j ? (u64)llroundl(v) : (1ull << 53) - 1;   // 2^53 would wrap the 32-bit top word in exp2neg
    }
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    double worst = 0, worst_chain = 0;
    for (int it = 0; it < 4000000; it++) {
        double T;
        const int mode = it % 4;
        if (mode == 0) T = U(rng) * 1100.0;
        else if (mode == 1) T = std::exp(-40.0 * U(rng));
        else if (mode == 2) T = U(rng) * 40.0;
        else T = std::ldexp(1.0 + U(rng), (int)(U(rng) * 24) - 12);
        u64 b;
        memcpy(&b, &T, 8);
        u64 F = fix_from_bits((u32)(b >> 32), (u32)b);
        for (int k = 0; k <= 8; k += 2) {
            u64 Fk = fix_shl(F, k);
            u64 M;
            int n;
            exp2neg(Fk, tab.data(), M, n);
            u64 tb = term_bits(M, n);
            double got;
            memcpy(&got, &tb, 8);
            long double Tk = (long double)T * (1 << k);
            long double ref = exp2l(-Tk);
            if (Tk >= 1021.9L) {
                if (Tk > 1023.0L && got != 0.0) { printf("expected 0 at T=%Lg got %g\n", Tk, got); return 1; }
                continue;
            }
            double err = (double)fabsl(((long double)got - ref) / ref);
            if (k == 0) worst = std::max(worst, err); else worst_chain = std::max(worst_chain, err);
        }
    }
    printf("exp2neg: worst relative error %.3e (base), %.3e (shifted by 2..8 bits)\n", worst, worst_chain);
    // specials
    {
        double inf = INFINITY, big = 5000.0, z = 0.0, nanv = NAN;
        for (double T : {inf, big, z, nanv}) {
            u64 b; memcpy(&b, &T, 8);
            b &= 0x7fffffffffffffffull;
            u64 F = fix_from_bits((u32)(b >> 32), (u32)b), M; int n;
            exp2neg(F, tab.data(), M, n);
            u64 tb = term_bits(M, n); double got; memcpy(&got, &tb, 8);
            printf("T = %g -> %g\n", T, got);
        }
    }
    // sums: terms over a wide range, random order
    double worst_sum = 0;
    for (int it = 0; it < 200000; it++) {
        const int cnt = 1 + (int)(U(rng) * 16);
        const double base = U(rng) * 900.0, spread = std::exp(U(rng) * 5.0) - 1.0;
        u64 acc = 0; int nacc = 4095;
        long double ref = 0;
        for (int i = 0; i < cnt; i++) {
            double T = base + U(rng) * spread;
            u64 b; memcpy(&b, &T, 8);
            u64 M; int n;
            exp2neg(fix_from_bits((u32)(b >> 32), (u32)b), tab.data(), M, n);
            add_term(acc, nacc, M, n);
            ref += exp2l(-(long double)T);
        }
        if (acc == 0 || nacc >= N_ZERO) continue;
        long double got = ldexpl((long double)acc, -(53 + nacc));
        worst_sum = std::max(worst_sum, (double)fabsl((got - ref) / ref));
    }
    printf("aligned integer sums of <= 16 terms: worst relative error %.3e\n", worst_sum);
    return (worst < 2e-12 && worst_chain < 2e-12 && worst_sum < 2e-12) ? 0 : 1;
}
