/* Plain-C client of the C ABI (no Python, no torch): what a non-Python binder (f2py replacement, cgo, JNI ...)
 * would do.  Links libaqml_b200.so, calls the f2py-equivalent entry points with host buffers and checks them
 * against direct loops.  Built and run by tests/test_gpu_c_abi.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "../include/aqml_b200.h"

static double frand(unsigned *s) { *s = *s * 1664525u + 1013904223u; return (double)(*s >> 8) / 16777216.0; }

int main(void) {
    const int na = 37, nb = 21, D = 301;
    unsigned seed = 7;
    double *a = malloc(sizeof(double) * na * D), *b = malloc(sizeof(double) * nb * D);
    double *k = malloc(sizeof(double) * na * nb), *d = malloc(sizeof(double) * na * nb);
    for (int i = 0; i < na * D; i++) a[i] = frand(&seed);
    for (int i = 0; i < nb * D; i++) b[i] = frand(&seed);
    const double sigma = 9.0;
    if (aqml_device_count() < 1) { fprintf(stderr, "no device\n"); return 2; }
    if (aqml_fgaussian_kernel(a, na, b, nb, D, k, sigma) != AQML_OK) { fprintf(stderr, "%s\n", aqml_last_error()); return 3; }
    if (aqml_fl2_distance(a, b, na, nb, D, d) != AQML_OK) { fprintf(stderr, "%s\n", aqml_last_error()); return 3; }
    double worst = 0.0;
    for (int i = 0; i < nb; i++)
        for (int j = 0; j < na; j++) {
            double s = 0.0;
            for (int f = 0; f < D; f++) { double t = a[j * D + f] - b[i * D + f]; s += t * t; }
            double ref = exp(-0.5 * s / (sigma * sigma));
            double e1 = fabs(k[j + na * i] - ref) / ref, e2 = fabs(d[j + na * i] - sqrt(s)) / sqrt(s);
            if (e1 > worst) worst = e1;
            if (e2 > worst) worst = e2;
        }
    /* local kernel: 3 + 2 molecules, two elements */
    const int nas1[3] = {2, 3, 1}, nas2[2] = {2, 2};
    const int zs1[6] = {1, 6, 1, 1, 6, 6}, zs2[4] = {6, 1, 1, 6};
    double sig[2 * 6];
    for (int s = 0; s < 2; s++) for (int z = 0; z < 6; z++) sig[s * 6 + z] = (s + 1) * (6.0 + z);
    double K[2 * 3 * 2];
    if (aqml_calc_lk_gaussian(a, b, zs1, zs2, nas1, nas2, 6, 0, sig, 3, 2, 2, D, K) != AQML_OK) {
        fprintf(stderr, "%s\n", aqml_last_error()); return 3;
    }
    int off1[4] = {0, 2, 5, 6}, off2[3] = {0, 2, 4};
    for (int s = 0; s < 2; s++)
        for (int m1 = 0; m1 < 3; m1++)
            for (int m2 = 0; m2 < 2; m2++) {
                double ref = 0.0;
                for (int i = off1[m1]; i < off1[m1 + 1]; i++)
                    for (int j = off2[m2]; j < off2[m2 + 1]; j++) {
                        if (zs1[i] != zs2[j]) continue;
                        double q = 0.0;
                        for (int f = 0; f < D; f++) { double t = a[i * D + f] - b[j * D + f]; q += t * t; }
                        double sg = sig[s * 6 + zs1[i] - 1];
                        ref += exp(-0.5 * q / (sg * sg));
                    }
                double got = K[(s * 3 + m1) * 2 + m2];
                double e = ref > 0 ? fabs(got - ref) / ref : fabs(got);
                if (e > worst) worst = e;
            }
    /* error path: bad element number is reported, not fatal */
    int badz[6] = {1, 6, 1, 1, 6, 99};
    int rc = aqml_calc_lk_gaussian(a, b, badz, zs2, nas1, nas2, 6, 0, sig, 3, 2, 2, D, K);
    printf("c_abi_client: worst relative error %.3e, bad-argument rc=%d (%s)\n", worst, rc, aqml_last_error());
    return (worst < 1e-10 && rc == AQML_EINVAL) ? 0 : 1;
}
