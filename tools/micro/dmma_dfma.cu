// Micro-benchmark: do DMMA.8x8x4 (tensor pipe) and DFMA (FP64 CUDA cores) overlap on sm_100a?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_dfma dmma_dfma.cu ; run: ./dmma_dfma
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode bit0: warps with (warp&1)==0 run DMMA ; bit1: warps with (warp&1)==1 run DFMA.  mode 1: all-even DMMA only (odd idle) etc.
__global__ void k(double *out, int iters, int mode, int split) {
    int warp = threadIdx.x >> 5;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[16];
    for (int i = 0; i < 16; i++) c[i] = i;
    bool is_mma = (split == 0) ? (mode == 1) : ((warp & 1) == 0);
    bool active = (split == 0) ? true : (is_mma ? (mode & 1) : (mode & 2));
    if (active) {
        if (is_mma) {
            for (int it = 0; it < iters; it++) {
#pragma unroll
                for (int i = 0; i < 16; i += 2) dmma(c[i], c[i + 1], a, b);
            }
        } else {
            for (int it = 0; it < iters; it++) {
#pragma unroll
                for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double *out;
    cudaMalloc(&out, 148 * 8 * 256 * 8);
    int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct { const char *name; int mode, split; } cfg[] = {
        {"all 8 warps DMMA", 1, 0}, {"all 8 warps DFMA", 2, 0},
        {"4 warps DMMA, 4 idle", 1, 1}, {"4 idle, 4 warps DFMA", 2, 1}, {"4 warps DMMA + 4 warps DFMA", 3, 1}};
    for (auto &c : cfg) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            k<<<148 * 2, 256>>>(out, iters, c.mode, c.split);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        // flops: DMMA = 512 flop per warp-instr; DFMA = 64 flop per warp-instr
        double warps = 148.0 * 2 * 8;
        double mma_w = c.split ? ((c.mode & 1) ? warps / 2 : 0) : (c.mode == 1 ? warps : 0);
        double fma_w = c.split ? ((c.mode & 2) ? warps / 2 : 0) : (c.mode == 2 ? warps : 0);
        double tf_mma = mma_w * iters * 8.0 * 512 / (ms * 1e-3) / 1e12, tf_fma = fma_w * iters * 16.0 * 64 / (ms * 1e-3) / 1e12;
        printf("%-32s %8.3f ms   DMMA %6.2f TF  DFMA %6.2f TF  total %6.2f TF\n", c.name, ms, tf_mma, tf_fma, tf_mma + tf_fma);
    }
    return 0;
}
