// Microbenchmark: DFMA throughput as a function of operand sources (registers vs uniform/constant).
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: x = fma(x, a, b)   a,b kernel params (uniform)
// MODE 1: x = fma(x, y, b)   y per-thread register, b uniform
// MODE 2: x = fma(x, y, z)   y,z per-thread registers (distinct per chain)
// MODE 3: x = fma(y, z, x)   accumulate form, y,z registers shared across chains
// MODE 4: DMUL x = x*y ; MODE 5: DADD x = x + y
template <int MODE, int NCH>
__global__ void __launch_bounds__(128) k(double* out, int iters, double a, double b, int flag) {
    double x[NCH], y[NCH], z[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { x[i] = threadIdx.x + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); z[i] = 1e-9 * (threadIdx.x + 2 * i); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                if (MODE == 0) x[i] = fma(x[i], a, b);
                if (MODE == 1) x[i] = fma(x[i], y[i], b);
                if (MODE == 2) x[i] = fma(x[i], y[i], z[i]);
                if (MODE == 3) x[i] = fma(y[0], z[i], x[i]);
                if (MODE == 4) x[i] = x[i] * y[i];
                if (MODE == 5) x[i] = x[i] + y[i];
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += x[i] + y[i] + z[i];
    if (flag) out[threadIdx.x] = s;
}

template <int MODE, int NCH>
void run(const char* name, int wps) {
    double* d; cudaMalloc(&d, 8 * 128);
    int blocks = 148 * wps, iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, NCH><<<blocks, 128>>>(d, 10, 1.0000001, 1e-9, 0);
    cudaEventRecord(e0);
    k<MODE, NCH><<<blocks, 128>>>(d, iters, 1.0000001, 1e-9, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * 1.95e9;
    printf("%-34s chains=%d warps/SMSP=%d  cycles per FP64 warp-instr = %.2f\n", name, NCH, wps,
           cyc / ((double)iters * 8 * NCH * wps));
    cudaFree(d);
}

int main() {
    for (int w : {1, 4, 8}) {
        run<0, 8>("fma(x, U, U)", w);
        run<1, 8>("fma(x, R, U)", w);
        run<2, 8>("fma(x, R, R) distinct", w);
        run<3, 8>("fma(R0, R, x) accumulate", w);
        run<4, 8>("mul(x, R)", w);
        run<5, 8>("add(x, R)", w);
        run<2, 3>("fma(x, R, R) 3 chains", w);
        run<2, 1>("fma(x, R, R) 1 chain (latency)", w);
    }
    return 0;
}
