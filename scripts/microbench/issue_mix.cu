// Microbenchmark: can non-FP64 instructions issue in the shadow of DFMA?
// Each thread runs K independent DFMA chains plus M independent integer (LOP3/IADD) or FP32 chains per iteration.
#include <cstdio>
#include <cuda_runtime.h>

template <int NF64, int NINT, int NF32>
__global__ void __launch_bounds__(128) mix(double* out, int iters, double a, double b, int ia, float fa, int flag) {
    double x[8]; int y[16]; float z[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) { y[i] = threadIdx.x * 7 + i; z[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NF64; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
            for (int i = 0; i < NINT; ++i) y[i] = (y[i] ^ ia) + (y[i] >> 3);
#pragma unroll
            for (int i = 0; i < NF32; ++i) z[i] = fmaf(z[i], fa, 1.0f);
        }
    }
    double s = 0; 
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    int t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { t += y[i]; s += z[i]; }
    if (flag) out[threadIdx.x] = s + (double)t;
}

template <int NF64, int NINT, int NF32>
void run(const char* name, int warps_per_smsp) {
    double* d; cudaMalloc(&d, 8 * 128);
    int dev_sms = 148;
    int blocks = dev_sms * warps_per_smsp;   // 128 threads = 4 warps = 1 per SMSP per block
    int iters = 5000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<NF64, NINT, NF32><<<blocks, 128>>>(d, 100, 1.0000001, 1e-9, 0x5bd1e995, 1.000001f, 0);
    cudaEventRecord(e0);
    mix<NF64, NINT, NF32><<<blocks, 128>>>(d, iters, 1.0000001, 1e-9, 0x5bd1e995, 1.000001f, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * 1.95e9;
    double per_iter = cyc / (iters * 4.0) / warps_per_smsp;   // cycles per (warp, inner r-iteration)
    printf("%-28s warps/SMSP=%d  ms=%.3f  cycles per warp-iter=%.2f  (F64=%d INT=%d(x3 instr) F32=%d)\n", name,
           warps_per_smsp, ms, per_iter, NF64, NINT, NF32);
    cudaFree(d);
}

int main() {
    for (int w : {1, 2, 4, 8}) {
        run<8, 0, 0>("8 DFMA", w);
        run<8, 4, 0>("8 DFMA + 4 int-chains", w);
        run<8, 8, 0>("8 DFMA + 8 int-chains", w);
        run<8, 0, 8>("8 DFMA + 8 FFMA", w);
        run<8, 0, 16>("8 DFMA + 16 FFMA", w);
        run<0, 8, 0>("8 int-chains", w);
        run<0, 0, 16>("16 FFMA", w);
        run<4, 0, 0>("4 DFMA", w);
        run<2, 0, 0>("2 DFMA", w);
        run<1, 0, 0>("1 DFMA", w);
    }
    return 0;
}
