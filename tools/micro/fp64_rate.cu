// Microbenchmark (diagnostic, not product): issue rate and dependent latency of DFMA and of the float <-> double conversions
// (F2F) on sm_100a, as warp-instructions per clock per SM sub-partition -- what the double-precision attitude path of
// integrate_thread (tvc_device.cuh) pays.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double dfma1(double a, double b, double c) { double d; asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ double f2d(float a) { double d; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(a)); return d; }
__device__ __forceinline__ float d2f(double a) { float d; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(d) : "d"(a)); return d; }

// MODE 0: DFMA chains; 1: FFMA chains (reference); 2: float -> double -> float round trips (two F2F per count)
template <int MODE, int CH>
__global__ void k(float *out, float x, float y, int iters) {
    double a[CH]; float f[CH];
#pragma unroll
    for (int j = 0; j < CH; j++) { a[j] = threadIdx.x * 0.001 + j; f[j] = threadIdx.x * 0.001f + j; }
    const double xd = x, yd = y;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int j = 0; j < CH; j++) {
                if (MODE == 0) a[j] = dfma1(a[j], xd, yd);
                if (MODE == 1) f[j] = ffma1(f[j], x, y);
                if (MODE == 2) f[j] = d2f(f2d(f[j]));
            }
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < CH; j++) s += (float)a[j] + f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int CH>
void run(const char *name, int warps_per_sm) {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount, block = 128, grid = sms * warps_per_sm / 4, iters = 1024;
    float *out; cudaMalloc(&out, sizeof(float) * grid * block);
    k<MODE, CH><<<grid, block>>>(out, 1.0001f, 0.5f, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE, CH><<<grid, block>>>(out, 1.0001f, 0.5f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double winstr = (double)grid * (block / 32) * iters * 8 * CH * (MODE == 2 ? 2 : 1);
    const double clk = ms * 1e-3 * khz * 1e3;
    const double per_sp = winstr / clk / (sms * 4);
    printf("%-30s warps/SM %2d chains %d: %.4f warp-instr/clk/sub-partition = %.1f lanes/clk/SM; %.1f clk per dependent instr per warp\n",
           name, warps_per_sm, CH, per_sp, per_sp * 128, (double)(warps_per_sm / 4) / per_sp / 1.0 / CH * 1.0);
    cudaFree(out);
}

int main() {
    for (int w : {4, 16, 32}) {
        run<0, 4>("DFMA", w);
        run<1, 4>("FFMA", w);
        run<2, 4>("F2F.F64.F32 + F2F.F32.F64", w);
    }
    run<0, 1>("DFMA 1 chain", 4);
    run<1, 1>("FFMA 1 chain", 4);
    run<2, 1>("F2F round trip 1 chain", 4);
    return 0;
}
