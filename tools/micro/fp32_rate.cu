// Microbenchmark (diagnostic, not product): issue rate of 3-register FFMA vs packed FFMA2 (fma.rn.f32x2) vs a
// FMUL/FFMA/FADD mix on sm_100a, as warp-instructions per clock per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_rate fp32_rate.cu ; run: ./fp32_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float fmul1(float a, float b) { float d; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fadd1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

template <int MODE, int CH>
__global__ void k(float *out, float x, float y, int iters, long long *cyc) {
    const long long t0 = clock64();
    float a[CH];
    unsigned long long p[CH];
#pragma unroll
    for (int j = 0; j < CH; j++) { a[j] = threadIdx.x * 0.001f + j; p[j] = ((unsigned long long)__float_as_uint(a[j]) << 32) | __float_as_uint(a[j] + 1.0f); }
    const unsigned long long xx = ((unsigned long long)__float_as_uint(x) << 32) | __float_as_uint(x);
    const unsigned long long yy = ((unsigned long long)__float_as_uint(y) << 32) | __float_as_uint(y);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int j = 0; j < CH; j++) {
                if (MODE == 0) a[j] = ffma1(a[j], x, y);
                if (MODE == 1) p[j] = ffma2(p[j], xx, yy);
                if (MODE == 2) a[j] = (u & 1) ? fmul1(a[j], x) : fadd1(a[j], y);
                if (MODE == 3) a[j] = ffma1(a[j], a[(j + 1) % CH], a[(j + 2) % CH]);   // three distinct register operands
            }
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < CH; j++) s += a[j] + __uint_as_float((unsigned)p[j]) + __uint_as_float((unsigned)(p[j] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = clock64() - t0;
}

template <int MODE, int CH>
void run(const char *name, int warps_per_sm) {
    int dev = 0; cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    const int sms = pr.multiProcessorCount;
    const int block = 128, grid = sms * warps_per_sm / 4, iters = 2048;
    float *out; cudaMalloc(&out, sizeof(float) * grid * block);
    long long *cyc; cudaMalloc(&cyc, 8);
    k<MODE, CH><<<grid, block>>>(out, 1.0001f, 0.5f, 16, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE, CH><<<grid, block>>>(out, 1.0001f, 0.5f, iters, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double winstr = (double)grid * (block / 32) * iters * 8 * CH;
    const double clk = ms * 1e-3 * khz * 1e3;   // at the maximum SM clock (bench.py samples 1965 MHz under load on this pool)
    printf("%-28s warps/SM %2d chains %d: %.3f warp-instr/clk/SM sub-partition (at %d MHz), %.1f us\n", name, warps_per_sm, CH,
           winstr / clk / (sms * 4), khz / 1000, ms * 1e3);
    cudaFree(cyc);
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0, 8>("FFMA r,r,r (2 shared srcs)", w);
        run<3, 8>("FFMA 3 distinct regs", w);
        run<1, 8>("FFMA2 (f32x2)", w);
        run<2, 8>("FMUL/FADD alternating", w);
    }
    run<0, 2>("FFMA 2 chains", 16);
    run<1, 2>("FFMA2 2 chains", 16);
    run<0, 1>("FFMA 1 chain", 16);
    run<1, 1>("FFMA2 1 chain", 16);
    return 0;
}
