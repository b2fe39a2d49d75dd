// Microbenchmark: issue cost of packed fp32x2 ops (FFMA2/FMUL2/FADD2) vs scalar on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int MODE> __global__ void k(float* out, int iters, float s) {
    float a[8]; u64 p[4];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int i = 0; i < 4; ++i) p[i] = ((u64)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
    u64 s2 = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {            // 8 scalar FFMA
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = ffma1(a[i], s, s);
        } else if (MODE == 1) {     // 4 packed FFMA2 (same flops)
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = ffma2(p[i], s2, s2);
        } else if (MODE == 2) {     // 8 scalar FFMA + 8 ALU (select-ish) ops: issue-bound mix
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = ffma1(a[i], s, s); }
        }
    }
    for (int i = 0; i < 8; ++i) acc += a[i];
    for (int i = 0; i < 4; ++i) acc += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> __global__ void kmix(float* out, int iters, float s) {
    float a[8]; u64 p[4]; float m[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; m[i] = i; }
    for (int i = 0; i < 4; ++i) p[i] = ((u64)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
    u64 s2 = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(s));   // 8 ALU-pipe ops
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = ffma1(a[i], s, s);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = ffma2(p[i], s2, s2);
        }
    }
    float acc = 0.f;
    for (int i = 0; i < 8; ++i) acc += a[i] + m[i];
    for (int i = 0; i < 4; ++i) acc += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int iters = 20000, grid = 148 * 8, block = 256;
    float t0 = timeit([&] { k<0><<<grid, block>>>(out, iters, 1.0001f); });
    float t1 = timeit([&] { k<1><<<grid, block>>>(out, iters, 1.0001f); });
    float t2 = timeit([&] { kmix<0><<<grid, block>>>(out, iters, 1.0001f); });
    float t3 = timeit([&] { kmix<1><<<grid, block>>>(out, iters, 1.0001f); });
    double flops = 2.0 * 8 * iters * (double)grid * block;
    printf("8xFFMA   : %.3f ms  %.1f TFLOP/s\n4xFFMA2  : %.3f ms  %.1f TFLOP/s\n", t0, flops / t0 / 1e9, t1, flops / t1 / 1e9);
    printf("8xFMNMX+8xFFMA : %.3f ms\n8xFMNMX+4xFFMA2: %.3f ms\n", t2, t3);
    printf("cuda err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
