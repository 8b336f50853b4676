// Issue-rate microbenchmark: FFMA (3-register form) vs FFMA2 (fma.rn.f32x2) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
// Prints warp-instructions per cycle per SM sub-partition for 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm volatile("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
    float r;
    asm volatile("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

template <int MODE> __global__ void k(float* out, long long* cyc, int iters, float seed) {
    float x[8];
    u64 y[8];
    float m = seed, n = seed * 0.5f;
    u64 m2 = ((u64)__float_as_uint(m) << 32) | __float_as_uint(n);
    for (int i = 0; i < 8; ++i) { x[i] = seed + i; y[i] = ((u64)__float_as_uint(seed + i) << 32) | __float_as_uint(seed - i); }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = ffma1(x[i], m, n);
                if (MODE == 1) y[i] = ffma2(y[i], m2, m2);
                if (MODE == 2) { // mixed: 1 FFMA2 + 1 FFMA (is the scalar pipe free while FFMA2 runs?)
                    if (i & 1) x[i] = ffma1(x[i], m, n);
                    else y[i] = ffma2(y[i], m2, m2);
                }
                if (MODE == 3) { // dependent chain latency: one chain only
                    if (i == 0) x[0] = ffma1(x[0], m, n);
                }
                if (MODE == 4) {
                    if (i == 0) y[0] = ffma2(y[0], m2, m2);
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float((unsigned)y[i]) + __uint_as_float((unsigned)(y[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int threads, int inst_per_iter) {
    float* out;
    long long* cyc;
    int blocks = 148, iters = 4096;
    cudaMalloc(&out, blocks * 1024 * 4);
    cudaMalloc(&cyc, blocks * 8);
    k<MODE><<<blocks, threads>>>(out, cyc, 16, 1.0f);
    k<MODE><<<blocks, threads>>>(out, cyc, iters, 1.0f);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < blocks; ++i) c += h[i];
    c /= blocks;
    double warps_per_smsp = threads / 32.0 / 4.0;
    double inst = (double)iters * inst_per_iter * warps_per_smsp;
    printf("%-28s %4d thr/SM (%.0f warps/SMSP): %8.0f cycles, %.3f warp-inst/cycle/SMSP, %.2f cycles/iter/warp\n", name, threads,
           warps_per_smsp, c, inst / c, c / iters);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int threads : {128, 256, 512, 1024}) {
        run<0>("FFMA  x32 independent", threads, 32);
        run<1>("FFMA2 x32 independent", threads, 32);
        run<2>("FFMA2+FFMA x16+16", threads, 32);
    }
    run<3>("FFMA  dependent chain x4", 128, 4);
    run<4>("FFMA2 dependent chain x4", 128, 4);
    return 0;
}
