// Microbenchmark: sustained ex2.approx.ftz.f32 (MUFU.EX2) rate on this GPU under the conditions of the
// disparity-head kernels (few resident warps, FP32 work interleaved).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/mufu_bench.cu -o /tmp/mufu_bench
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: ex2 only.  MODE 1: per ex2: 1 FFMA before (z), FADD + FFMA after (den, num) -- the head's mix.
// MODE 2: like 1 plus 3 extra FFMA per ex2 (the blend/bookkeeping share).
template <int ILP, int MODE>
__global__ void k(float* out, int iters, float seed) {
    float a[ILP], dg[ILP], ng[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = seed + 0.001f * (threadIdx.x + i); dg[i] = 0.f; ng[i] = 0.f; }
    float l = 0.333f + seed * 1e-6f, dlt = 0.01f, kf = 1.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            float z = MODE == 0 ? a[i] : __fmaf_rn(l, dlt, a[i]);
            float e;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z));
            if (MODE == 0) a[i] = e - 1.0f;
            if (MODE >= 1) { dg[i] += e; ng[i] = __fmaf_rn(e, kf, ng[i]); }
            if (MODE == 2) { a[i] = __fmaf_rn(a[i], 0.9999f, 1e-5f); a[i] = __fmaf_rn(a[i], 0.9999f, 1e-5f); a[i] = __fmaf_rn(a[i], 0.9999f, 1e-5f); }
        }
        kf += 1.f; dlt = -dlt;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + dg[i] + ng[i];
    if (s == 123.456f) out[0] = s;
}

template <int ILP, int MODE>
void run(const char* name, int blocks_per_sm, int threads) {
    float* d; cudaMalloc(&d, 4);
    const int iters = 2048, blocks = 148 * blocks_per_sm;
    k<ILP, MODE><<<blocks, threads>>>(d, iters, -0.5f);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<ILP, MODE><<<blocks, threads>>>(d, iters, -0.5f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)blocks * threads * iters * ILP;
    printf("%-44s warps/SM=%2d: %.3f ms, %.2f ex2/clk/SM @1.965GHz\n", name, blocks_per_sm * threads / 32, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
}

int main() {
    run<8, 0>("ex2 only ILP8", 8, 256);
    run<8, 0>("ex2 only ILP8", 4, 128);
    run<27, 0>("ex2 only ILP27", 4, 128);
    run<27, 1>("ffma+ex2+fadd+ffma ILP27", 4, 128);
    run<27, 1>("ffma+ex2+fadd+ffma ILP27", 8, 128);
    run<27, 1>("ffma+ex2+fadd+ffma ILP27", 2, 128);
    run<27, 2>("ffma+ex2+fadd+ffma+3ffma ILP27", 4, 128);
    run<27, 2>("ffma+ex2+fadd+ffma+3ffma ILP27", 8, 128);
    run<9, 1>("ffma+ex2+fadd+ffma ILP9", 4, 128);
    run<9, 1>("ffma+ex2+fadd+ffma ILP9", 8, 128);
    run<9, 2>("ffma+ex2+fadd+ffma+3ffma ILP9", 8, 128);
    run<9, 2>("ffma+ex2+fadd+ffma+3ffma ILP9", 16, 128);
    return 0;
}
