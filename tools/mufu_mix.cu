// Microbenchmark: cost of MUFU.EX2 when mixed with N other instructions of a given kind per ex2.
// kind 0: FFMA (independent per-lane chain)  1: FADD  2: FFMA2 (packed; N counts instructions)  3: IADD (ALU pipe)
// Reports clk per ex2 per SMSP (8.0 = SFU floor).
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda.h>

template <int ILP, int N, int KIND>
__global__ void k(float* out, int iters, float seed) {
    float a[ILP];
    float2 p[ILP];
    int q[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = seed + 0.001f * (threadIdx.x + i); p[i] = make_float2(seed, seed * 0.5f + i); q[i] = i + threadIdx.x; }
    const float2 c1 = make_float2(0.9999f, 0.9998f), c2 = make_float2(1e-5f, 2e-5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            a[i] = a[i] - 1.0f;  // keep the chain bounded (1 FADD always present)
#pragma unroll
            for (int n = 0; n < N; ++n) {
                if (KIND == 0) p[i].x = __fmaf_rn(p[i].x, 0.9999f, 1e-5f);
                if (KIND == 1) p[i].x = p[i].x + 1e-5f;
                if (KIND == 2) p[i] = __ffma2_rn(p[i], c1, c2);
                if (KIND == 3) q[i] = q[i] * 3 + it;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + p[i].x + p[i].y + q[i];
    if (s == 123.456f) out[0] = s;
}

template <int ILP, int N, int KIND>
void run(int blocks_per_sm, int threads) {
    float* d; cudaMalloc(&d, 4);
    const int iters = 2048, blocks = 148 * blocks_per_sm;
    k<ILP, N, KIND><<<blocks, threads>>>(d, iters, -0.5f);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<ILP, N, KIND><<<blocks, threads>>>(d, iters, -0.5f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double warp_ex2_per_smsp = (double)blocks_per_sm * threads / 32 / 4 * iters * ILP;
    const char* names[] = {"FFMA", "FADD", "FFMA2", "IMAD"};
    printf("ILP%-2d warps/SM=%2d  ex2 + 1 FADD + %d %-5s : %.2f clk/ex2/SMSP (@1.965GHz)\n", ILP, blocks_per_sm * threads / 32, N, names[KIND],
           ms * 1e-3 * 1.965e9 / warp_ex2_per_smsp);
    cudaFree(d);
}

int main() {
    run<9, 0, 0>(4, 128);
    run<9, 1, 0>(4, 128); run<9, 2, 0>(4, 128); run<9, 3, 0>(4, 128); run<9, 4, 0>(4, 128); run<9, 6, 0>(4, 128); run<9, 8, 0>(4, 128);
    run<9, 3, 0>(8, 128); run<9, 6, 0>(8, 128); run<9, 6, 0>(16, 128);
    run<9, 3, 1>(4, 128); run<9, 6, 1>(4, 128);
    run<9, 1, 2>(4, 128); run<9, 2, 2>(4, 128); run<9, 3, 2>(4, 128); run<9, 4, 2>(4, 128);
    run<9, 2, 2>(8, 128); run<9, 3, 2>(8, 128);
    run<9, 3, 3>(4, 128); run<9, 6, 3>(4, 128);
    run<27, 3, 0>(4, 128); run<27, 6, 0>(4, 128); run<27, 3, 2>(4, 128);
    return 0;
}
