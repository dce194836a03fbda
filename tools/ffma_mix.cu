// Does sm_100a run scalar FP32 (fmalite) concurrently with packed FFMA2 (fmaheavy)?
// Each thread owns NP packed and NS scalar independent FMA chains; prints FP32 lane-ops per clk per
// SM sub-partition (peak with scalar only = 32).   nvcc -O3 -arch=sm_100a tools/ffma_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NP, int NS, int NM>
__global__ void __launch_bounds__(128) k(float* out, const float* in, int iters) {
    float2 p[NP > 0 ? NP : 1];
    float s[NS > 0 ? NS : 1];
    float m[NM > 0 ? NM : 1];
    const float2 a2 = make_float2(in[0], in[1]), b2 = make_float2(in[2], in[3]);
    const float a = in[4], b = in[5];
#pragma unroll
    for (int i = 0; i < NP; ++i) p[i] = make_float2(in[6 + i] + threadIdx.x, in[7 + i]);
#pragma unroll
    for (int i = 0; i < NS; ++i) s[i] = in[8 + i] + threadIdx.x;
#pragma unroll
    for (int i = 0; i < NM; ++i) m[i] = in[9 + i] * 1e-3f;
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NP; ++i) p[i] = __ffma2_rn(p[i], a2, b2);
#pragma unroll
        for (int i = 0; i < NS; ++i) s[i] = __fmaf_rn(s[i], a, b);
#pragma unroll
        for (int i = 0; i < NM; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) r += p[i].x + p[i].y;
#pragma unroll
    for (int i = 0; i < NS; ++i) r += s[i];
#pragma unroll
    for (int i = 0; i < NM; ++i) r += m[i];
    if (r == 123.456f) out[0] = r;
}

template <int NP, int NS, int NM>
void run(const float* din, float* dout, int bps, double ghz) {
    const int iters = 4096, blocks = 148 * bps, threads = 128;
    k<NP, NS, NM><<<blocks, threads>>>(dout, din, iters);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<NP, NS, NM><<<blocks, threads>>>(dout, din, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double warps_per_smsp = (double)bps * threads / 32 / 4;
    const double clk = ms * 1e-3 * ghz * 1e9;
    const double laneops = warps_per_smsp * iters * (NP * 64.0 + NS * 32.0);
    const double instr = warps_per_smsp * iters * (NP + NS + NM);
    printf("NP=%2d NS=%2d NM=%d warps/SM=%2d: %.3f ms  %.1f fp32 lane-ops/clk/SMSP  %.2f instr/clk/SMSP  (%.2f clk per iteration per warp-slot)\n",
           NP, NS, NM, bps * 4, ms, laneops / clk, instr / clk, clk / (warps_per_smsp * iters));
}

int main() {
    float h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.5f + 0.001f * i;
    float *din, *dout; cudaMalloc(&din, sizeof(h)); cudaMalloc(&dout, 4);
    cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
    int dev = 0; cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal\n", pr.name, pr.multiProcessorCount, ghz);
    for (int bps : {4, 8}) {
        run<0, 16, 0>(din, dout, bps, ghz);
        run<8, 0, 0>(din, dout, bps, ghz);
        run<8, 2, 0>(din, dout, bps, ghz);
        run<8, 4, 0>(din, dout, bps, ghz);
        run<8, 8, 0>(din, dout, bps, ghz);
        run<6, 6, 0>(din, dout, bps, ghz);
        run<4, 8, 0>(din, dout, bps, ghz);
        run<8, 16, 0>(din, dout, bps, ghz);
        run<8, 0, 2>(din, dout, bps, ghz);
        run<8, 8, 2>(din, dout, bps, ghz);
        run<6, 6, 2>(din, dout, bps, ghz);
    }
    return 0;
}
