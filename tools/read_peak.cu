// Experiment: the read-only HBM ceiling the cost-volume backward should be held against (companion of write_peak.cu).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/read_peak.cu -o tools/read_peak && tools/read_peak
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

// linear stream: every thread sums float4s with a grid stride, U loads in flight
template <int U, int MODE>   // MODE 0 plain, 1 ld.global.cs, 2 __ldg
__global__ void __launch_bounds__(256) read_linear(const float4* __restrict__ p, size_t n4, float* __restrict__ sink) {
    const size_t stride = (size_t)gridDim.x * 256;
    float acc = 0.f;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = MODE == 1 ? __ldcs(p + i + u * stride) : MODE == 2 ? __ldg(p + i + u * stride) : p[i + u * stride];
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < n4; i += stride) { const float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) sink[0] = acc;
}

// the backward's pattern: volume [BC2][Df][PV] float4; a thread owns vector p of a (b, c) plane and walks the Df planes of the
// LEFT channel and of the RIGHT channel (vector p + q), U disparities in flight; FULL = read everything (no triangle skip)
template <int U>
__global__ void __launch_bounds__(128) read_cvbwd(const float4* __restrict__ g, int C, int Df, int PV, int Wv, float* __restrict__ sink) {
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= PV) return;
    const int c = blockIdx.y, b = blockIdx.z;
    const float4* gl = g + (size_t)(b * 2 * C + c) * Df * PV + p;
    const float4* gr = gl + (size_t)C * Df * PV;
    float acc = 0.f;
    for (int d0 = Df - U; d0 >= 0; d0 -= U) {
        float4 a[U], r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int d = d0 + u, q = d >> 2;
            a[u] = __ldcs(gl + (size_t)d * PV);
            r[u] = (p + q < PV) ? __ldg(gr + (size_t)d * PV + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += a[u].x + a[u].w + r[u].y + r[u].z;
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main() {
    const int B = 8, C = 12, Df = 64, Hf = 160, Wf = 320, Wv = Wf / 4, PV = Hf * Wv;
    const size_t n = (size_t)B * 2 * C * Df * Hf * Wf;
    float *buf, *sink;
    CK(cudaMalloc(&buf, n * 4)); CK(cudaMalloc(&sink, 16));
    CK(cudaMemset(buf, 0x3c, n * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    auto time_it = [&](const char* name, auto launch) -> int {
        std::vector<float> ts;
        for (int it = 0; it < 25; ++it) {
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it >= 5) ts.push_back(ms);
        }
        CK(cudaGetLastError());
        std::sort(ts.begin(), ts.end());
        printf("{\"kernel\": \"%s\", \"ms_median\": %.4f, \"ms_best\": %.4f, \"TBps\": %.3f}\n", name, ts[ts.size() / 2], ts[0], n * 4 / ts[ts.size() / 2] * 1e-9);
        return 0;
    };
    const size_t n4 = n / 4;
    char nm[96];
    for (int mult : {8, 16, 32}) {
        snprintf(nm, 96, "linear U=4 plain grid=%dxSM", mult); time_it(nm, [&] { read_linear<4, 0><<<sms * mult, 256>>>((const float4*)buf, n4, sink); });
        snprintf(nm, 96, "linear U=4 ld.cs grid=%dxSM", mult); time_it(nm, [&] { read_linear<4, 1><<<sms * mult, 256>>>((const float4*)buf, n4, sink); });
        snprintf(nm, 96, "linear U=8 ld.cs grid=%dxSM", mult); time_it(nm, [&] { read_linear<8, 1><<<sms * mult, 256>>>((const float4*)buf, n4, sink); });
        snprintf(nm, 96, "linear U=8 ldg grid=%dxSM", mult); time_it(nm, [&] { read_linear<8, 2><<<sms * mult, 256>>>((const float4*)buf, n4, sink); });
    }
    dim3 grid((PV + 127) / 128, C, B);
    time_it("cv_bwd pattern, full volume, U=4", [&] { read_cvbwd<4><<<grid, 128>>>((const float4*)buf, C, Df, PV, Wv, sink); });
    time_it("cv_bwd pattern, full volume, U=8", [&] { read_cvbwd<8><<<grid, 128>>>((const float4*)buf, C, Df, PV, Wv, sink); });
    return 0;
}
