// Experiment: what is the write-only HBM ceiling the cost-volume forward kernel should be held against?
// torch.fill_ of the same bytes runs at ~7.4 TB/s, cv_fwd_lean at ~6.7 TB/s.  Is the difference the constant data,
// the store flavour, the grid shape, or the access pattern (2 rows x 16 planes x 2 halves per item)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/write_peak.cu -o tools/write_peak && tools/write_peak
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

template <int MODE>   // 0 constant, 1 varying plain, 2 varying .cs
__global__ void __launch_bounds__(256) fill_kernel(float4* __restrict__ p, size_t n4) {
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
        float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
        if (MODE >= 1) { const float f = __uint_as_float(0x3f800000u | ((unsigned)(i * 2654435761u) >> 9)); v = make_float4(f, f + 1.f, f * 2.f, f - 3.f); }
        if (MODE == 2) __stcs(p + i, v); else p[i] = v;
    }
}

// cv-like geometry: volume [BC2][Df][Hf][Wf] floats; item = (bc, tile of R rows, chunk of 16 planes), a CTA writes
// planes d of its rows in the LEFT channel bc and the RIGHT channel bc + C*... (two streams 157 MB apart)
// PAT 0: the lean kernel's pattern (item order: bc, tile, dchunk fastest);  PAT 1: dchunk slowest inside bc
// (all CTAs on the same 16 planes);  PAT 2: linear items of the same byte count (contiguous 80 KB pieces)
template <int PAT>
__global__ void __launch_bounds__(256) pattern_kernel(float* __restrict__ out, int B, int C, int Df, int Hf, int Wf, int R,
                                                      unsigned int* __restrict__ ctr) {
    __shared__ int s_item[2];
    const int tid = threadIdx.x;
    const int n_tiles = (Hf + R - 1) / R, n_dch = Df / 16;
    const int n_items = B * C * n_tiles * n_dch;
    const size_t plane = (size_t)Hf * Wf;
    const int Wv = Wf / 4;
    if (tid == 0) s_item[0] = (int)atomicAdd(ctr, 1u);
    __syncthreads();
    int item = s_item[0];
    for (int k = 0; item < n_items; ++k) {
        if (tid == 0) s_item[(k + 1) & 1] = (int)atomicAdd(ctr, 1u);
        const float f = __uint_as_float(0x3f800000u | ((unsigned)((item * 256 + tid) * 2654435761u) >> 9));
        float4 v = make_float4(f, f + 1.f, f * 2.f, f - 3.f);
        if (PAT == 2) {
            // same bytes per item, contiguous: 2 halves x 16 planes x R rows x Wf floats
            const size_t per = (size_t)2 * 16 * R * Wf / 4;
            float4* p = reinterpret_cast<float4*>(out) + (size_t)item * per;
            for (size_t i = tid; i < per; i += 256) { __stcs(p + i, v); v.x += 1.f; }
        } else {
            int bc, tile, dc;
            if (PAT == 0) { dc = item % n_dch; const int it = item / n_dch; tile = it % n_tiles; bc = it / n_tiles; }
            else { tile = item % n_tiles; const int it = item / n_tiles; dc = it % n_dch; bc = it / n_dch; }
            const int b = bc / C, c = bc - b * C;
            float* outL = out + ((size_t)(b * 2 * C + c) * Df + dc * 16) * plane + (size_t)tile * R * Wf;
            float* outR = outL + (size_t)C * Df * plane;
            const int nv = R * Wv;                   // vectors per plane tile
            for (int d = 0; d < 16; ++d) {
                for (int i = tid; i < nv; i += 256) {
                    __stcs(reinterpret_cast<float4*>(outR + (size_t)d * plane) + i, v);
                    __stcs(reinterpret_cast<float4*>(outL + (size_t)d * plane) + i, v);
                }
                v.x += 1.f;
            }
        }
        __syncthreads();
        item = s_item[(k + 1) & 1];
        __syncthreads();
    }
}

int main() {
    const int B = 8, C = 12, Df = 64, Hf = 160, Wf = 320;
    const size_t n = (size_t)B * 2 * C * Df * Hf * Wf;     // floats
    float* buf;
    unsigned int* ctr;
    CK(cudaMalloc(&buf, n * 4));
    CK(cudaMalloc(&ctr, 4096));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    auto time_it = [&](const char* name, auto launch) -> int {
        std::vector<float> ts;
        for (int it = 0; it < 25; ++it) {
            CK(cudaMemsetAsync(ctr, 0, 4096));
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it >= 5) ts.push_back(ms);
        }
        CK(cudaGetLastError());
        std::sort(ts.begin(), ts.end());
        const float med = ts[ts.size() / 2];
        printf("{\"kernel\": \"%s\", \"ms_median\": %.4f, \"ms_best\": %.4f, \"TBps\": %.3f}\n", name, med, ts[0], n * 4 / med * 1e-9);
        return 0;
    };
    const size_t n4 = n / 4;
    for (int mult : {4, 8, 16, 0}) {
        const int grid = mult ? sms * mult : (int)((n4 + 255) / 256 / 4);
        char nm[96];
        snprintf(nm, 96, "fill_const grid=%d", grid);  time_it(nm, [&] { fill_kernel<0><<<grid, 256>>>((float4*)buf, n4); });
        snprintf(nm, 96, "fill_varying grid=%d", grid); time_it(nm, [&] { fill_kernel<1><<<grid, 256>>>((float4*)buf, n4); });
        snprintf(nm, 96, "fill_varying_cs grid=%d", grid); time_it(nm, [&] { fill_kernel<2><<<grid, 256>>>((float4*)buf, n4); });
    }
    time_it("cudaMemset", [&] { cudaMemsetAsync(buf, 0, n * 4); });
    for (int R : {2, 4, 8}) {
        for (int per_sm : {1, 2}) {
            char nm[96];
            snprintf(nm, 96, "pattern lean (bc,tile,dchunk) R=%d ctas/sm=%d", R, per_sm);
            time_it(nm, [&] { pattern_kernel<0><<<sms * per_sm, 256>>>(buf, B, C, Df, Hf, Wf, R, ctr); });
            snprintf(nm, 96, "pattern (bc,dchunk,tile) R=%d ctas/sm=%d", R, per_sm);
            time_it(nm, [&] { pattern_kernel<1><<<sms * per_sm, 256>>>(buf, B, C, Df, Hf, Wf, R, ctr); });
            snprintf(nm, 96, "pattern linear items R=%d ctas/sm=%d", R, per_sm);
            time_it(nm, [&] { pattern_kernel<2><<<sms * per_sm, 256>>>(buf, B, C, Df, Hf, Wf, R, ctr); });
        }
    }
    return 0;
}
