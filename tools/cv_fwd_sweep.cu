// Experiment: launch geometry of the product cost-volume forward kernel (cv_fwd_lean_kernel, included from the
// library sources unchanged): threads x vectors per thread, CTAs per SM, disparities per item.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/cv_fwd_sweep.cu -o tools/cv_fwd_sweep
#include <algorithm>
#include <vector>

#include "../rag_b200/csrc/cv_lean.cuh"

namespace rag {
thread_local char g_last_error[512];
std::atomic<uint64_t> g_launches{0};
}  // namespace rag

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    int B = 8, C = 12, Df = 64, Hf = 160, Wf = 320;
    if (argc > 1 && atoi(argv[1]) == 1) { B = 4; Hf = 96; Wf = 192; }
    if (argc > 1 && atoi(argv[1]) == 2) { B = 4; Hf = 128; Wf = 416; }
    if (argc > 1 && atoi(argv[1]) == 3) { B = 4; Hf = 128; Wf = 416; Df = 96; }
    const size_t nf = (size_t)B * C * Hf * Wf, nv = (size_t)B * 2 * C * Df * Hf * Wf;
    float *x, *y, *cost;
    unsigned int* ctr;
    CK(cudaMalloc(&x, nf * 4)); CK(cudaMalloc(&y, nf * 4)); CK(cudaMalloc(&cost, nv * 4)); CK(cudaMalloc(&ctr, 256));
    CK(cudaMemset(x, 0x3c, nf * 4)); CK(cudaMemset(y, 0x3d, nf * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t alg = nv * 4 + 2 * nf * 4;

    auto run = [&](auto kern, int NT, int VPT, int per_sm, int dchunk, const char* name) -> int {
        const int Wv = Wf / 4, Df4 = (Df + 3) & ~3;
        const size_t row_bytes = (size_t)16 * (Df4 + Wf + 4);
        int R = std::min(Hf, (NT * VPT) / Wv);
        const size_t smem = R * row_bytes;
        if (smem > 200 * 1024 || R < 1) return 0;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int n_tiles = (Hf + R - 1) / R;
        const int n_dch = (Df + dchunk - 1) / dchunk;
        const long long n_items = (long long)B * C * n_tiles * n_dch;
        const int grid = (int)std::min<long long>(n_items, (long long)sms * per_sm);
        std::vector<float> ts;
        for (int it = 0; it < 25; ++it) {
            CK(cudaMemsetAsync(ctr, 0, 16));
            CK(cudaEventRecord(e0));
            kern<<<grid, NT, smem>>>(x, y, cost, B * C, C, Df, Hf, Wf, R, n_tiles, dchunk, n_dch, ctr);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it >= 5) ts.push_back(ms);
        }
        CK(cudaGetLastError());
        std::sort(ts.begin(), ts.end());
        printf("{\"kernel\": \"%s\", \"R\": %d, \"ctas_per_sm\": %d, \"dchunk\": %d, \"ms_median\": %.4f, \"ms_best\": %.4f, \"TBps\": %.3f}\n",
               name, R, per_sm, dchunk, ts[ts.size() / 2], ts[0], alg / ts[ts.size() / 2] * 1e-9);
        return 0;
    };
    for (int dchunk : {16, 32, 64})
        for (int per_sm : {1, 2, 3}) {
            run(rag::cv_fwd_lean_kernel<256, 2, true>, 256, 2, per_sm, dchunk, "lean<256,2>");
            run(rag::cv_fwd_lean_kernel<512, 1, true>, 512, 1, per_sm, dchunk, "lean<512,1>");
            run(rag::cv_fwd_lean_kernel<128, 2, true>, 128, 2, per_sm, dchunk, "lean<128,2>");
            run(rag::cv_fwd_lean_kernel<256, 1, true>, 256, 1, per_sm, dchunk, "lean<256,1>");
            run(rag::cv_fwd_lean_kernel<256, 4, true>, 256, 4, per_sm, dchunk, "lean<256,4>");
            run(rag::cv_fwd_lean_kernel<512, 2, true>, 512, 2, per_sm, dchunk, "lean<512,2>");
        }
    return 0;
}
