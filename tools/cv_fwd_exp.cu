// Experiment: which part of the product cost-volume forward kernel keeps it at 6.7-6.8 TB/s when a pure store
// stream of the same pattern reaches 7.3 TB/s (tools/write_peak.cu)?  The kernel below is a copy of
// cv_fwd_lean_kernel with switches (EXP bits): 1 = no global loads, 2 = no shared-memory staging / LDS (the right
// half stores the left vector), 4 = no mask logic, 8 = every item reads the same (L2-resident) feature rows,
// 16 = clustered L2 prefetch of the features one batch element ahead (32: with evict_last).
#include <algorithm>
#include <vector>

#include "../rag_b200/csrc/cv_lean.cuh"

namespace rag {
thread_local char g_last_error[512];
std::atomic<uint64_t> g_launches{0};
#include "cv_fwd_exp_kernel.inc"
}  // namespace rag

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    int B = 8, C = 12, Df = 64, Hf = 160, Wf = 320;
    if (argc > 1 && atoi(argv[1]) == 1) { B = 4; Hf = 96; Wf = 192; }
    if (argc > 1 && atoi(argv[1]) == 2) { B = 32; }
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
    auto run = [&](auto kern, int NT, int VPT, int per_sm, const char* name) -> int {
        const int Wv = Wf / 4, Df4 = (Df + 3) & ~3, dchunk = 16;
        const size_t row_bytes = (size_t)16 * (Df4 + Wf + 4);
        int R = std::min(Hf, (NT * VPT) / Wv);
        const size_t smem = R * row_bytes;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int n_tiles = (Hf + R - 1) / R, n_dch = (Df + dchunk - 1) / dchunk;
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
        printf("{\"kernel\": \"%s\", \"ctas_per_sm\": %d, \"ms_median\": %.4f, \"ms_best\": %.4f, \"TBps\": %.3f}\n", name, per_sm,
               ts[ts.size() / 2], ts[0], alg / ts[ts.size() / 2] * 1e-9);
        return 0;
    };
    for (int per_sm : {1, 2}) {
        run(rag::cv_fwd_exp_kernel<256, 2, true, 0>, 256, 2, per_sm, "product");
        run(rag::cv_fwd_exp_kernel<256, 2, true, 1>, 256, 2, per_sm, "no global loads");
        run(rag::cv_fwd_exp_kernel<256, 2, true, 8>, 256, 2, per_sm, "L2-resident loads");
        run(rag::cv_fwd_exp_kernel<256, 2, true, 2>, 256, 2, per_sm, "no smem staging");
        run(rag::cv_fwd_exp_kernel<256, 2, true, 4>, 256, 2, per_sm, "no masks");
        run(rag::cv_fwd_exp_kernel<256, 2, true, 3>, 256, 2, per_sm, "no loads, no smem");
        run(rag::cv_fwd_exp_kernel<256, 2, true, 7>, 256, 2, per_sm, "no loads, no smem, no masks");
        run(rag::cv_fwd_exp_kernel<256, 2, true, 7, 1>, 256, 2, per_sm, "no loads, no smem, no masks, plain st");
        run(rag::cv_fwd_exp_kernel<256, 2, true, 16>, 256, 2, per_sm, "prefetch bursts");
        run(rag::cv_fwd_exp_kernel<256, 2, true, 48>, 256, 2, per_sm, "prefetch bursts evict_last");
        run(rag::cv_fwd_exp_kernel<512, 1, true, 0>, 512, 1, per_sm, "product 512x1");
        run(rag::cv_fwd_exp_kernel<512, 1, true, 16>, 512, 1, per_sm, "512x1 prefetch bursts");
        run(rag::cv_fwd_exp_kernel<512, 1, true, 48>, 512, 1, per_sm, "512x1 prefetch bursts evict_last");
    }
    return 0;
}
