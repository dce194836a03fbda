// Experiment: what does the cost-volume BACKWARD (a pure read stream) pay for the dirty L2 lines the volume FORWARD (a pure
// write stream) leaves behind, and does the store flavour of the forward change it?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/dirty_l2.cu -o tools/dirty_l2 && tools/dirty_l2
#include <algorithm>
#include <vector>

#include "../rag_b200/csrc/cost_volume.cu"

namespace rag {
thread_local char g_last_error[512];
std::atomic<uint64_t> g_launches{0};
}  // namespace rag

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void spin_kernel(long long cycles) {   // stands for the head kernels between the two: FP32 work, no memory traffic
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
}

int main() {
    const int B = 4, C = 12, Df = 64, Hf = 96, Wf = 192;
    const size_t nf = (size_t)B * C * Hf * Wf, nv = (size_t)B * 2 * C * Df * Hf * Wf;
    float *x, *y, *cost, *g, *gx, *gy;
    unsigned int* ctr;
    CK(cudaMalloc(&x, nf * 4)); CK(cudaMalloc(&y, nf * 4)); CK(cudaMalloc(&cost, nv * 4)); CK(cudaMalloc(&g, nv * 4));
    CK(cudaMalloc(&gx, nf * 4)); CK(cudaMalloc(&gy, nf * 4)); CK(cudaMalloc(&ctr, 256));
    CK(cudaMemset(x, 0x3c, nf * 4)); CK(cudaMemset(y, 0x3d, nf * 4)); CK(cudaMemset(g, 0x3c, nv * 4));
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int Wv = Wf / 4, Df4 = (Df + 3) & ~3, R = 256 / Wv > Hf ? Hf : 256 / Wv;
    const size_t smem = (size_t)R * 16 * (Df4 + Wf + 4);
    const int n_tiles = (Hf + R - 1) / R, dchunk = 32, n_dch = (Df + dchunk - 1) / dchunk;
    const int n_items = B * C * n_tiles * n_dch;
    auto fwd = [&](int st) {
        cudaMemsetAsync(ctr, 0, 16);
        if (st == 0) rag::cv_fwd_lean_kernel<256, 1, true, 0><<<std::min(n_items, sms), 256, smem>>>(x, y, cost, B * C, C, Df, Hf, Wf, R, n_tiles, dchunk, n_dch, ctr);
        if (st == 1) rag::cv_fwd_lean_kernel<256, 1, true, 1><<<std::min(n_items, sms), 256, smem>>>(x, y, cost, B * C, C, Df, Hf, Wf, R, n_tiles, dchunk, n_dch, ctr);
        if (st == 2) rag::cv_fwd_lean_kernel<256, 1, true, 2><<<std::min(n_items, sms), 256, smem>>>(x, y, cost, B * C, C, Df, Hf, Wf, R, n_tiles, dchunk, n_dch, ctr);
    };
    auto bwd = [&]() {
        const int PV = Hf * Wv;
        rag::cv_bwd_v4_kernel<128, 4, 1><<<dim3((PV + 127) / 128, C, B), 128>>>(g, gx, gy, C, Df, Hf, Wf);
    };
    CK(cudaFuncSetAttribute(rag::cv_fwd_lean_kernel<256, 1, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(rag::cv_fwd_lean_kernel<256, 1, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(rag::cv_fwd_lean_kernel<256, 1, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto med = [](std::vector<float>& v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; };
    {   // backward alone, back to back
        std::vector<float> t;
        for (int it = 0; it < 25; ++it) {
            CK(cudaEventRecord(e0)); bwd(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it >= 5) t.push_back(ms);
        }
        printf("{\"case\": \"cv_bwd alone, back to back\", \"cv_bwd_ms\": %.4f}\n", med(t));
    }
    for (int st = 0; st < 3; ++st)
        for (long long gap : {0LL, 200000LL}) {   // 0 or ~0.1 ms of memory-silent work between forward and backward
            std::vector<float> tf, tb;
            for (int it = 0; it < 25; ++it) {
                CK(cudaEventRecord(e0)); fwd(st); CK(cudaEventRecord(e1));
                if (gap) spin_kernel<<<sms, 32>>>(gap);
                CK(cudaEventRecord(e2)); bwd();
                cudaEvent_t e3; CK(cudaEventCreate(&e3)); CK(cudaEventRecord(e3)); CK(cudaEventSynchronize(e3));
                float a, b2; CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b2, e2, e3));
                CK(cudaEventDestroy(e3));
                if (it >= 5) { tf.push_back(a); tb.push_back(b2); }
            }
            CK(cudaGetLastError());
            printf("{\"case\": \"cv_fwd (%s) -> %s -> cv_bwd\", \"cv_fwd_ms\": %.4f, \"cv_bwd_ms\": %.4f}\n",
                   st == 0 ? "st.global.cs" : st == 1 ? "st.global" : "st.global.wt", gap ? "0.1 ms of FP32-only work" : "nothing", med(tf), med(tb));
        }
    return 0;
}
