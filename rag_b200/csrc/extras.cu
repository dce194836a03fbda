// Adjacent rows of the hot path: the masked loss + metric sums
// that consume the disparity map, and the eval-time input staging in front of the Feature Net.
#include "common.cuh"

namespace rag {

// ---------------------------------------------------------------------------------------------
// masked smooth-L1 + EPE / D1 / Thres1-3 sums, per image, deterministic two-stage reduction.
// Reference: src/approaches/rag.py:418-430, src/utilstool/metrics.py:22-65.
// ---------------------------------------------------------------------------------------------
constexpr int kLmNT = 256;
constexpr int kLmItems = 16;  // elements per thread

__host__ __device__ inline int lm_blocks(int H, int W) { return (H * W + kLmNT * kLmItems - 1) / (kLmNT * kLmItems); }

__global__ void __launch_bounds__(kLmNT)
loss_metrics_partial_kernel(const float* __restrict__ est, const float* __restrict__ gt, double* __restrict__ scratch,
                            int HW, float maxdisp) {
    const int b = blockIdx.y;
    const float* e = est + (size_t)b * HW;
    const float* g = gt + (size_t)b * HW;
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int start = blockIdx.x * kLmNT * kLmItems;
#pragma unroll 4
    for (int i = 0; i < kLmItems; ++i) {
        const int idx = start + i * kLmNT + threadIdx.x;
        if (idx < HW) {
            const float gv = __ldg(g + idx), ev = __ldg(e + idx);
            if (gv > 0.f) s[1] += 1.0;
            if (gv > 0.f && gv < maxdisp) {
                const float d = ev - gv;
                const float E = fabsf(gv - ev);
                const float ad = fabsf(d);
                s[0] += 1.0;
                s[2] += (double)(ad < 1.f ? 0.5f * d * d : ad - 0.5f);
                s[3] += (double)E;
                if (E > 3.f && E / fabsf(gv) > 0.05f) s[4] += 1.0;
                if (E > 1.f) s[5] += 1.0;
                if (E > 2.f) s[6] += 1.0;
                if (E > 3.f) s[7] += 1.0;
            }
        }
    }
    // fixed-order block reduction: warp shuffle tree, then warp 0 adds the 8 warp results in order
    __shared__ double sh[8][kLmNT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        double v = s[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sh[q][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double v = 0.0;
        for (int wi = 0; wi < kLmNT / 32; ++wi) v += sh[threadIdx.x][wi];
        scratch[((size_t)b * gridDim.x + blockIdx.x) * 8 + threadIdx.x] = v;
    }
}

__global__ void loss_metrics_final_kernel(const double* __restrict__ scratch, double* __restrict__ sums, int nblk) {
    const int b = blockIdx.x, q = threadIdx.x;
    if (q >= 8) return;
    double v = 0.0;
    for (int i = 0; i < nblk; ++i) v += scratch[((size_t)b * nblk + i) * 8 + q];
    sums[(size_t)b * 8 + q] = v;
}

__global__ void __launch_bounds__(256)
smooth_l1_bwd_kernel(const float* __restrict__ est, const float* __restrict__ gt, const double* __restrict__ sums,
                     const float* __restrict__ gloss, float* __restrict__ gest, int B, size_t n, float maxdisp) {
    double nm = 0.0;
    for (int b = 0; b < B; ++b) nm += sums[(size_t)b * 8];
    const float scale = __ldg(gloss) / (float)nm;
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= n) return;
    const float gv = __ldg(gt + idx);
    float out = 0.f;
    if (gv > 0.f && gv < maxdisp) {
        const float d = __ldg(est + idx) - gv;
        out = (fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f)) * scale;
    }
    gest[idx] = out;
}

int loss_metrics_scratch(int H, int W) { return (H > 0 && W > 0) ? lm_blocks(H, W) * 8 : 0; }

int loss_metrics_sums(const float* est, const float* gt, double* sums, double* scratch, int B, int H, int W,
                      float maxdisp, cudaStream_t st) {
    if (!est || !gt || !sums || !scratch) return fail(RAG_E_NULL, "loss_metrics_sums: null pointer");
    if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || (size_t)H * W >= ((size_t)1 << 31))
        return fail(RAG_E_SHAPE, "loss_metrics_sums: bad shape B=%d H=%d W=%d", B, H, W);
    if (!aligned(sums, 8) || !aligned(scratch, 8)) return fail(RAG_E_ALIGN, "loss_metrics_sums: sums/scratch must be 8-byte aligned");
    const int nblk = lm_blocks(H, W);
    loss_metrics_partial_kernel<<<dim3(nblk, B), kLmNT, 0, st>>>(est, gt, scratch, H * W, maxdisp);
    if (int e = check_launch("loss_metrics_partial")) return e;
    loss_metrics_final_kernel<<<B, 32, 0, st>>>(scratch, sums, nblk);
    return check_launch("loss_metrics_final");
}

int smooth_l1_bwd(const float* est, const float* gt, const double* sums, const float* gloss, float* gest,
                  int B, int H, int W, float maxdisp, cudaStream_t st) {
    if (!est || !gt || !sums || !gloss || !gest) return fail(RAG_E_NULL, "smooth_l1_bwd: null pointer");
    if (B <= 0 || H <= 0 || W <= 0) return fail(RAG_E_SHAPE, "smooth_l1_bwd: bad shape B=%d H=%d W=%d", B, H, W);
    const size_t n = (size_t)B * H * W;
    smooth_l1_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(est, gt, sums, gloss, gest, B, n, maxdisp);
    return check_launch("smooth_l1_bwd");
}

// ---------------------------------------------------------------------------------------------
// uint8 HWC -> normalised fp32 CHW with top/right zero padding.
// Reference: src/dataloaders/data_io.py:6-13 + src/dataloaders/stereo_dataset.py:88-102.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
normalize_pad_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int H, int W, int top, int right) {
    const int Ho = H + top, Wo = W + right;
    const size_t n = (size_t)3 * Ho * Wo;
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= n) return;
    const int b = blockIdx.y;
    const int wo = idx % Wo, ho = (idx / Wo) % Ho, ch = idx / ((size_t)Wo * Ho);
    float v = 0.f;
    if (ho >= top && wo < W) {
        const float mean = ch == 0 ? 0.485f : ch == 1 ? 0.456f : 0.406f;
        const float stdv = ch == 0 ? 0.229f : ch == 1 ? 0.224f : 0.225f;
        const float p = (float)img[(((size_t)b * H + (ho - top)) * W + wo) * 3 + ch];
        v = __fdiv_rn(__fsub_rn(__fdiv_rn(p, 255.f), mean), stdv);
    }
    out[(size_t)b * n + idx] = v;
}

int normalize_pad(const uint8_t* img, float* out, int B, int H, int W, int top, int right, cudaStream_t st) {
    if (!img || !out) return fail(RAG_E_NULL, "normalize_pad: null pointer");
    if (B <= 0 || H <= 0 || W <= 0 || top < 0 || right < 0 || B > 65535)
        return fail(RAG_E_SHAPE, "normalize_pad: bad shape B=%d H=%d W=%d top=%d right=%d", B, H, W, top, right);
    const size_t n = (size_t)3 * (H + top) * (W + right);
    normalize_pad_kernel<<<dim3((unsigned)((n + 255) / 256), B), 256, 0, st>>>(img, out, H, W, top, right);
    return check_launch("normalize_pad");
}

}  // namespace rag
