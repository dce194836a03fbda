// x3 disparity-head forward, "column" version: one thread per 3x1 output-pixel column
// (rows 3r+1..3r+3 of one output column w).  Same arithmetic as disp_head_x3.cuh, a third of the
// per-thread state: ~64 registers -> 32 resident warps per SM instead of 16, which is what the SFU
// needs to stay busy (measured: 4 warps/SMSP leave the MUFU pipe ~65% utilised).
// Lanes of a warp are consecutive output columns, so every global access is coalesced and there are
// no partially filled warps for the image widths of the reference (W = 3*Wl is a multiple of 32 for
// 576, 960 and 1248).  Totals are (hi,lo) fp32 pairs folded with an error-free TwoSum: no FP64, no
// F2F on the SFU pipe.
#pragma once
#include "disp_head_x3.cuh"

namespace rag {

__device__ __forceinline__ void two_sum_f(float& hi, float& lo, float g) {
    const float s = hi + g;
    const float bb = s - hi;
    lo += (hi - (s - bb)) + (g - bb);
    hi = s;
}

// grid: x = ceil((Hl+1)*W / NT), y = B.  smem: float lam1[D].
template <int NT>
__global__ void __launch_bounds__(NT, 1024 / NT)
head_fwd_x3c_kernel(const float* __restrict__ cost, float* __restrict__ disp, float* __restrict__ stats,
                    int Dl, int Hl, int Wl, float scale) {
    extern __shared__ float x3c_lam[];   // [D] lambda1 of full-res bin k
    const int D = 3 * Dl, H = 3 * Hl, W = 3 * Wl;
    for (int k = threadIdx.x; k < D; k += NT) {
        int t0, t1;
        float l0, l1;
        src_index<true>(scale, k, Dl, t0, t1, l0, l1);
        x3c_lam[k] = l1;
    }
    __syncthreads();
    const int id = blockIdx.x * NT + threadIdx.x;
    if (id >= (Hl + 1) * W) return;
    const int rr = id / W;
    const int w = id - rr * W;
    const int r = rr - 1;
    const int b = blockIdx.y;

    X3Axis ah;
    x3_axis(scale, r, Hl, ah);
    float hs0[3], hs1[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { hs0[i] = ah.l0[i] * kX3NegLog2e; hs1[i] = ah.l1[i] * kX3NegLog2e; }
    int c0, c1;
    float wl0, wl1;
    src_index<true>(scale, w, Wl, c0, c1, wl0, wl1);
    const bool two = c1 != c0;           // false only on the clamped last columns

    const size_t plane = (size_t)Hl * Wl;
    const float* p0 = cost + (size_t)b * Dl * plane + (size_t)ah.lo0 * Wl + c0;
    const float* p1 = cost + (size_t)b * Dl * plane + (size_t)ah.lo1 * Wl + c0;
    float v0, v1, v2, v3;
    auto fetch = [&]() {
        v0 = __ldg(p0); v2 = __ldg(p1);
        v1 = two ? __ldg(p0 + 1) : v0;
        v3 = two ? __ldg(p1 + 1) : v2;
    };
    int jl = 0;                          // low-res bin the pointers sit on
    auto next = [&]() {
        if (jl < Dl - 1) { ++jl; p0 += plane; p1 += plane; }
        fetch();
    };

    float a[3], m[3], dg[3], ng[3], t[3];
    float dhi[3], dlo[3], nhi[3], nlo[3];
    const float kc = 0.5f * (float)D;
    fetch();
    {
        const float x0 = __fmaf_rn(wl0, v0, wl1 * v1), x1 = __fmaf_rn(wl0, v2, wl1 * v3);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            m[i] = __fmaf_rn(hs0[i], x0, hs1[i] * x1);  // reference exponent = exponent of low-res bin 0
            a[i] = 0.f;
            dg[i] = 1.f;                                // full-res bin 0 (lambda1 == 0): 2^0 ...
            ng[i] = -kc;                                //   ... times (0 - kc)
            dhi[i] = dlo[i] = nhi[i] = nlo[i] = 0.f;
        }
    }
    next();
    float kf = 1.f - kc;

#pragma unroll 2
    for (int j = 0; j < Dl - 1; ++j) {
        const float x0 = __fmaf_rn(wl0, v0, wl1 * v1), x1 = __fmaf_rn(wl0, v2, wl1 * v3);
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = __fmaf_rn(hs0[i], x0, __fmaf_rn(hs1[i], x1, -m[i]));
        next();                                          // prefetch bin min(j+2, Dl-1)
        const float l1 = x3c_lam[3 * j + 1], l2 = x3c_lam[3 * j + 2], l3 = x3c_lam[3 * j + 3];
        if (fmaxf(fmaxf(t[0], t[1]), t[2]) > kX3Tau) {  // rare: move the reference exponent(s) up
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (t[i] > kX3Tau) {
                    const float f = ex2_approx(-t[i]);
                    dg[i] *= f; ng[i] *= f; dhi[i] *= f; dlo[i] *= f; nhi[i] *= f; nlo[i] *= f;
                    m[i] += t[i]; a[i] -= t[i]; t[i] = 0.f;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float dlt = t[i] - a[i];
            const float e1 = ex2_approx(__fmaf_rn(l1, dlt, a[i]));
            const float e2 = ex2_approx(__fmaf_rn(l2, dlt, a[i]));
            const float e3 = ex2_approx(__fmaf_rn(l3, dlt, a[i]));
            dg[i] += e1; ng[i] = __fmaf_rn(e1, kf, ng[i]);
            dg[i] += e2; ng[i] = __fmaf_rn(e2, kf + 1.f, ng[i]);
            dg[i] += e3; ng[i] = __fmaf_rn(e3, kf + 2.f, ng[i]);
            a[i] = t[i];
        }
        kf += 3.f;
        if ((j & 7) == 7) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                two_sum_f(dhi[i], dlo[i], dg[i]); two_sum_f(nhi[i], nlo[i], ng[i]);
                dg[i] = 0.f; ng[i] = 0.f;
            }
        }
    }
    const size_t img = (size_t)H * W;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        // last k-block (j = Dl-1): bins 3Dl-2 and 3Dl-1 both sit on low-res bin Dl-1
        const float e = ex2_approx(a[i]);
        dg[i] += e + e;
        ng[i] = __fmaf_rn(e, kf, ng[i]);
        ng[i] = __fmaf_rn(e, kf + 1.f, ng[i]);
        two_sum_f(dhi[i], dlo[i], dg[i]); two_sum_f(nhi[i], nlo[i], ng[i]);
        if (ah.valid[i]) {
            const size_t o = (size_t)ah.idx[i] * W + w;
            const float inv = 1.f / (dhi[i] + dlo[i]);
            const float q = nhi[i] * inv;                                   // num/den, low parts to first order
            const float rr2 = __fmaf_rn(-q, dhi[i], nhi[i]) + (nlo[i] - q * dlo[i]);
            disp[(size_t)b * img + o] = kc + (q + rr2 * inv);
            if (stats) {
                stats[(size_t)b * 2 * img + o] = m[i];
                stats[(size_t)b * 2 * img + img + o] = inv;
            }
        }
    }
}

}  // namespace rag
