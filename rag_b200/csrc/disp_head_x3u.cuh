// x3 disparity-head forward, tiled + lean addressing (default).  Needs Wl % 4 == 0, 16B-aligned cost_lr.
//
// Same plan as disp_head_x3t.cuh (centred 3x3 blocks, 4x32 blocks per CTA, 3-stage cp.async pipeline
// over 8-bin chunks) with the per-bin instruction count cut to the bone, because with MUFU.EX2 in
// the mix an SM sub-partition sustains only ~0.75 instructions/clk (tools/mufu_mix2.cu): the
// clamp-to-edge columns are patched into the shared-memory window once per chunk, so every thread
// reads its 3x3 low-res neighbourhood at COMPILE-TIME offsets from one pointer that is bumped per
// bin (9 LDS, no address arithmetic), and the three bin lambdas come from one 16-byte LDS.
#pragma once
#include <cuda_pipeline.h>

#include "disp_head_x3t.cuh"

namespace rag {

// grid: x = ceil(Wl/32), y = ceil(Hl/4), z = B; 128 threads.
// smem: tile[kTStages][kTBins][kTRows][kTCols] | float4 lam[Dl] (l(3j+1), l(3j+2), l(3j+3), -)
__global__ void __launch_bounds__(128, 4)
head_fwd_x3u_kernel(const float* __restrict__ cost, float* __restrict__ disp, float* __restrict__ stats,
                    int Dl, int Hl, int Wl, float scale) {
    extern __shared__ __align__(16) float x3u_smem[];
    const int D = 3 * Dl, W = 3 * Wl;
    float* tile = x3u_smem;
    float4* lam = reinterpret_cast<float4*>(x3u_smem + kTStages * kTStageFloats);   // [Dl]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int C0 = blockIdx.x * 32, R0 = blockIdx.y * 4;
    const int T0 = C0 - 4;                       // image column held in tile column 0
    const size_t plane = (size_t)Hl * Wl;
    const float* base = cost + (size_t)b * Dl * plane;
    const int Wv = Wl >> 2;
    const int n_chunks = (Dl + kTBins - 1) / kTBins;
    // clamp-to-edge patches: tile column 3 is image column -1 in the first strip; tile column Wl-T0 is
    // image column Wl in the strip that contains the last image column
    const bool patch_l = C0 == 0;
    const bool patch_r = (Wl - T0) <= 36;
    const int pr = Wl - T0;

    auto issue_chunk = [&](int ch) {
        if (ch < n_chunks) {
            float* dst = tile + (ch % kTStages) * kTStageFloats;
            for (int u = tid; u < kTBins * kTRows * (kTCols / 4); u += 128) {
                const int vec = u % (kTCols / 4);
                const int row = (u / (kTCols / 4)) % kTRows;
                const int bin = u / ((kTCols / 4) * kTRows);
                const int gj = min(ch * kTBins + bin, Dl - 1);
                const int gr = min(max(R0 - 1 + row, 0), Hl - 1);
                const int gv = min(max((T0 >> 2) + vec, 0), Wv - 1);
                __pipeline_memcpy_async(dst + (bin * kTRows + row) * kTCols + vec * 4,
                                        base + (size_t)gj * plane + (size_t)gr * Wl + gv * 4, 16);
            }
        }
        __pipeline_commit();
    };
    issue_chunk(0);
    issue_chunk(1);
    for (int j = tid; j < Dl; j += 128) {
        float l[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            int t0, t1;
            float l0;
            src_index<true>(scale, min(3 * j + 1 + q, D - 1), Dl, t0, t1, l0, l[q]);
        }
        lam[j] = make_float4(l[0], l[1], l[2], 0.f);
    }

    // ---- per-thread geometry: block (r,c); threads past the ragged edge shadow the last valid block ----
    const int r_raw = R0 + warp, c_raw = C0 + lane;
    const bool active = r_raw < Hl && c_raw < Wl;
    const int r = min(r_raw, Hl - 1), c = min(c_raw, Wl - 1);
    float hs0[3], hs1[3], wl0[3], wl1[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int i0, i1;
        float l0, l1;
        src_index<true>(scale, 3 * r + i, Hl, i0, i1, l0, l1);
        hs0[i] = l0 * kX3NegLog2e; hs1[i] = l1 * kX3NegLog2e;
        src_index<true>(scale, 3 * c + i, Wl, i0, i1, wl0[i], wl1[i]);
    }
    // window origin of the thread inside a bin slab: tile row (r - R0) holds image row r-1 (clamped at
    // fill time), tile column (c - 1 - T0) = (c - C0 + 3) holds image column c-1 (patched at the edges)
    const int org = (r - R0) * kTCols + (c - C0 + 3);

    float a[9], m[9], dg[9], ng[9], t[9];
    float dhi[9], dlo[9], nhi[9], nlo[9];
    const float kc = 0.5f * (float)D;
    float kf1 = 1.f - kc, kf2 = 2.f - kc, kf3 = 3.f - kc;

    // t = (blend of the bin at p) - mm
    auto blend = [&](const float* p, const float (&mm)[9], float (&o)[9]) {
        float x[3][3];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
            const float v0 = p[rr * kTCols + 0], v1 = p[rr * kTCols + 1], v2 = p[rr * kTCols + 2];
            x[rr][0] = __fmaf_rn(wl0[0], v0, wl1[0] * v1);
            x[rr][1] = __fmaf_rn(wl0[1], v1, wl1[1] * v2);
            x[rr][2] = __fmaf_rn(wl0[2], v1, wl1[2] * v2);
        }
#pragma unroll
        for (int pw = 0; pw < 3; ++pw) {
            o[0 * 3 + pw] = __fmaf_rn(hs0[0], x[0][pw], __fmaf_rn(hs1[0], x[1][pw], -mm[0 * 3 + pw]));
            o[1 * 3 + pw] = __fmaf_rn(hs0[1], x[1][pw], __fmaf_rn(hs1[1], x[2][pw], -mm[1 * 3 + pw]));
            o[2 * 3 + pw] = __fmaf_rn(hs0[2], x[1][pw], __fmaf_rn(hs1[2], x[2][pw], -mm[2 * 3 + pw]));
        }
    };

    for (int ch = 0; ch < n_chunks; ++ch) {
        issue_chunk(ch + 2);
        __pipeline_wait_prior(2);
        __syncthreads();
        float* stw = tile + (ch % kTStages) * kTStageFloats;
        if (patch_l || patch_r) {                // uniform per CTA: only the first/last strip
            for (int u = tid; u < kTBins * kTRows; u += 128) {
                float* rowp = stw + u * kTCols;
                if (patch_l) rowp[3] = rowp[4];
                if (patch_r) rowp[pr] = rowp[pr - 1];
            }
            __syncthreads();
        }
        const int jbeg = ch * kTBins, jend = min(jbeg + kTBins, Dl);
        const float* p = stw + org;
        if (ch == 0) {
            // low-res bin 0 defines the reference exponent; full-res bin 0 sits exactly on it
#pragma unroll
            for (int i = 0; i < 9; ++i) m[i] = 0.f;
            blend(p, m, t);
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                m[i] = t[i]; a[i] = 0.f; dg[i] = 1.f; ng[i] = -kc;
                dhi[i] = dlo[i] = nhi[i] = nlo[i] = 0.f;
            }
            p += kTRows * kTCols;
        }
        const int n_it = jend - max(jbeg, 1);
        const float4* lp = lam + max(jbeg, 1) - 1;
#pragma unroll 2
        for (int it = 0; it < n_it; ++it) {
            blend(p, m, t);                      // exponent of the k-block's upper low-res bin, relative to m
            p += kTRows * kTCols;
            const float4 l = *lp++;
            float mx = t[0];
#pragma unroll
            for (int i = 1; i < 9; ++i) mx = fmaxf(mx, t[i]);
            if (mx > kX3Tau) {                   // rare: some pixel's running maximum ran away from its reference
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    if (t[i] > kX3Tau) {
                        const float f = ex2_approx(-t[i]);
                        dg[i] *= f; ng[i] *= f; dhi[i] *= f; dlo[i] *= f; nhi[i] *= f; nlo[i] *= f;
                        m[i] += t[i]; a[i] -= t[i]; t[i] = 0.f;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const float dlt = t[i] - a[i];
                const float e1 = ex2_approx(__fmaf_rn(l.x, dlt, a[i]));
                const float e2 = ex2_approx(__fmaf_rn(l.y, dlt, a[i]));
                const float e3 = ex2_approx(__fmaf_rn(l.z, dlt, a[i]));
                dg[i] += e1; ng[i] = __fmaf_rn(e1, kf1, ng[i]);
                dg[i] += e2; ng[i] = __fmaf_rn(e2, kf2, ng[i]);
                dg[i] += e3; ng[i] = __fmaf_rn(e3, kf3, ng[i]);
                a[i] = t[i];
            }
            kf1 += 3.f; kf2 += 3.f; kf3 += 3.f;
        }
        // fold the short fp32 group sums (<= 24 bins) into the compensated totals once per chunk
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            two_sum_f(dhi[i], dlo[i], dg[i]); two_sum_f(nhi[i], nlo[i], ng[i]);
            dg[i] = 0.f; ng[i] = 0.f;
        }
        __syncthreads();                         // the stage may be refilled from here on
    }
    if (!active) return;
    const size_t img = (size_t)3 * Hl * W;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        // last k-block (j = Dl-1): bins 3Dl-2 and 3Dl-1 both sit on low-res bin Dl-1
        const float e = ex2_approx(a[i]);
        float g = e + e, n = __fmaf_rn(e, kf2, e * kf1);
        two_sum_f(dhi[i], dlo[i], g); two_sum_f(nhi[i], nlo[i], n);
        const int ph = i / 3, pw = i - 3 * ph;
        const size_t o = (size_t)(3 * r + ph) * W + (3 * c + pw);
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

}  // namespace rag
