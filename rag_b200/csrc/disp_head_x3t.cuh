// x3 disparity-head forward, tiled version (default).  Needs Wl % 4 == 0 and 16-byte aligned cost_lr.
//
// What the profiles of the first versions showed (profiles/r1_head_fwd_variants.md): the SFU
// (MUFU.EX2, 8 clk per warp instruction per SM sub-partition) is the busiest pipe, but it idled
// ~30% of the time behind global-load latency (cost_lr bins are 200 KB apart, every k-block
// touches new lines) and 9-16% of all warps were almost empty (the "+1" block column).  So:
//   * blocks are CENTRED: block (r,c) owns output rows 3r..3r+2 and cols 3c..3c+2 and interpolates
//     within low-res rows {r-1,r,r+1} / cols {c-1,c,c+1} (clamped).  There are exactly Hl x Wl
//     blocks -- no partially filled warps when Wl is a multiple of 32 (192, 320, 416 are).
//   * a CTA (4 warps = 4 block rows x 32 block cols) streams its 6 x 40 low-res window through
//     shared memory with a 3-stage cp.async (LDGSTS.128) pipeline, 8 bins per stage, so the inner
//     loop only ever waits on shared memory.
// Everything else (fp32 lambda replication, lazy per-pixel reference exponent, grouped + compensated
// sums, centred regression) is as in disp_head_x3.cuh.
#pragma once
#include <cuda_pipeline.h>

#include "disp_head_x3c.cuh"

namespace rag {

constexpr int kTRows = 6, kTCols = 40, kTBins = 8, kTStages = 3;
constexpr int kTStageFloats = kTBins * kTRows * kTCols;

// grid: x = ceil(Wl/32), y = ceil(Hl/4), z = B; 128 threads.
// smem: float lam1[D] | float tile[kTStages][kTBins][kTRows][kTCols]
__global__ void __launch_bounds__(128, 4)
head_fwd_x3t_kernel(const float* __restrict__ cost, float* __restrict__ disp, float* __restrict__ stats,
                    int Dl, int Hl, int Wl, float scale) {
    extern __shared__ __align__(16) float x3t_smem[];
    const int D = 3 * Dl, W = 3 * Wl;
    float* tile = x3t_smem;                            // 16-byte aligned stages first
    float* lam1 = x3t_smem + kTStages * kTStageFloats; // [D]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int C0 = blockIdx.x * 32, R0 = blockIdx.y * 4;   // first block col / row of the CTA
    const int T0 = C0 - 4;                                  // first low-res col held in the tile (16B aligned)
    const size_t plane = (size_t)Hl * Wl;
    const float* base = cost + (size_t)b * Dl * plane;
    const int Wv = Wl >> 2;
    const int n_chunks = (Dl + kTBins - 1) / kTBins;

    // cooperative async copy of bins [ch*8, ch*8+8) x rows R0-1..R0+4 x cols T0..T0+39 (clamped into the image)
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
    for (int k = tid; k < D; k += 128) {
        int t0, t1;
        float l0, l1;
        src_index<true>(scale, k, Dl, t0, t1, l0, l1);
        lam1[k] = l1;
    }

    // ---- per-thread geometry: block (r,c); inactive threads (ragged right/bottom edge) shadow a valid block ----
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
    // tile offsets of the 3x3 low-res neighbourhood {r-1,r,r+1} x {c-1,c,c+1}, clamped like the reference's i0/i1
    int ro[3], co[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ro[i] = (min(max(r - 1 + i, 0), Hl - 1) - (R0 - 1)) * kTCols;
        co[i] = min(max(c - 1 + i, 0), Wl - 1) - T0;
    }
    // pixel (ph,pw): rows (ro[0],ro[1]) for ph=0, (ro[1],ro[2]) for ph=1,2; cols likewise

    float a[9], m[9], dg[9], ng[9], t[9];
    float dhi[9], dlo[9], nhi[9], nlo[9];
    const float kc = 0.5f * (float)D;
    float kf = 1.f - kc;

    // blend of one low-res bin held at `s` (tile stage base + bin offset): t = z - m (m = 0 -> plain exponent)
    auto blend = [&](const float* s, const float (&mm)[9], float (&o)[9]) {
        float x[3][3];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
            const float v0 = s[ro[rr] + co[0]], v1 = s[ro[rr] + co[1]], v2 = s[ro[rr] + co[2]];
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
        __pipeline_wait_prior(2);          // chunk ch has landed (for this thread's copies)
        __syncthreads();                   // ... and for everybody's
        const float* st = tile + (ch % kTStages) * kTStageFloats;
        const int jbeg = ch * kTBins, jend = min(jbeg + kTBins, Dl);
        if (ch == 0) {
            // low-res bin 0 defines the reference exponent; full-res bin 0 sits exactly on it
#pragma unroll
            for (int i = 0; i < 9; ++i) m[i] = 0.f;
            blend(st, m, t);
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                m[i] = t[i]; a[i] = 0.f; dg[i] = 1.f; ng[i] = -kc;
                dhi[i] = dlo[i] = nhi[i] = nlo[i] = 0.f;
            }
        }
        // k-block j (bins 3j+1..3j+3) interpolates low-res bins j and j+1; bin j+1 is `jn`.
        // Within this chunk we can advance while bin j+1 is in the chunk; the block whose upper bin is the
        // first bin of the NEXT chunk is done at the top of the next chunk iteration (jn = jbeg).
        for (int jn = max(jbeg, 1); jn < jend; ++jn) {
            blend(st + (jn - jbeg) * (kTRows * kTCols), m, t);   // exponent of low-res bin jn relative to m
            const int j = jn - 1;
            const float l1 = lam1[3 * j + 1], l2 = lam1[3 * j + 2], l3 = lam1[3 * j + 3];
            float mx = t[0];
#pragma unroll
            for (int i = 1; i < 9; ++i) mx = fmaxf(mx, t[i]);
            if (mx > kX3Tau) {             // rare: some pixel's running maximum ran away from its reference
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
                const float e1 = ex2_approx(__fmaf_rn(l1, dlt, a[i]));
                const float e2 = ex2_approx(__fmaf_rn(l2, dlt, a[i]));
                const float e3 = ex2_approx(__fmaf_rn(l3, dlt, a[i]));
                dg[i] += e1; ng[i] = __fmaf_rn(e1, kf, ng[i]);
                dg[i] += e2; ng[i] = __fmaf_rn(e2, kf + 1.f, ng[i]);
                dg[i] += e3; ng[i] = __fmaf_rn(e3, kf + 2.f, ng[i]);
                a[i] = t[i];
            }
            kf += 3.f;
        }
        // fold the short fp32 group sums (<= 24 bins) into the compensated totals once per chunk
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            two_sum_f(dhi[i], dlo[i], dg[i]); two_sum_f(nhi[i], nlo[i], ng[i]);
            dg[i] = 0.f; ng[i] = 0.f;
        }
        __syncthreads();                   // everyone is done with this stage before it is refilled
    }
    if (!active) return;
    const size_t img = (size_t)3 * Hl * W;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        // last k-block (j = Dl-1): bins 3Dl-2 and 3Dl-1 both sit on low-res bin Dl-1
        const float e = ex2_approx(a[i]);
        float g = e + e, n = __fmaf_rn(e, kf + 1.f, e * kf);
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
