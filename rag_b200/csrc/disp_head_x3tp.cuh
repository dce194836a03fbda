// x3 disparity-head forward: tiled (disp_head_x3t.cuh) + packed FP32 (disp_head_x3p.cuh).
//
// Measured on B200 (tools/mufu_mix2.cu, profiles/): with MUFU.EX2 in the mix an SM sub-partition
// sustains ~0.75 issued instructions per clock, and one MUFU.EX2 per warp occupies the SFU for 8 clk.
// The scalar kernels need ~9.5 instructions per exp2, i.e. ~13-14 clk per exp2 -- issue bound, not
// SFU bound.  This version keeps the nine pixels of a block in four register pairs + one scalar
//     P0=(p00,p01) P1=(p10,p11) P2=(p20,p21) P3=(p02,p12) S=p22        (p<row><col>)
// and runs the blend's row stage, the exponent interpolation and both accumulations on FFMA2/FADD2
// (two fp32 lanes per issue slot): ~5 instructions per exp2, which puts the SFU back in charge.
// Shared-memory staging (3-stage cp.async pipeline over 8-bin chunks) and centred blocks as in x3t;
// compensated (hi,lo) totals live in shared memory and are touched once per chunk.
#pragma once
#include "disp_head_x3p.cuh"
#include "disp_head_x3t.cuh"

namespace rag {

// grid: x = ceil(Wl/32), y = ceil(Hl/4), z = B; 128 threads.
// smem: tile[kTStages][kTBins][kTRows][kTCols] | float2 tot[18][128] | float2 lam[D]
__global__ void __launch_bounds__(128, 4)
head_fwd_x3tp_kernel(const float* __restrict__ cost, float* __restrict__ disp, float* __restrict__ stats,
                     int Dl, int Hl, int Wl, float scale) {
    extern __shared__ __align__(16) float x3tp_smem[];
    const int D = 3 * Dl, W = 3 * Wl;
    float* tile = x3tp_smem;
    float2* tot = reinterpret_cast<float2*>(x3tp_smem + kTStages * kTStageFloats);  // [18][128]
    float2* lam = tot + 18 * 128;                                                    // [D] (lambda1, lambda1)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int C0 = blockIdx.x * 32, R0 = blockIdx.y * 4;
    const int T0 = C0 - 4;
    const size_t plane = (size_t)Hl * Wl;
    const float* base = cost + (size_t)b * Dl * plane;
    const int Wv = Wl >> 2;
    const int n_chunks = (Dl + kTBins - 1) / kTBins;

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
        lam[k] = f2b(l1);
    }
    float2* mytot = tot + tid;   // slot s at mytot[s * 128]
#pragma unroll
    for (int s = 0; s < 18; ++s) mytot[s * 128] = f2b(0.f);

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
    int ro[3], co[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ro[i] = (min(max(r - 1 + i, 0), Hl - 1) - (R0 - 1)) * kTCols;
        co[i] = min(max(c - 1 + i, 0), Wl - 1) - T0;
    }
    // packed weights.  Columns 0,1 of a row form a pair; column 2 stays scalar and is paired across rows.
    const float2 wA = f2(wl0[0], wl0[1]), wB = f2(wl1[0], wl1[1]);
    const float2 h0b[3] = {f2b(hs0[0]), f2b(hs0[1]), f2b(hs0[2])};
    const float2 h1b[3] = {f2b(hs1[0]), f2b(hs1[1]), f2b(hs1[2])};
    const float2 h0p = f2(hs0[0], hs0[1]), h1p = f2(hs1[0], hs1[1]);

    float2 a[4], mneg[4], dg[4], ng[4], t[4];
    float aS, mnegS, dgS, ngS, tS;
    const float kc = 0.5f * (float)D;
    float2 kf1 = f2b(1.f - kc), kf2 = f2b(2.f - kc), kf3 = f2b(3.f - kc);
    const float2 three = f2b(3.f), neg1 = f2b(-1.f);

    // t = blend(bin at s) + mn   (mn = -reference exponent; 0 for the very first bin)
    auto blend = [&](const float* s, const float2 (&mn)[4], float mnS, float2 (&o)[4], float& oS) {
        float2 x01[3];
        float x2[3];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
            const float v0 = s[ro[rr] + co[0]], v1 = s[ro[rr] + co[1]], v2 = s[ro[rr] + co[2]];
            x01[rr] = fma2(wA, f2(v0, v1), mul2(wB, f2(v1, v2)));     // cols 0,1: (c-1,c) and (c,c+1)
            x2[rr] = __fmaf_rn(wl0[2], v1, wl1[2] * v2);             // col 2: (c,c+1)
        }
        o[0] = fma2(h0b[0], x01[0], fma2(h1b[0], x01[1], mn[0]));    // row 0: rows (r-1,r)
        o[1] = fma2(h0b[1], x01[1], fma2(h1b[1], x01[2], mn[1]));    // row 1: rows (r,r+1)
        o[2] = fma2(h0b[2], x01[1], fma2(h1b[2], x01[2], mn[2]));    // row 2: rows (r,r+1)
        o[3] = fma2(h0p, f2(x2[0], x2[1]), fma2(h1p, f2(x2[1], x2[2]), mn[3]));   // (p02, p12)
        oS = __fmaf_rn(hs0[2], x2[1], __fmaf_rn(hs1[2], x2[2], mnS));             // p22
    };
    auto fold = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 hi = mytot[(4 * i + 0) * 128], lo = mytot[(4 * i + 1) * 128];
            two_sum_acc(hi, lo, dg[i]);
            mytot[(4 * i + 0) * 128] = hi; mytot[(4 * i + 1) * 128] = lo;
            hi = mytot[(4 * i + 2) * 128]; lo = mytot[(4 * i + 3) * 128];
            two_sum_acc(hi, lo, ng[i]);
            mytot[(4 * i + 2) * 128] = hi; mytot[(4 * i + 3) * 128] = lo;
            dg[i] = f2b(0.f); ng[i] = f2b(0.f);
        }
        float2 d = mytot[16 * 128], n = mytot[17 * 128];   // scalar pixel: (hi, lo)
        two_sum_acc(d.x, d.y, dgS);
        two_sum_acc(n.x, n.y, ngS);
        mytot[16 * 128] = d; mytot[17 * 128] = n;
        dgS = 0.f; ngS = 0.f;
    };

    for (int ch = 0; ch < n_chunks; ++ch) {
        issue_chunk(ch + 2);
        __pipeline_wait_prior(2);
        __syncthreads();
        const float* st = tile + (ch % kTStages) * kTStageFloats;
        const int jbeg = ch * kTBins, jend = min(jbeg + kTBins, Dl);
        if (ch == 0) {
            const float2 z4[4] = {f2b(0.f), f2b(0.f), f2b(0.f), f2b(0.f)};
            blend(st, z4, 0.f, t, tS);       // plain exponents of low-res bin 0 = reference exponents
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                mneg[i] = f2(-t[i].x, -t[i].y);
                a[i] = f2b(0.f); dg[i] = f2b(1.f); ng[i] = f2b(-kc);   // full-res bin 0: 2^0 * (0 - kc)
            }
            mnegS = -tS; aS = 0.f; dgS = 1.f; ngS = -kc;
        }
        for (int jn = max(jbeg, 1); jn < jend; ++jn) {
            blend(st + (jn - jbeg) * (kTRows * kTCols), mneg, mnegS, t, tS);
            const int j = jn - 1;
            const float2 l1 = lam[3 * j + 1], l2 = lam[3 * j + 2], l3 = lam[3 * j + 3];
            const float mx = fmaxf(fmaxf(fmaxf(fmaxf(t[0].x, t[0].y), fmaxf(t[1].x, t[1].y)),
                                         fmaxf(fmaxf(t[2].x, t[2].y), fmaxf(t[3].x, t[3].y))), tS);
            if (mx > kX3Tau) {               // rare: move the runaway pixels' reference exponents up
                fold();
                auto fix = [&](float& tt, float& aa, float& mn, int slot, bool hi_lane) {
                    if (tt > kX3Tau) {
                        const float f = ex2_approx(-tt);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float2 v = mytot[(slot + q) * 128];
                            if (hi_lane) v.y *= f; else v.x *= f;
                            mytot[(slot + q) * 128] = v;
                        }
                        mn -= tt; aa -= tt; tt = 0.f;
                    }
                };
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    fix(t[i].x, a[i].x, mneg[i].x, 4 * i, false);
                    fix(t[i].y, a[i].y, mneg[i].y, 4 * i, true);
                }
                if (tS > kX3Tau) {
                    const float f = ex2_approx(-tS);
                    float2 d = mytot[16 * 128], n = mytot[17 * 128];
                    d.x *= f; d.y *= f; n.x *= f; n.y *= f;
                    mytot[16 * 128] = d; mytot[17 * 128] = n;
                    mnegS -= tS; aS -= tS; tS = 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 dlt = fma2(a[i], neg1, t[i]);
                const float2 e1 = ex2_2(fma2(l1, dlt, a[i]));
                const float2 e2 = ex2_2(fma2(l2, dlt, a[i]));
                const float2 e3 = ex2_2(fma2(l3, dlt, a[i]));
                dg[i] = add2(dg[i], e1); ng[i] = fma2(e1, kf1, ng[i]);
                dg[i] = add2(dg[i], e2); ng[i] = fma2(e2, kf2, ng[i]);
                dg[i] = add2(dg[i], e3); ng[i] = fma2(e3, kf3, ng[i]);
                a[i] = t[i];
            }
            {
                const float dlt = tS - aS;
                const float e1 = ex2_approx(__fmaf_rn(l1.x, dlt, aS));
                const float e2 = ex2_approx(__fmaf_rn(l2.x, dlt, aS));
                const float e3 = ex2_approx(__fmaf_rn(l3.x, dlt, aS));
                dgS += e1; ngS = __fmaf_rn(e1, kf1.x, ngS);
                dgS += e2; ngS = __fmaf_rn(e2, kf2.x, ngS);
                dgS += e3; ngS = __fmaf_rn(e3, kf3.x, ngS);
                aS = tS;
            }
            kf1 = add2(kf1, three); kf2 = add2(kf2, three); kf3 = add2(kf3, three);
        }
        fold();                              // once per 8-bin chunk
        __syncthreads();
    }
    // last k-block (j = Dl-1): bins 3Dl-2 and 3Dl-1 both sit on low-res bin Dl-1
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 e = ex2_2(a[i]);
        dg[i] = add2(e, e);
        ng[i] = fma2(e, kf2, mul2(e, kf1));
    }
    {
        const float e = ex2_approx(aS);
        dgS = e + e;
        ngS = __fmaf_rn(e, kf2.x, e * kf1.x);
    }
    fold();
    if (!active) return;

    const size_t img = (size_t)3 * Hl * W;
    auto emit = [&](int ph, int pw, float dhi, float dlo, float nhi, float nlo, float mn) {
        const size_t o = (size_t)(3 * r + ph) * W + (3 * c + pw);
        const float inv = 1.f / (dhi + dlo);
        const float q = nhi * inv;                                       // num/den, low parts to first order
        const float rr2 = __fmaf_rn(-q, dhi, nhi) + (nlo - q * dlo);
        disp[(size_t)b * img + o] = kc + (q + rr2 * inv);
        if (stats) {
            stats[(size_t)b * 2 * img + o] = -mn;
            stats[(size_t)b * 2 * img + img + o] = inv;
        }
    };
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 dhi = mytot[(4 * i + 0) * 128], dlo = mytot[(4 * i + 1) * 128];
        const float2 nhi = mytot[(4 * i + 2) * 128], nlo = mytot[(4 * i + 3) * 128];
        if (i < 3) {
            emit(i, 0, dhi.x, dlo.x, nhi.x, nlo.x, mneg[i].x);
            emit(i, 1, dhi.y, dlo.y, nhi.y, nlo.y, mneg[i].y);
        } else {
            emit(0, 2, dhi.x, dlo.x, nhi.x, nlo.x, mneg[i].x);
            emit(1, 2, dhi.y, dlo.y, nhi.y, nlo.y, mneg[i].y);
        }
    }
    {
        const float2 d = mytot[16 * 128], n = mytot[17 * 128];
        emit(2, 2, d.x, d.y, n.x, n.y, mnegS);
    }
}

}  // namespace rag
