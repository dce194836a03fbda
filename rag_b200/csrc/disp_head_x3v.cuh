// x3 disparity-head forward: lean tiled kernel (disp_head_x3u.cuh) with the per-bin arithmetic on packed
// FP32 (FFMA2/FADD2, sm_100a).  All forward variants so far turned out ISSUE bound (profiles/:
// 190-240 M warp instructions at 63-77% issue utilisation, the SFU at ~60%), so this version minimises
// issued instructions: nine pixels in four register pairs + one scalar
//     P0=(p00,p01) P1=(p10,p11) P2=(p20,p21) P3=(p02,p12) S=p22        (p<row><col>)
// with every broadcast operand loop-invariant (weights, lambdas, bin indices), so pairs never have to
// be assembled by MOVs inside the loop; the k-block loop is unrolled by two with ping-pong exponent
// registers; cp.async slots are precomputed; compensated totals are folded in packed form.
#pragma once
#include "disp_head_x3p.cuh"
#include "disp_head_x3u.cuh"

namespace rag {

struct X3vAcc {          // per-thread softmax state for the 4 pairs + scalar
    float2 dg[4], ng[4];
    float dgS, ngS;
};

// one k-block: bins (l.x,l.y,l.z) between relative exponents a (lower bin) and t (upper bin)
__device__ __forceinline__ void x3v_bins(const float2 (&a)[4], float aS, const float2 (&t)[4], float tS,
                                         float2 l1, float2 l2, float2 l3, float2 kf1, float2 kf2, float2 kf3, X3vAcc& s) {
    const float2 neg1 = f2b(-1.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 dlt = fma2(a[i], neg1, t[i]);
        const float2 e1 = ex2_2(fma2(l1, dlt, a[i]));
        const float2 e2 = ex2_2(fma2(l2, dlt, a[i]));
        const float2 e3 = ex2_2(fma2(l3, dlt, a[i]));
        s.dg[i] = add2(s.dg[i], e1); s.ng[i] = fma2(e1, kf1, s.ng[i]);
        s.dg[i] = add2(s.dg[i], e2); s.ng[i] = fma2(e2, kf2, s.ng[i]);
        s.dg[i] = add2(s.dg[i], e3); s.ng[i] = fma2(e3, kf3, s.ng[i]);
    }
    const float dlt = tS - aS;
    const float e1 = ex2_approx(__fmaf_rn(l1.x, dlt, aS));
    const float e2 = ex2_approx(__fmaf_rn(l2.x, dlt, aS));
    const float e3 = ex2_approx(__fmaf_rn(l3.x, dlt, aS));
    s.dgS += e1; s.ngS = __fmaf_rn(e1, kf1.x, s.ngS);
    s.dgS += e2; s.ngS = __fmaf_rn(e2, kf2.x, s.ngS);
    s.dgS += e3; s.ngS = __fmaf_rn(e3, kf3.x, s.ngS);
}

// grid: x = ceil(Wl/32), y = ceil(Hl/4), z = B; 128 threads.
// BINS low-res bins per pipeline stage, STAGES stages (8x3 or 16x2: the latter halves the per-chunk
// barriers and total folds and is the default).
// smem: tile[STAGES][BINS][kTRows][kTCols] | float2 tot[18][128] | float2 lam[Dl][4] ((l1,l1),(l2,l2),(l3,l3),-)
template <int MINB, int BINS, int STAGES>
__global__ void __launch_bounds__(128, MINB)
head_fwd_x3v_kernel(const float* __restrict__ cost, float* __restrict__ disp, float* __restrict__ stats,
                    int Dl, int Hl, int Wl, float scale) {
    extern __shared__ __align__(16) float x3v_smem[];
    constexpr int kStageFloats = BINS * kTRows * kTCols;
    constexpr int kSlots = (BINS * kTRows * (kTCols / 4) + 127) / 128;
    const int D = 3 * Dl, W = 3 * Wl;
    float* tile = x3v_smem;
    float2* tot = reinterpret_cast<float2*>(x3v_smem + STAGES * kStageFloats);   // [18][128]
    float2* lam = tot + 18 * 128;                                                    // [Dl][4]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int C0 = blockIdx.x * 32, R0 = blockIdx.y * 4;
    const int T0 = C0 - 4;
    const size_t plane = (size_t)Hl * Wl;
    const float* base = cost + (size_t)b * Dl * plane;
    const int Wv = Wl >> 2;
    const int n_chunks = (Dl + BINS - 1) / BINS;
    const bool patch_l = C0 == 0;
    const bool patch_r = (Wl - T0) <= 36;
    const int pr = Wl - T0;

    // cp.async slots: a chunk is BINS bins x 6 rows x 10 vectors; thread tid owns vectors tid, tid+128, ...
    // Source offsets inside a bin plane are fixed.
    int s_off[kSlots], g_off[kSlots], s_bin[kSlots];
#pragma unroll
    for (int q = 0; q < kSlots; ++q) {
        const int u = tid + q * 128;
        const int vec = u % 10, row = (u / 10) % kTRows, bin = u / 60;
        s_bin[q] = bin;
        s_off[q] = (bin * kTRows + row) * kTCols + vec * 4;
        const int gr = min(max(R0 - 1 + row, 0), Hl - 1);
        const int gv = min(max((T0 >> 2) + vec, 0), Wv - 1);
        g_off[q] = gr * Wl + gv * 4;
    }
    auto issue_chunk = [&](int ch) {
        if (ch < n_chunks) {
            float* dst = tile + (ch % STAGES) * kStageFloats;
#pragma unroll
            for (int q = 0; q < kSlots; ++q) {
                if (tid + q * 128 < BINS * kTRows * (kTCols / 4)) {
                    const int gj = min(ch * BINS + s_bin[q], Dl - 1);
                    __pipeline_memcpy_async(dst + s_off[q], base + (size_t)gj * plane + g_off[q], 16);
                }
            }
        }
        __pipeline_commit();
    };
    issue_chunk(0);
    if (STAGES > 2) issue_chunk(1);
    for (int j = tid; j < Dl; j += 128) {
        float l[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            int t0, t1;
            float l0;
            src_index<true>(scale, min(3 * j + 1 + q, D - 1), Dl, t0, t1, l0, l[q]);
        }
        lam[4 * j + 0] = f2b(l[0]); lam[4 * j + 1] = f2b(l[1]); lam[4 * j + 2] = f2b(l[2]); lam[4 * j + 3] = f2b(0.f);
    }
    float2* mytot = tot + tid;
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
    const int org = (r - R0) * kTCols + (c - C0 + 3);
    const float2 wA = f2(wl0[0], wl0[1]), wB = f2(wl1[0], wl1[1]);
    const float2 h0b[3] = {f2b(hs0[0]), f2b(hs0[1]), f2b(hs0[2])};
    const float2 h1b[3] = {f2b(hs1[0]), f2b(hs1[1]), f2b(hs1[2])};
    const float2 h0p = f2(hs0[0], hs0[1]), h1p = f2(hs1[0], hs1[1]);

    float2 mneg[4], ea[4], eb[4];   // ea/eb: ping-pong relative exponents (lower / upper bin of the k-block)
    float mnegS = 0.f, eaS = 0.f, ebS = 0.f;
    X3vAcc acc;
    const float kc = 0.5f * (float)D;
    float2 kf1 = f2b(1.f - kc), kf2 = f2b(2.f - kc), kf3 = f2b(3.f - kc);
    const float2 three = f2b(3.f);

    auto blend = [&](const float* p, const float2 (&mn)[4], float mnS, float2 (&o)[4], float& oS) {
        float2 x01[3];
        float x2[3];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
            const float v0 = p[rr * kTCols + 0], v1 = p[rr * kTCols + 1], v2 = p[rr * kTCols + 2];
            x01[rr] = fma2(wA, f2(v0, v1), mul2(wB, f2(v1, v2)));
            x2[rr] = __fmaf_rn(wl0[2], v1, wl1[2] * v2);
        }
        o[0] = fma2(h0b[0], x01[0], fma2(h1b[0], x01[1], mn[0]));
        o[1] = fma2(h0b[1], x01[1], fma2(h1b[1], x01[2], mn[1]));
        o[2] = fma2(h0b[2], x01[1], fma2(h1b[2], x01[2], mn[2]));
        o[3] = fma2(h0p, f2(x2[0], x2[1]), fma2(h1p, f2(x2[1], x2[2]), mn[3]));
        oS = __fmaf_rn(hs0[2], x2[1], __fmaf_rn(hs1[2], x2[2], mnS));
    };
    auto fold = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 hi = mytot[(4 * i + 0) * 128], lo = mytot[(4 * i + 1) * 128];
            two_sum_acc(hi, lo, acc.dg[i]);
            mytot[(4 * i + 0) * 128] = hi; mytot[(4 * i + 1) * 128] = lo;
            hi = mytot[(4 * i + 2) * 128]; lo = mytot[(4 * i + 3) * 128];
            two_sum_acc(hi, lo, acc.ng[i]);
            mytot[(4 * i + 2) * 128] = hi; mytot[(4 * i + 3) * 128] = lo;
            acc.dg[i] = f2b(0.f); acc.ng[i] = f2b(0.f);
        }
        float2 d = mytot[16 * 128], n = mytot[17 * 128];
        two_sum_acc(d.x, d.y, acc.dgS);
        two_sum_acc(n.x, n.y, acc.ngS);
        mytot[16 * 128] = d; mytot[17 * 128] = n;
        acc.dgS = 0.f; acc.ngS = 0.f;
    };
    // rare path: upper-bin exponents in (t,tS) exceed the reference by > kX3Tau for some pixel
    auto rescale = [&](float2 (&t)[4], float& tS, float2 (&a)[4], float& aS) {
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
    };
    auto maxof = [](const float2 (&t)[4], float tS) {
        return fmaxf(fmaxf(fmaxf(fmaxf(t[0].x, t[0].y), fmaxf(t[1].x, t[1].y)),
                           fmaxf(fmaxf(t[2].x, t[2].y), fmaxf(t[3].x, t[3].y))), tS);
    };

    for (int ch = 0; ch < n_chunks; ++ch) {
        issue_chunk(ch + STAGES - 1);
        __pipeline_wait_prior(STAGES - 1);
        __syncthreads();
        float* stw = tile + (ch % STAGES) * kStageFloats;
        if (patch_l || patch_r) {
            if (tid < BINS * kTRows) {
                float* rowp = stw + tid * kTCols;
                if (patch_l) rowp[3] = rowp[4];
                if (patch_r) rowp[pr] = rowp[pr - 1];
            }
            __syncthreads();
        }
        const int jbeg = ch * BINS, jend = min(jbeg + BINS, Dl);
        const float* p = stw + org;
        if (ch == 0) {
            const float2 z4[4] = {f2b(0.f), f2b(0.f), f2b(0.f), f2b(0.f)};
            blend(p, z4, 0.f, eb, ebS);          // plain exponents of low-res bin 0 = reference exponents
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                mneg[i] = f2(-eb[i].x, -eb[i].y);
                ea[i] = f2b(0.f); acc.dg[i] = f2b(1.f); acc.ng[i] = f2b(-kc);   // full-res bin 0: 2^0 * (0 - kc)
            }
            mnegS = -ebS; eaS = 0.f; acc.dgS = 1.f; acc.ngS = -kc;
            p += kTRows * kTCols;
        }
        int n_it = jend - max(jbeg, 1);
        const float2* lp = lam + 4 * (max(jbeg, 1) - 1);
        // two k-blocks per trip, exponent registers ping-pong (ea -> eb -> ea) so nothing is copied
        for (; n_it >= 2; n_it -= 2) {
            blend(p, mneg, mnegS, eb, ebS);
            if (maxof(eb, ebS) > kX3Tau) rescale(eb, ebS, ea, eaS);
            x3v_bins(ea, eaS, eb, ebS, lp[0], lp[1], lp[2], kf1, kf2, kf3, acc);
            kf1 = add2(kf1, three); kf2 = add2(kf2, three); kf3 = add2(kf3, three);
            blend(p + kTRows * kTCols, mneg, mnegS, ea, eaS);
            if (maxof(ea, eaS) > kX3Tau) rescale(ea, eaS, eb, ebS);
            x3v_bins(eb, ebS, ea, eaS, lp[4], lp[5], lp[6], kf1, kf2, kf3, acc);
            kf1 = add2(kf1, three); kf2 = add2(kf2, three); kf3 = add2(kf3, three);
            p += 2 * kTRows * kTCols;
            lp += 8;
        }
        if (n_it == 1) {
            blend(p, mneg, mnegS, eb, ebS);
            if (maxof(eb, ebS) > kX3Tau) rescale(eb, ebS, ea, eaS);
            x3v_bins(ea, eaS, eb, ebS, lp[0], lp[1], lp[2], kf1, kf2, kf3, acc);
            kf1 = add2(kf1, three); kf2 = add2(kf2, three); kf3 = add2(kf3, three);
#pragma unroll
            for (int i = 0; i < 4; ++i) ea[i] = eb[i];
            eaS = ebS;
        }
        fold();
        __syncthreads();
    }
    // last k-block (j = Dl-1): bins 3Dl-2 and 3Dl-1 both sit on low-res bin Dl-1 (exponent ea)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 e = ex2_2(ea[i]);
        acc.dg[i] = add2(e, e);
        acc.ng[i] = fma2(e, kf2, mul2(e, kf1));
    }
    {
        const float e = ex2_approx(eaS);
        acc.dgS = e + e;
        acc.ngS = __fmaf_rn(e, kf2.x, e * kf1.x);
    }
    fold();
    if (!active) return;

    const size_t img = (size_t)3 * Hl * W;
    auto emit = [&](int ph, int pw, float dhi, float dlo, float nhi, float nlo, float mn) {
        const size_t o = (size_t)(3 * r + ph) * W + (3 * c + pw);
        const float inv = 1.f / (dhi + dlo);
        const float q = nhi * inv;
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
