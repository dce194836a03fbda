// x3 disparity-head forward, packed-FP32 version (sm_100a FFMA2/FADD2/FMUL2: two fp32 lanes per
// issue slot).  Same geometry and numerics plan as disp_head_x3.cuh; what changes is the issue
// budget.  The scalar kernel needs ~8 issue slots per pixel-bin and the SFU needs 8 clk per
// pixel-bin (one MUFU.EX2 per warp = 32 lanes / 4 SFU lanes), so neither pipe can saturate.  Here the
// nine pixels of a block are held as four register pairs + one scalar,
//     P0=(p00,p01) P1=(p10,p11) P2=(p20,p21) P3=(p02,p12) S=p22        (p<row><col>)
// chosen so that every operand of the blend is a natural pair (row weights pair with P3, column
// weights pair with P0..P2); blend, exponent interpolation and both accumulations run on FFMA2 /
// FADD2, leaving MUFU.EX2 as the only scalar instruction of the inner loop: ~4.5 issue slots per
// pixel-bin, so the kernel runs at the SFU floor.
//
// Accumulation: per pixel, fp32 group sums over 8 k-blocks (24 bins), folded with an error-free
// TwoSum into (hi,lo) fp32 totals that live in shared memory (keeps the register file for 16+
// resident warps per SM and keeps F2F/FP64 off the SFU pipe).
#pragma once
#include "disp_head_x3.cuh"

namespace rag {

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2b(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 ex2_2(float2 z) { return make_float2(ex2_approx(z.x), ex2_approx(z.y)); }

// error-free accumulation of g into (hi, lo):  hi + lo + g  ==  hi' + lo'  up to O(eps^2)
__device__ __forceinline__ void two_sum_acc(float2& hi, float2& lo, float2 g) {
    const float2 neg1 = f2b(-1.f);
    const float2 s = add2(hi, g);
    const float2 bb = fma2(hi, neg1, s);                       // s - hi
    const float2 e1 = fma2(fma2(bb, neg1, s), neg1, hi);       // hi - (s - bb)
    const float2 e2 = fma2(bb, neg1, g);                       // g - bb
    lo = add2(lo, add2(e1, e2));
    hi = s;
}
__device__ __forceinline__ void two_sum_acc(float& hi, float& lo, float g) {
    const float s = hi + g;
    const float bb = s - hi;
    lo += (hi - (s - bb)) + (g - bb);
    hi = s;
}

// grid: x = ceil((Wl+1)/32), y = ceil((Hl+1)/WARPS), z = B.  smem: float2 lam[D] + totals.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 512 / (WARPS * 32))
head_fwd_x3p_kernel(const float* __restrict__ cost, float* __restrict__ disp, float* __restrict__ stats,
                    int Dl, int Hl, int Wl, float scale) {
    constexpr int NT = WARPS * 32;
    extern __shared__ float2 x3p_smem[];
    float2* lam = x3p_smem;                 // [D]  (lambda1, lambda1) of full-res bin k
    const int D = 3 * Dl, H = 3 * Hl, W = 3 * Wl;
    float2* tot = x3p_smem + D;             // [20][NT]: den hi/lo, num hi/lo as pairs P0..P3 + S
    for (int k = threadIdx.x; k < D; k += NT) {
        int t0, t1;
        float l0, l1;
        src_index<true>(scale, k, Dl, t0, t1, l0, l1);
        lam[k] = f2b(l1);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane - 1;
    const int r = blockIdx.y * WARPS + warp - 1;
    const int b = blockIdx.z;
    if (c > Wl - 1 || r > Hl - 1) return;
    float2* mytot = tot + threadIdx.x;      // slot s at mytot[s * NT]

    X3Axis ah, aw;
    x3_axis(scale, r, Hl, ah);
    x3_axis(scale, c, Wl, aw);
    // column weights: pair (pw0,pw1) and broadcast pw2; row weights (scaled by -log2e): broadcast per
    // row for P0..P2, pair (ph0,ph1) for P3, scalar ph2 for S
    const float2 w0p = f2(aw.l0[0], aw.l0[1]), w1p = f2(aw.l1[0], aw.l1[1]);
    const float2 w0c = f2b(aw.l0[2]), w1c = f2b(aw.l1[2]);
    float2 h0b[3], h1b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { h0b[i] = f2b(ah.l0[i] * kX3NegLog2e); h1b[i] = f2b(ah.l1[i] * kX3NegLog2e); }
    const float2 h0p = f2(h0b[0].x, h0b[1].x), h1p = f2(h1b[0].x, h1b[1].x);

    const size_t plane = (size_t)Hl * Wl;
    X3Loader ld;
    ld.init(cost + (size_t)b * Dl * plane, Wl, plane, Dl, 0, ah.lo0, ah.lo1, aw.lo0, aw.lo1);

    // t = blend - m for the 4 pairs + scalar
    float2 mneg[4], a[4], dg[4], ng[4], t[4];
    float mnegS, aS, dgS, ngS, tS;
    auto blend = [&](const float2 (&mn)[4], float mnS, float2 (&o)[4], float& oS) {
        const float2 v0 = f2b(ld.v[0]), v1 = f2b(ld.v[1]), v2 = f2b(ld.v[2]), v3 = f2b(ld.v[3]);
        const float2 x0p = fma2(w0p, v0, mul2(w1p, v1)), x1p = fma2(w0p, v2, mul2(w1p, v3));   // cols 0,1
        const float2 x0c = fma2(w0c, v0, mul2(w1c, v1)), x1c = fma2(w0c, v2, mul2(w1c, v3));   // col 2 (broadcast)
#pragma unroll
        for (int ph = 0; ph < 3; ++ph) o[ph] = fma2(h0b[ph], x0p, fma2(h1b[ph], x1p, mn[ph]));
        o[3] = fma2(h0p, x0c, fma2(h1p, x1c, mn[3]));
        oS = __fmaf_rn(h0b[2].x, x0c.x, __fmaf_rn(h1b[2].x, x1c.x, mnS));
    };
    {
        const float2 zero4[4] = {f2b(0.f), f2b(0.f), f2b(0.f), f2b(0.f)};
        blend(zero4, 0.f, t, tS);   // plain exponents of low-res bin 0
    }
    const float kc = 0.5f * (float)D;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        mneg[i] = f2(-t[i].x, -t[i].y);   // reference exponent = exponent of bin 0
        a[i] = f2b(0.f);
        dg[i] = f2b(1.f);                 // full-res bin 0 (lambda1 == 0): 2^0 ...
        ng[i] = f2b(-kc);                 //   ... times (0 - kc)
    }
    mnegS = -tS; aS = 0.f; dgS = 1.f; ngS = -kc;
#pragma unroll
    for (int s = 0; s < 20; ++s) mytot[s * NT] = f2b(0.f);
    ld.next();
    float2 kf1 = f2b(1.f - kc), kf2 = f2b(2.f - kc), kf3 = f2b(3.f - kc);   // centred bin indices of the k-block
    const float2 three = f2b(3.f), neg1 = f2b(-1.f);

    auto fold = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 hi = mytot[(4 * i + 0) * NT], lo = mytot[(4 * i + 1) * NT];
            two_sum_acc(hi, lo, dg[i]);
            mytot[(4 * i + 0) * NT] = hi; mytot[(4 * i + 1) * NT] = lo;
            hi = mytot[(4 * i + 2) * NT]; lo = mytot[(4 * i + 3) * NT];
            two_sum_acc(hi, lo, ng[i]);
            mytot[(4 * i + 2) * NT] = hi; mytot[(4 * i + 3) * NT] = lo;
            dg[i] = f2b(0.f); ng[i] = f2b(0.f);
        }
        float2 d = mytot[16 * NT], n = mytot[17 * NT];   // scalar pixel: (hi, lo) packed in one float2 each
        two_sum_acc(d.x, d.y, dgS);
        two_sum_acc(n.x, n.y, ngS);
        mytot[16 * NT] = d; mytot[17 * NT] = n;
        dgS = 0.f; ngS = 0.f;
    };

#pragma unroll 2
    for (int j = 0; j < Dl - 1; ++j) {
        blend(mneg, mnegS, t, tS);        // exponent of low-res bin j+1 relative to m
        ld.next();                        // prefetch bin min(j+2, Dl-1)
        const float2 l1 = lam[3 * j + 1], l2 = lam[3 * j + 2], l3 = lam[3 * j + 3];
        const float mx = fmaxf(fmaxf(fmaxf(fmaxf(t[0].x, t[0].y), fmaxf(t[1].x, t[1].y)),
                                     fmaxf(fmaxf(t[2].x, t[2].y), fmaxf(t[3].x, t[3].y))), tS);
        if (mx > kX3Tau) {                // rare: some pixel's running maximum ran away from its reference
            fold();
            auto fix = [&](float& tt, float& aa, float& mn, int slot, bool hi_lane) {
                if (tt > kX3Tau) {
                    const float f = ex2_approx(-tt);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float2 v = mytot[(slot + q) * NT];
                        if (hi_lane) v.y *= f; else v.x *= f;
                        mytot[(slot + q) * NT] = v;
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
                float2 d = mytot[16 * NT], n = mytot[17 * NT];
                d.x *= f; d.y *= f; n.x *= f; n.y *= f;
                mytot[16 * NT] = d; mytot[17 * NT] = n;
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
        if ((j & 7) == 7) fold();
    }
    // last k-block (j = Dl-1): bins 3Dl-2 and 3Dl-1 both sit on low-res bin Dl-1
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 e = ex2_2(a[i]);
        dg[i] = add2(dg[i], add2(e, e));
        ng[i] = fma2(e, kf1, ng[i]);
        ng[i] = fma2(e, kf2, ng[i]);
    }
    {
        const float e = ex2_approx(aS);
        dgS += e + e;
        ngS = __fmaf_rn(e, kf1.x, ngS);
        ngS = __fmaf_rn(e, kf2.x, ngS);
    }
    fold();

    const size_t img = (size_t)H * W;
    auto emit = [&](int ph, int pw, float dhi, float dlo, float nhi, float nlo, float mn) {
        if (!ah.valid[ph] || !aw.valid[pw]) return;
        const size_t o = (size_t)ah.idx[ph] * W + aw.idx[pw];
        const float inv = 1.f / (dhi + dlo);
        // num/den with the low parts folded in to first order
        const float q = nhi * inv;
        const float rr = __fmaf_rn(-q, dhi, nhi) + (nlo - q * dlo);
        disp[(size_t)b * img + o] = kc + (q + rr * inv);
        if (stats) {
            stats[(size_t)b * 2 * img + o] = -mn;
            stats[(size_t)b * 2 * img + img + o] = inv;
        }
    };
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 dhi = mytot[(4 * i + 0) * NT], dlo = mytot[(4 * i + 1) * NT];
        const float2 nhi = mytot[(4 * i + 2) * NT], nlo = mytot[(4 * i + 3) * NT];
        if (i < 3) {
            emit(i, 0, dhi.x, dlo.x, nhi.x, nlo.x, mneg[i].x);
            emit(i, 1, dhi.y, dlo.y, nhi.y, nlo.y, mneg[i].y);
        } else {
            emit(0, 2, dhi.x, dlo.x, nhi.x, nlo.x, mneg[i].x);
            emit(1, 2, dhi.y, dlo.y, nhi.y, nlo.y, mneg[i].y);
        }
    }
    {
        const float2 d = mytot[16 * NT], n = mytot[17 * NT];
        emit(2, 2, d.x, d.y, n.x, n.y, mnegS);
    }
}

}  // namespace rag
