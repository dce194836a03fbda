// x3 disparity-head kernels (maxdisp == 3*Dl, the only ratio the reference produces).
//
// Geometry.  Output pixel rows 3r+1..3r+3 (r in [-1,Hl-1]) all interpolate between low-res rows
// {max(r,0), min(max(r,0)+1,Hl-1)}, likewise for columns, and full-res bins 3j+1..3j+3 interpolate
// between low-res bins j and min(j+1,Dl-1) (bin 0 uses low-res bin 0 alone).  So one thread owns a
// 3x3 pixel block: per low-res bin it loads 4 values, forms the nine bilinear blends with 30 FP ops
// (3.3 per pixel instead of 10) and walks the 3 full-res bins of that k-block for its 9 pixels.
// A warp = 32 consecutive block columns.
//
// All weights come from src_index<true>() -- PyTorch's fp32 lambda arithmetic -- never from 1/3, 2/3.
#pragma once
#include "head_common.cuh"

namespace rag {

// blends are produced directly as log2-domain exponents: z = -log2(e) * v

// Tables of one axis of block `blk`: lambdas of its 3 full-res indices (invalid ones clamped into
// range) and, for the backward, the weights with which each index feeds low-res cell `blk` (wa) and
// cell `blk+1` (wb) -- the general form covers the clamped first/last blocks.
struct X3Axis {
    float l0[3], l1[3], wa[3], wb[3];
    bool valid[3];
    int lo0, lo1, idx[3];
};

__device__ __forceinline__ void x3_axis(float scale, int blk, int n_lo, X3Axis& ax) {
    const int n_hi = 3 * n_lo;
    ax.lo0 = max(blk, 0);
    ax.lo1 = min(ax.lo0 + 1, n_lo - 1);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int raw = 3 * blk + 1 + i;
        ax.valid[i] = raw >= 0 && raw < n_hi;
        const int id = min(max(raw, 0), n_hi - 1);
        ax.idx[i] = id;
        int i0, i1;
        src_index<true>(scale, id, n_lo, i0, i1, ax.l0[i], ax.l1[i]);
        ax.wa[i] = ax.valid[i] ? ((i0 == blk ? ax.l0[i] : 0.f) + (i1 == blk ? ax.l1[i] : 0.f)) : 0.f;
        ax.wb[i] = ax.valid[i] ? ((i0 == blk + 1 ? ax.l0[i] : 0.f) + (i1 == blk + 1 ? ax.l1[i] : 0.f)) : 0.f;
    }
}

struct X3Loader {
    const float* p00; const float* p01; const float* p10; const float* p11;
    size_t plane;
    int j, jmax;
    float v[4];
    // positions the four corner pointers on low-res bin j0 and loads it
    __device__ __forceinline__ void init(const float* base, int Wl, size_t plane_, int Dl, int j0, int rl0, int rl1, int cl0, int cl1) {
        plane = plane_;
        j = j0;
        jmax = Dl - 1;
        base += (size_t)j0 * plane;
        p00 = base + rl0 * Wl + cl0; p01 = base + rl0 * Wl + cl1;
        p10 = base + rl1 * Wl + cl0; p11 = base + rl1 * Wl + cl1;
        fetch();
    }
    __device__ __forceinline__ void fetch() { v[0] = __ldg(p00); v[1] = __ldg(p01); v[2] = __ldg(p10); v[3] = __ldg(p11); }
    // loads bin min(j+1, Dl-1): pointer bumps instead of re-deriving 64-bit addresses every bin
    __device__ __forceinline__ void next() {
        if (j < jmax) { ++j; p00 += plane; p01 += plane; p10 += plane; p11 += plane; }
        fetch();
    }
};

// nine blends z[ph*3+pw] = hs0[ph]*(wl0*v0 + wl1*v1) + hs1[ph]*(wl0*v2 + wl1*v3), hs = -log2e * h-lambda
__device__ __forceinline__ void x3_blend9(const float (&v)[4], const float (&wl0)[3], const float (&wl1)[3],
                                          const float (&hs0)[3], const float (&hs1)[3], float (&z)[9]) {
    float x0[3], x1[3];
#pragma unroll
    for (int pw = 0; pw < 3; ++pw) {
        x0[pw] = __fmaf_rn(wl0[pw], v[0], wl1[pw] * v[1]);
        x1[pw] = __fmaf_rn(wl0[pw], v[2], wl1[pw] * v[3]);
    }
#pragma unroll
    for (int ph = 0; ph < 3; ++ph)
#pragma unroll
        for (int pw = 0; pw < 3; ++pw) z[ph * 3 + pw] = __fmaf_rn(hs0[ph], x0[pw], hs1[ph] * x1[pw]);
}

// same blends, minus the pixel's reference exponent, in 2 FFMA per pixel: t = hs0*x0 + (hs1*x1 - m)
__device__ __forceinline__ void x3_blend9_rel(const float (&v)[4], const float (&wl0)[3], const float (&wl1)[3],
                                              const float (&hs0)[3], const float (&hs1)[3], const float (&m)[9], float (&t)[9]) {
    float x0[3], x1[3];
#pragma unroll
    for (int pw = 0; pw < 3; ++pw) {
        x0[pw] = __fmaf_rn(wl0[pw], v[0], wl1[pw] * v[1]);
        x1[pw] = __fmaf_rn(wl0[pw], v[2], wl1[pw] * v[3]);
    }
#pragma unroll
    for (int ph = 0; ph < 3; ++ph)
#pragma unroll
        for (int pw = 0; pw < 3; ++pw)
            t[ph * 3 + pw] = __fmaf_rn(hs0[ph], x0[pw], __fmaf_rn(hs1[ph], x1[pw], -m[ph * 3 + pw]));
}

// ---------------------------------------------------------------------------------------------
// forward.  grid: x = ceil((Wl+1)/32), y = ceil((Hl+1)/WARPS), z = B.
// Per pixel: exponents relative to a reference m that moves only when the running maximum exceeds
// it by more than kX3Tau (rare, handled in one out-of-line branch per k-block); sums in short fp32
// group accumulators folded into fp64 totals every 8 k-blocks; centred regression.
// ---------------------------------------------------------------------------------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 512 / (WARPS * 32))
head_fwd_x3_kernel(const float* __restrict__ cost, float* __restrict__ disp, float* __restrict__ stats,
                   int Dl, int Hl, int Wl, float scale) {
    extern __shared__ float x3_smem[];
    float* lam1 = x3_smem;  // [D] lambda1 of full-res bin k
    const int D = 3 * Dl, H = 3 * Hl, W = 3 * Wl;
    for (int k = threadIdx.x; k < D; k += WARPS * 32) {
        int t0, t1;
        float l0, l1;
        src_index<true>(scale, k, Dl, t0, t1, l0, l1);
        lam1[k] = l1;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane - 1;
    const int r = blockIdx.y * WARPS + warp - 1;
    const int b = blockIdx.z;
    if (c > Wl - 1 || r > Hl - 1) return;

    X3Axis ah, aw;
    x3_axis(scale, r, Hl, ah);
    x3_axis(scale, c, Wl, aw);
    float hs0[3], hs1[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { hs0[i] = ah.l0[i] * kX3NegLog2e; hs1[i] = ah.l1[i] * kX3NegLog2e; }

    const size_t plane = (size_t)Hl * Wl;
    X3Loader ld;
    ld.init(cost + (size_t)b * Dl * plane, Wl, plane, Dl, 0, ah.lo0, ah.lo1, aw.lo0, aw.lo1);

    float a[9], m[9], dg[9], ng[9], nxt[9];
    double dend[9], numd[9];
    const float kc = 0.5f * (float)D;

    x3_blend9(ld.v, aw.l0, aw.l1, hs0, hs1, nxt);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        m[i] = nxt[i];   // reference exponent = exponent of bin 0
        a[i] = 0.f;      // current low-res exponent relative to m
        dg[i] = 1.f;     // full-res bin 0 (lambda1 == 0): 2^0
        ng[i] = -kc;     //   ... times (0 - kc)
        dend[i] = 0.0; numd[i] = 0.0;
    }
    ld.next();
    float kf = 1.f - kc;  // centred index of the first bin of the k-block

#pragma unroll 2
    for (int j = 0; j < Dl - 1; ++j) {
        x3_blend9_rel(ld.v, aw.l0, aw.l1, hs0, hs1, m, nxt);  // exponent of low-res bin j+1 relative to m
        ld.next();                                             // prefetch bin min(j+2, Dl-1)
        const float l1 = lam1[3 * j + 1], l2 = lam1[3 * j + 2], l3 = lam1[3 * j + 3];
        float mx = nxt[0];
#pragma unroll
        for (int i = 1; i < 9; ++i) mx = fmaxf(mx, nxt[i]);
        if (mx > kX3Tau) {                                     // rare: move the reference(s) up
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                if (nxt[i] > kX3Tau) {
                    const float f = ex2_approx(-nxt[i]);
                    dg[i] *= f; ng[i] *= f;
                    dend[i] *= (double)f; numd[i] *= (double)f;
                    m[i] += nxt[i];
                    a[i] -= nxt[i];
                    nxt[i] = 0.f;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const float dlt = nxt[i] - a[i];
            const float e1 = ex2_approx(__fmaf_rn(l1, dlt, a[i]));
            const float e2 = ex2_approx(__fmaf_rn(l2, dlt, a[i]));
            const float e3 = ex2_approx(__fmaf_rn(l3, dlt, a[i]));
            dg[i] += e1; ng[i] = __fmaf_rn(e1, kf, ng[i]);
            dg[i] += e2; ng[i] = __fmaf_rn(e2, kf + 1.f, ng[i]);
            dg[i] += e3; ng[i] = __fmaf_rn(e3, kf + 2.f, ng[i]);
            a[i] = nxt[i];
        }
        kf += 3.f;
        if ((j & 7) == 7) {  // fold the short fp32 group sums into the fp64 totals
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                dend[i] += (double)dg[i]; numd[i] += (double)ng[i];
                dg[i] = 0.f; ng[i] = 0.f;
            }
        }
    }
    // last k-block (j = Dl-1): bins 3Dl-2 and 3Dl-1 both sit on low-res bin Dl-1
    const size_t img = (size_t)H * W;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const float e = ex2_approx(a[i]);
        dg[i] += e + e;
        ng[i] = __fmaf_rn(e, kf, ng[i]);
        ng[i] = __fmaf_rn(e, kf + 1.f, ng[i]);
        dend[i] += (double)dg[i]; numd[i] += (double)ng[i];
    }
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
        if (!ah.valid[ph]) continue;
#pragma unroll
        for (int pw = 0; pw < 3; ++pw) {
            if (!aw.valid[pw]) continue;
            const int i = ph * 3 + pw;
            const size_t o = (size_t)ah.idx[ph] * W + aw.idx[pw];
            const float inv = 1.f / (float)dend[i];
            disp[(size_t)b * img + o] = kc + (float)numd[i] * inv;
            if (stats) {
                stats[(size_t)b * 2 * img + o] = m[i];
                stats[(size_t)b * 2 * img + img + o] = inv;
            }
        }
    }
}

}  // namespace rag
