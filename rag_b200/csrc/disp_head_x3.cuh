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
#include "common.cuh"

namespace rag {

// blends are produced directly as log2-domain exponents: z = -log2(e) * v
constexpr float kX3NegLog2e = -1.4426950408889634f;
constexpr float kX3Tau = 24.0f;  // lazy-rescale threshold (log2 units)

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

// ---------------------------------------------------------------------------------------------
// backward: deterministic, atomics-free.
//   gcost[j,r,c] = sum_{k,h,w} Wd(j,k) Wh(r,h) Ww(c,w) * ( -g[h,w] p_k[h,w] (k - disp[h,w]) ),
//   p_k = 2^(z_k - m) * inv  rebuilt from cost_lr and the forward's stats.
// A warp owns 31 low-res columns (32 block columns c0-1..c0+30: lane l -> block c0-1+l, and each
// low-res cell column gets the "A" part of its own block plus the "B" part of the block to its
// left via one shuffle), TR low-res rows (it sweeps block rows r0-1..r0+TR-1 and carries each
// block row's "B" part to the next row in shared memory) and a chunk of J low-res bins (k-blocks
// j0-1..j0+J-1, the "B" part of a k-block is carried to the next in registers).  Every gcost element
// is produced by exactly one lane with a fixed summation order.
// grid: x = ceil(Wl/31), y = ceil(n_tasks/4), z = B; task = row_tile * nJ + bin_chunk.
// ---------------------------------------------------------------------------------------------
template <int J>
__global__ void __launch_bounds__(128)
head_bwd_x3_kernel(const float* __restrict__ cost, const float* __restrict__ gdisp, const float* __restrict__ disp,
                   const float* __restrict__ stats, float* __restrict__ gcost,
                   int Dl, int Hl, int Wl, float scale, int TR, int nJ, int n_tasks) {
    extern __shared__ float x3_smem[];
    const int D = 3 * Dl, H = 3 * Hl, W = 3 * Wl;
    float* lz = x3_smem;            // [D+3] lambda1 of bin k (exponent interpolation)
    float* dA = lz + (D + 3);       // [D+3] weight of bin k into low-res cell (k-1)/3
    float* dB = dA + (D + 3);       // [D+3] weight of bin k into low-res cell (k-1)/3 + 1
    float* carry_all = dB + (D + 3);
    for (int k = threadIdx.x; k < D + 3; k += 128) {
        float l0 = 0.f, l1 = 0.f, wa = 0.f, wb = 0.f;
        if (k < D) {
            int t0, t1;
            src_index<true>(scale, k, Dl, t0, t1, l0, l1);
            const int jb = k == 0 ? 0 : (k - 1) / 3;
            wa = (t0 == jb ? l0 : 0.f) + (t1 == jb ? l1 : 0.f);
            wb = (t0 == jb + 1 ? l0 : 0.f) + (t1 == jb + 1 ? l1 : 0.f);
        }
        lz[k] = l1; dA[k] = wa; dB[k] = wb;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int task = blockIdx.y * 4 + warp;
    if (task >= n_tasks) return;
    float* carry = carry_all + warp * (J * 32) + lane;  // [J] per lane, stride 32
    const int tile = task / nJ, jc = task - tile * nJ;
    const int r0 = tile * TR, r1 = min(r0 + TR, Hl);
    const int j0 = jc * J, j1 = min(j0 + J, Dl);
    const int b = blockIdx.z;
    const int c_raw = blockIdx.x * 31 + lane - 1;
    const bool lane_on = c_raw <= Wl - 1;
    const int c = min(c_raw, Wl - 1);

    X3Axis aw;
    x3_axis(scale, c, Wl, aw);
    if (!lane_on) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { aw.wa[i] = 0.f; aw.wb[i] = 0.f; aw.valid[i] = false; }
    }
    const size_t plane = (size_t)Hl * Wl;
    const size_t img = (size_t)H * W;
    const float* base = cost + (size_t)b * Dl * plane;
    const float* gd = gdisp + (size_t)b * img;
    const float* dp = disp + (size_t)b * img;
    const float* sm = stats + (size_t)b * 2 * img;
    float* gout = gcost + (size_t)b * Dl * plane;
    const int jb0 = max(j0 - 1, 0);

    for (int rb = r0 - 1; rb < r1; ++rb) {
        X3Axis ah;
        x3_axis(scale, rb, Hl, ah);
        float hs0[3], hs1[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { hs0[i] = ah.l0[i] * kX3NegLog2e; hs1[i] = ah.l1[i] * kX3NegLog2e; }
        float m[9], gneg[9], dsp[9], a[9], nxt[9];
#pragma unroll
        for (int ph = 0; ph < 3; ++ph)
#pragma unroll
            for (int pw = 0; pw < 3; ++pw) {
                const int i = ph * 3 + pw;
                const size_t o = (size_t)ah.idx[ph] * W + aw.idx[pw];   // clamped -> always a real pixel
                const bool ok = ah.valid[ph] && aw.valid[pw];
                m[i] = __ldg(sm + o);
                dsp[i] = __ldg(dp + o);
                gneg[i] = ok ? -__ldg(gd + o) * __ldg(sm + img + o) : 0.f;
            }
        // Ground-truth masks are sparse and spatially coherent (no LiDAR returns in the sky, occlusions): if no
        // pixel of the warp's 32 blocks carries an upstream gradient, the whole block row contributes zeros --
        // skip its exp2 work but keep the carries and the stores going.
        float gabs = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) gabs = fmaxf(gabs, fabsf(gneg[i]));
        if (!__any_sync(0xffffffffu, gabs != 0.f)) {
            for (int jb = max(jb0, j0); jb < j1; ++jb) {
                const int jj = jb - j0;
                if (rb >= r0 && lane >= 1 && lane_on) gout[(size_t)jb * plane + (size_t)rb * Wl + c] = carry[jj * 32];
                carry[jj * 32] = 0.f;
            }
            continue;
        }
        X3Loader ld;
        ld.init(base, Wl, plane, Dl, jb0, ah.lo0, ah.lo1, aw.lo0, aw.lo1);
        x3_blend9_rel(ld.v, aw.l0, aw.l1, hs0, hs1, m, a);
        ld.next();
        float pA = 0.f, pB = 0.f;  // "B" bin-part of the previous k-block, split by row part

        for (int jb = jb0; jb < j1; ++jb) {
            x3_blend9_rel(ld.v, aw.l0, aw.l1, hs0, hs1, m, nxt);
            ld.next();
            const int k1 = 3 * jb + 1;
            const float lz1 = lz[k1], lz2 = lz[k1 + 1], lz3 = lz[k1 + 2];
            const float a1 = dA[k1], a2 = dA[k1 + 1], a3 = dA[k1 + 2];
            const float b1 = dB[k1], b2 = dB[k1 + 1], b3 = dB[k1 + 2];
            const float kf = (float)k1;
            float g0[9], g1[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const float t = nxt[i];
                const float dlt = t - a[i];
                const float dk = kf - dsp[i];
                float s0 = 0.f, s1 = 0.f;
                if (jb == 0) {  // full-res bin 0 rides with k-block 0
                    const float u = ex2_approx(a[i]) * (dk - 1.f);
                    s0 = dA[0] * u;
                    s1 = dB[0] * u;
                }
                const float u1 = ex2_approx(__fmaf_rn(lz1, dlt, a[i])) * dk;
                const float u2 = ex2_approx(__fmaf_rn(lz2, dlt, a[i])) * (dk + 1.f);
                const float u3 = ex2_approx(__fmaf_rn(lz3, dlt, a[i])) * (dk + 2.f);
                s0 = __fmaf_rn(a1, u1, s0); s1 = __fmaf_rn(b1, u1, s1);
                s0 = __fmaf_rn(a2, u2, s0); s1 = __fmaf_rn(b2, u2, s1);
                s0 = __fmaf_rn(a3, u3, s0); s1 = __fmaf_rn(b3, u3, s1);
                g0[i] = gneg[i] * s0;
                g1[i] = gneg[i] * s1;
                a[i] = t;
            }
            // separable transpose of the bilinear blend: rows (A -> cell row rb, B -> rb+1), then columns
            float v0A, v0B, v1A, v1B;
            {
                float tAA = 0.f, tAB = 0.f, tBA = 0.f, tBB = 0.f;
#pragma unroll
                for (int pw = 0; pw < 3; ++pw) {
                    const float uA = __fmaf_rn(ah.wa[2], g0[6 + pw], __fmaf_rn(ah.wa[1], g0[3 + pw], ah.wa[0] * g0[pw]));
                    const float uB = __fmaf_rn(ah.wb[2], g0[6 + pw], __fmaf_rn(ah.wb[1], g0[3 + pw], ah.wb[0] * g0[pw]));
                    tAA = __fmaf_rn(aw.wa[pw], uA, tAA); tAB = __fmaf_rn(aw.wb[pw], uA, tAB);
                    tBA = __fmaf_rn(aw.wa[pw], uB, tBA); tBB = __fmaf_rn(aw.wb[pw], uB, tBB);
                }
                v0A = tAA + __shfl_up_sync(0xffffffffu, tAB, 1);
                v0B = tBA + __shfl_up_sync(0xffffffffu, tBB, 1);
            }
            {
                float tAA = 0.f, tAB = 0.f, tBA = 0.f, tBB = 0.f;
#pragma unroll
                for (int pw = 0; pw < 3; ++pw) {
                    const float uA = __fmaf_rn(ah.wa[2], g1[6 + pw], __fmaf_rn(ah.wa[1], g1[3 + pw], ah.wa[0] * g1[pw]));
                    const float uB = __fmaf_rn(ah.wb[2], g1[6 + pw], __fmaf_rn(ah.wb[1], g1[3 + pw], ah.wb[0] * g1[pw]));
                    tAA = __fmaf_rn(aw.wa[pw], uA, tAA); tAB = __fmaf_rn(aw.wb[pw], uA, tAB);
                    tBA = __fmaf_rn(aw.wa[pw], uB, tBA); tBB = __fmaf_rn(aw.wb[pw], uB, tBB);
                }
                v1A = tAA + __shfl_up_sync(0xffffffffu, tAB, 1);
                v1B = tBA + __shfl_up_sync(0xffffffffu, tBB, 1);
            }
            if (jb >= j0) {
                const int jj = jb - j0;
                const float rowA = v0A + pA;   // cell (jb, rb): own k-block's A part + previous k-block's B part
                const float rowB = v0B + pB;   // cell (jb, rb+1)
                if (rb >= r0) {
                    const float out = rowA + carry[jj * 32];
                    if (lane >= 1 && lane_on) gout[(size_t)jb * plane + (size_t)rb * Wl + c] = out;
                }
                carry[jj * 32] = rowB;
            }
            pA = v1A;
            pB = v1B;
        }
    }
}

}  // namespace rag
