// Backward of `last_3_3d`: the bias-free Conv3d C -> 1, 3x3x3, pad 1 that produces the head's
// input (src/models/rag_model.py:269, applied at :361-365; src/automl/operations_3d.py:31-47 with bn=False, relu=False).
//     gin[b,c,d,h,w]      = sum_{kd,kh,kw} W[0,c,kd,kh,kw] * g[b,0,d-kd+1,h-kh+1,w-kw+1]          (data gradient)
//     gW[0,c,kd,kh,kw]    = sum_{b,d,h,w}  in[b,c,d+kd-1,h+kh-1,w+kw-1] * g[b,0,d,h,w]            (weight gradient)
// Both read the ONE-channel upstream gradient g through the same 3x3x6 register window per thread (4 adjacent w) and both are
// MARCH kernels: a warp owns a 4 (h) x 32 (w) column of 16 planes and walks it along d with the window ROLLING in registers
// (one new plane = three row loads per step), fed by a warp-private cp.async ring, no block barrier.  The data gradient re-uses
// the window for all C channels (1296 FMAs per step, output streamed with 128-bit stores: the C-times larger tensor is
// written once); the weight gradient keeps 2 x 27 running sums per thread, reduced once per CTA (fixed order, fp64 final pass:
// deterministic).  cuDNN has no good kernel for a 1-channel side.
#include <cuda_pipeline.h>

#include "common.cuh"

namespace rag {

// Weight gradient: gW[c][k] = sum_p in[c][p] g[p - k + 1]; grid (n_groups, ceil(C/2)), 128 threads; a CTA accumulates the 27 sums
// of channels 2cp and 2cp+1 (one g window feeds 216 FMAs instead of 108) over its share of the positions and writes one
// partial per channel and tap: part [C][n_groups][27], reduced by the final kernel in a fixed order (fp64): deterministic.
// History: a tile kernel (4 x 8 x 32 positions per CTA step, g tile double-buffered by cp.async, 0.353 -> 0.179 ms at B=4
// 288x576 over the round) whose ncu capture showed 39 M shared-memory wavefronts in 50 M SM cycles -- every thread re-read
// a 3x3x6 window of g per tile (54 values for 216 FMAs; the 18 scalar loads among
// them are 4-way bank conflicts) behind two block barriers.  Here a WARP owns a 4 (h) x 32 (w) column of 32 input planes of
// one pair and walks it along d: the thread's g window ROLLS (three planes of 3 x 6 values in registers; one new plane = three
// row loads per step, the oldest plane's registers are overwritten -- the step is unrolled by three so that the roles rotate
// without moves), the neighbours' elements come by shuffle, and the g plane plus the thread's two input vectors of the step
// arrive through a warp-private cp.async ring (no block barrier).  A g row in the ring mirrors its global 128-byte line
// (columns w0..w0+31) and keeps the two halo vectors in a second chunk at a position that rotates with the row (the four
// rows of a warp's edge loads fall into different banks).  Warps take items warp, warp + n_warps, ...; partial sums as above.
constexpr int kLbGroups = 296;                                // CTAs per channel pair = partial sums per channel and tap
constexpr int kWmDS = 32, kWmR = 4;                           // input planes per item (16: 0.121 ms, 32: 0.115 ms), ring slots per warp
constexpr int kWmRowF = 64, kWmG = 6 * kWmRowF, kWmSlot = kWmG + 2 * 32 * 4;   // floats: g rows | x vectors [2 channels][32 lanes]
__device__ __forceinline__ int wm_edge(int row) { return 32 + 8 * (row & 3); }   // + 0: columns w0+32..35, + 4: columns w0-4..w0-1

__global__ void __launch_bounds__(128, 3)
conv3d_c1_bwd_weight_march_kernel(const float* __restrict__ g, const float* __restrict__ in, float* __restrict__ part,
                                  int B, int C, int D, int H, int W, int n_w32, int n_h4, int n_ds) {
    extern __shared__ __align__(128) float wm_smem[];
    __shared__ float red[4][54];
    const int lane = threadIdx.x & 31;
    // the warp index as a value the compiler knows is warp-uniform (from threadIdx.x >> 5 everything a warp decides per item looks
    // divergent to it: every shuffle gets a convergence barrier, the bookkeeping stays out of the uniform registers)
    const int warp = (__ballot_sync(0xffffffffu, threadIdx.x >= 32) != 0u) + (__ballot_sync(0xffffffffu, threadIdx.x >= 64) != 0u) +
                     (__ballot_sync(0xffffffffu, threadIdx.x >= 96) != 0u);
    const int c0 = 2 * blockIdx.y, n_groups = gridDim.x;
    const bool two = c0 + 1 < C;
    const int tw = lane & 7, th = lane >> 3;
    const size_t vol = (size_t)D * H * W;
    const int n_items = B * n_ds * n_h4 * n_w32, stride = n_groups * 4;
    float* ring = wm_smem + warp * (kWmR * kWmSlot);

    // the lane's copies of a unit: two g vectors (k = 0: rows 0-3, eight lanes per row; k = 1: rows 4-5 on lanes 0-15, the
    // twelve halo vectors on lanes 16-27) and its own input vector of both channels
    const bool k1_main = lane < 16, k1_on = lane < 28;
    const int r0 = lane >> 3, r1 = k1_main ? 4 + (lane >> 3) : (lane - 16) >> 1;
    const int cq0 = 4 * (lane & 7), cq1 = k1_main ? 4 * (lane & 7) : ((lane & 1) ? -4 : 32);          // column offset from w0
    const int so0 = r0 * kWmRowF + 4 * (lane & 7), so1 = r1 * kWmRowF + (k1_main ? 4 * (lane & 7) : wm_edge(r1) + ((lane & 1) ? 4 : 0));

    // producer state: unit pu of item pi; a unit = g plane d0 - 1 + pu (p_gd) and the input plane d0 + pu - 2.  The three source
    // pointers are set per item and advance by one plane per unit (they may point outside the tensors while the plane or the
    // lane's row / column is outside: then the copy is a zero fill that reads nothing)
    const ptrdiff_t HW = (ptrdiff_t)H * W;
    int pi = blockIdx.x * 4 + warp, pu = 0, ps = 0, p_nu = 0, p_gd = 0;
    const float *pg0 = g, *pg1 = g, *px = in;
    bool v0 = false, v1 = false, vx = false;
    auto decode = [&](int item, int& b, int& d0, int& h0, int& w0, int& nu) {
        int t = item;
        const int wq = t % n_w32; t /= n_w32;
        const int hq = t % n_h4; t /= n_h4;
        const int dq = t % n_ds; b = t / n_ds;
        d0 = dq * kWmDS; h0 = hq * 4; w0 = wq * 32;
        nu = min(kWmDS, D - d0) + 2;
    };
    auto p_begin = [&]() {
        int b, d0, h0, w0;
        decode(pi, b, d0, h0, w0, p_nu);
        p_gd = d0 - 1;
        const float* gb = g + (ptrdiff_t)b * (ptrdiff_t)vol + (ptrdiff_t)p_gd * HW;
        const int gh0 = h0 - 1 + r0, gw0 = w0 + cq0, gh1 = h0 - 1 + r1, gw1 = w0 + cq1, h = h0 + th, w = w0 + 4 * tw;
        pg0 = gb + (ptrdiff_t)gh0 * W + gw0;
        pg1 = gb + (ptrdiff_t)gh1 * W + gw1;
        v0 = gh0 >= 0 && gh0 < H && gw0 < W;                                        // gw0 >= 0; W % 4 == 0
        v1 = k1_on && gh1 >= 0 && gh1 < H && gw1 >= 0 && gw1 < W;
        px = in + ((ptrdiff_t)b * C + c0) * (ptrdiff_t)vol + (ptrdiff_t)(d0 - 2) * HW + (ptrdiff_t)h * W + w;
        vx = h < H && w < W;
    };
    if (pi < n_items) p_begin();
    auto copy_next = [&]() {
        if (pi < n_items) {
            float* slot = ring + ps * kWmSlot;
            const bool pl = (unsigned)p_gd < (unsigned)D;
            const bool ok0 = pl && v0, ok1 = pl && v1, okx = vx && pu >= 2, okx1 = okx && two;
            __pipeline_memcpy_async(slot + so0, ok0 ? pg0 : g, 16, ok0 ? 0 : 16);
            if (k1_on) __pipeline_memcpy_async(slot + so1, ok1 ? pg1 : g, 16, ok1 ? 0 : 16);
            __pipeline_memcpy_async(slot + kWmG + lane * 4, okx ? px : in, 16, okx ? 0 : 16);
            __pipeline_memcpy_async(slot + kWmG + 128 + lane * 4, okx1 ? px + vol : in, 16, okx1 ? 0 : 16);
            pg0 += HW; pg1 += HW; px += HW; ++p_gd;
            if (++ps == kWmR) ps = 0;
            if (++pu == p_nu) {
                pu = 0;
                pi += stride;
                if (pi < n_items) p_begin();
            }
        }
        __pipeline_commit();
    };
#pragma unroll
    for (int i = 0; i < kWmR - 1; ++i) copy_next();

    float acc[2][27];
#pragma unroll
    for (int k = 0; k < 27; ++k) { acc[0][k] = 0.f; acc[1][k] = 0.f; }
    float W0[3][6], W1[3][6], W2[3][6];                         // g planes of the rolling window: [row h-1, h, h+1][column -1 .. 4]
    int cs = 0;
    const bool edge = tw == 0 || tw == 7;
    // one unit: the new g plane goes into `nw`; with (old, mid, nw) = g planes d-1, d, d+1 the input plane d is accumulated
    auto unit = [&](bool fma, float (&nw)[3][6], const float (&old)[3][6], const float (&mid)[3][6]) {
        __pipeline_wait_prior(kWmR - 2);
        __syncwarp();
        copy_next();
        const float* slot = ring + cs * kWmSlot;
        if (++cs == kWmR) cs = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int row = th + j;
            const float* p = slot + row * kWmRowF;
            const float4 m = *reinterpret_cast<const float4*>(p + 4 * tw);
            const float e = edge ? p[wm_edge(row) + (tw == 0 ? 7 : 0)] : 0.f;
            const float l = __shfl_up_sync(0xffffffffu, m.w, 1, 8), r = __shfl_down_sync(0xffffffffu, m.x, 1, 8);
            nw[j][0] = tw == 0 ? e : l;
            nw[j][1] = m.x; nw[j][2] = m.y; nw[j][3] = m.z; nw[j][4] = m.w;
            nw[j][5] = tw == 7 ? e : r;
        }
        if (!fma) return;
        const float4 x0 = *reinterpret_cast<const float4*>(slot + kWmG + lane * 4);
        const float4 x1 = *reinterpret_cast<const float4*>(slot + kWmG + 128 + lane * 4);
        // gW[k] += in[p] g[p - k + 1]: for input column i (0..3) tap (kd,kh,kw) pairs with plane 2-kd, row 2-kh, element i + 2 - kw.
        // Column-major: 54 independent FMAs per input column (tap-major order chains four dependent FMAs per accumulator and,
        // with three warps per scheduler, stalls on their latency)
        const float xa[4] = {x0.x, x0.y, x0.z, x0.w}, xb[4] = {x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int kd = 0; kd < 3; ++kd)
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float* r = kd == 0 ? nw[2 - kh] : (kd == 1 ? mid[2 - kh] : old[2 - kh]);
                        const int k = (kd * 3 + kh) * 3 + kw;
                        acc[0][k] = __fmaf_rn(xa[i], r[i + 2 - kw], acc[0][k]);
                        acc[1][k] = __fmaf_rn(xb[i], r[i + 2 - kw], acc[1][k]);
                    }
    };
    for (int item = blockIdx.x * 4 + warp; item < n_items; item += stride) {
        int b, d0, h0, w0, nu;
        decode(item, b, d0, h0, w0, nu);
        for (int u = 0; u < nu; u += 3) {                         // unit u: planes (u-2, u-1, u) = (old, mid, new)
            unit(u >= 2, W0, W1, W2);
            if (u + 1 < nu) unit(u + 1 >= 2, W1, W2, W0);
            if (u + 2 < nu) unit(true, W2, W0, W1);
        }
    }
    __pipeline_wait_prior(0);
    // CTA reduction, fixed order: a value-halving butterfly over the lanes (at every stage a lane hands half of its values to its
    // partner and adds the half it receives: 62 shuffles for 54 values instead of 270; lane L ends with the sums of values 2L
    // and 2L+1), then the 4 warps in order
    {
        float v[64];
#pragma unroll
        for (int k = 0; k < 64; ++k) v[k] = k < 54 ? acc[k / 27][k % 27] : 0.f;
#pragma unroll
        for (int st = 0; st < 5; ++st) {
            const int m = 16 >> st, n = 32 >> st;                 // partner distance, values kept after this stage
            const bool up = (lane & m) != 0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
                const float keep = up ? v[i + n] : v[i], send = up ? v[i] : v[i + n];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
            }
        }
        if (2 * lane < 54) red[warp][2 * lane] = v[0];
        if (2 * lane + 1 < 54) red[warp][2 * lane + 1] = v[1];
    }
    __syncthreads();
    if (threadIdx.x < 54 && (threadIdx.x < 27 || two)) {
        float v = 0.f;
#pragma unroll
        for (int wv = 0; wv < 4; ++wv) v += red[wv][threadIdx.x];
        const int c = c0 + threadIdx.x / 27, k = threadIdx.x % 27;
        part[((size_t)c * n_groups + blockIdx.x) * 27 + k] = v;
    }
}

// The data gradient (same column walk and ring as the weight gradient; an item per warp): per step the rolling window of g
// gives, for each of the C channels, 108 FMAs against that channel's 27 weights (seven broadcast LDS.128 from the CTA's padded
// table) and one streamed 128-bit store.  Its predecessor, a 4 x 8 x 32 tile per CTA, spent a quarter of its instructions in
// the per-tile prologue (staging through registers, weight table, 27 window loads) for 1296 FMAs per thread: 0.088 -> 0.079 ms
// at B=4 288x576, 0.489 -> 0.405 ms at B=8 480x960.
// grid ceil(items / 4), 128 threads.  smem: ring [4 warps][kWmR][6 x 64] | weights [C][28]
constexpr int kDmSlot = kWmG, kDmDS = 16;                  // 16 planes per item here: twice the warps (0.079 vs 0.090 ms with 32)
__global__ void __launch_bounds__(128, 4)
conv3d_c1_bwd_data_march_kernel(const float* __restrict__ g, const float* __restrict__ wgt, float* __restrict__ gin,
                                int B, int C, int D, int H, int W, int n_w32, int n_h4, int n_ds) {
    extern __shared__ __align__(128) float dm_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = (__ballot_sync(0xffffffffu, threadIdx.x >= 32) != 0u) + (__ballot_sync(0xffffffffu, threadIdx.x >= 64) != 0u) +
                     (__ballot_sync(0xffffffffu, threadIdx.x >= 96) != 0u);
    float* ring = dm_smem + warp * (kWmR * kDmSlot);
    float* ws = dm_smem + 4 * kWmR * kDmSlot;
    for (int i = threadIdx.x; i < C * 28; i += 128) {
        const int c = i / 28, k = i - c * 28;
        ws[i] = k < 27 ? __ldg(wgt + c * 27 + k) : 0.f;
    }
    __syncthreads();
    const int n_items = B * n_ds * n_h4 * n_w32;
    const int item = blockIdx.x * 4 + warp;
    if (item >= n_items) return;
    const int tw = lane & 7, th = lane >> 3;
    const size_t vol = (size_t)D * H * W;
    const ptrdiff_t HW = (ptrdiff_t)H * W;
    int t = item;
    const int wq = t % n_w32; t /= n_w32;
    const int hq = t % n_h4; t /= n_h4;
    const int dq = t % n_ds, b = t / n_ds;
    const int d0 = dq * kDmDS, h0 = hq * 4, w0 = wq * 32, nu = min(kDmDS, D - d0) + 2;

    const bool k1_main = lane < 16, k1_on = lane < 28;
    const int r0 = lane >> 3, r1 = k1_main ? 4 + (lane >> 3) : (lane - 16) >> 1;
    const int cq0 = 4 * (lane & 7), cq1 = k1_main ? 4 * (lane & 7) : ((lane & 1) ? -4 : 32);
    const int so0 = r0 * kWmRowF + 4 * (lane & 7), so1 = r1 * kWmRowF + (k1_main ? 4 * (lane & 7) : wm_edge(r1) + ((lane & 1) ? 4 : 0));
    const int gh0 = h0 - 1 + r0, gw0 = w0 + cq0, gh1 = h0 - 1 + r1, gw1 = w0 + cq1;
    const float* gb = g + (ptrdiff_t)b * (ptrdiff_t)vol + (ptrdiff_t)(d0 - 1) * HW;
    const float *pg0 = gb + (ptrdiff_t)gh0 * W + gw0, *pg1 = gb + (ptrdiff_t)gh1 * W + gw1;
    const bool v0 = gh0 >= 0 && gh0 < H && gw0 < W, v1 = k1_on && gh1 >= 0 && gh1 < H && gw1 >= 0 && gw1 < W;
    int pu = 0, ps = 0;
    auto copy_next = [&]() {
        if (pu < nu) {
            float* slot = ring + ps * kDmSlot;
            const bool pl = (unsigned)(d0 - 1 + pu) < (unsigned)D;
            const bool ok0 = pl && v0, ok1 = pl && v1;
            __pipeline_memcpy_async(slot + so0, ok0 ? pg0 : g, 16, ok0 ? 0 : 16);
            if (k1_on) __pipeline_memcpy_async(slot + so1, ok1 ? pg1 : g, 16, ok1 ? 0 : 16);
            pg0 += HW; pg1 += HW; ++pu;
            if (++ps == kWmR) ps = 0;
        }
        __pipeline_commit();
    };
#pragma unroll
    for (int i = 0; i < kWmR - 1; ++i) copy_next();

    const int h = h0 + th, w = w0 + 4 * tw;
    const bool on = h < H && w < W;
    float* o = gin + (size_t)b * C * vol + (size_t)d0 * H * W + (size_t)(on ? h : 0) * W + (on ? w : 0);   // plane d0, channel 0
    float W0[3][6], W1[3][6], W2[3][6];
    int cs = 0;
    const bool edge = tw == 0 || tw == 7;
    auto unit = [&](bool out, float (&nw)[3][6], const float (&old)[3][6], const float (&mid)[3][6]) {
        __pipeline_wait_prior(kWmR - 2);
        __syncwarp();
        copy_next();
        const float* slot = ring + cs * kDmSlot;
        if (++cs == kWmR) cs = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int row = th + j;
            const float* p = slot + row * kWmRowF;
            const float4 m = *reinterpret_cast<const float4*>(p + 4 * tw);
            const float e = edge ? p[wm_edge(row) + (tw == 0 ? 7 : 0)] : 0.f;
            const float l = __shfl_up_sync(0xffffffffu, m.w, 1, 8), r = __shfl_down_sync(0xffffffffu, m.x, 1, 8);
            nw[j][0] = tw == 0 ? e : l;
            nw[j][1] = m.x; nw[j][2] = m.y; nw[j][3] = m.z; nw[j][4] = m.w;
            nw[j][5] = tw == 7 ? e : r;
        }
        if (!out) return;
        // gin[p] = sum_k W[k] g[p - k + 1]: tap (kd,kh,kw) reads plane 2-kd (old / mid / nw = g at d-1 / d / d+1), row 2-kh, element i + 2 - kw
        for (int c = 0; c < C; ++c) {
            float wc[28];
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(ws + c * 28 + 4 * q);
                wc[4 * q] = v.x; wc[4 * q + 1] = v.y; wc[4 * q + 2] = v.z; wc[4 * q + 3] = v.w;
            }
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int kd = 0; kd < 3; ++kd)
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float wv = wc[(kd * 3 + kh) * 3 + kw];
                        const float* r = kd == 0 ? nw[2 - kh] : (kd == 1 ? mid[2 - kh] : old[2 - kh]);
                        acc.x = __fmaf_rn(wv, r[0 + 2 - kw], acc.x);
                        acc.y = __fmaf_rn(wv, r[1 + 2 - kw], acc.y);
                        acc.z = __fmaf_rn(wv, r[2 + 2 - kw], acc.z);
                        acc.w = __fmaf_rn(wv, r[3 + 2 - kw], acc.w);
                    }
            if (on) st_stream(reinterpret_cast<float4*>(o + (size_t)c * vol), acc);
        }
        o += HW;
    };
    for (int u = 0; u < nu; u += 3) {                             // unit u: planes (u-2, u-1, u) = (old, mid, new) -> output plane d0 + u - 2
        unit(u >= 2, W0, W1, W2);
        if (u + 1 < nu) unit(u + 1 >= 2, W1, W2, W0);
        if (u + 2 < nu) unit(true, W2, W0, W1);
    }
    __pipeline_wait_prior(0);
}

// gw[c][k] = sum over groups of part[c][grp][k], fixed order, fp64.  grid C; 27 warps (one per tap): the lanes take every
// 32nd group, then a shuffle tree in a fixed order (the first version walked the 296 partials of a tap with ONE thread:
// 23 us of dependent loads).
__global__ void __launch_bounds__(27 * 32)
conv3d_c1_bwd_weight_final_kernel(const float* __restrict__ part, float* __restrict__ gw, int n_groups) {
    const int c = blockIdx.x, k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double a = 0.0;
    for (int i = lane; i < n_groups; i += 32) a += (double)part[((size_t)c * n_groups + i) * 27 + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if (lane == 0) gw[c * 27 + k] = (float)a;
}


size_t conv3d_c1_bwd_workspace_bytes(int C) { return C > 0 ? (size_t)C * kLbGroups * 27 * sizeof(float) : 0; }

// gin [B,C,D,H,W] (nullable; needs w) and gw [1,C,3,3,3] (nullable; needs in + workspace) from g [B,1,D,H,W]
int conv3d_c1_bwd(const float* g, const float* in, const float* w, float* gin, float* gw, float* workspace,
                  int B, int C, int D, int H, int W, cudaStream_t st) {
    if (!g) return fail(RAG_E_NULL, "conv3d_c1_bwd: null pointer");
    if (gin && !w) return fail(RAG_E_NULL, "conv3d_c1_bwd: the data gradient needs the weight");
    if (gw && (!in || !workspace)) return fail(RAG_E_NULL, "conv3d_c1_bwd: the weight gradient needs the layer input and a workspace");
    if (B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return fail(RAG_E_SHAPE, "conv3d_c1_bwd: non-positive dimension");
    if (W % 4 != 0) return fail(RAG_E_SHAPE, "conv3d_c1_bwd: W=%d must be a multiple of 4", W);
    if ((size_t)D * H * W >= ((size_t)1 << 31) || C > 1024) return fail(RAG_E_SHAPE, "conv3d_c1_bwd: D*H*W must be < 2^31 and C <= 1024");
    if (!aligned(g, 16) || (gin && !aligned(gin, 16)) || (in && !aligned(in, 16))) return fail(RAG_E_ALIGN, "conv3d_c1_bwd: g, in, gin must be 16-byte aligned");
    const int n_w32 = (W + 31) / 32, n_h4 = (H + 3) / 4, n_ds = (D + kWmDS - 1) / kWmDS, n_dsd = (D + kDmDS - 1) / kDmDS;
    const long long items = (long long)B * n_dsd * n_h4 * n_w32;              // data gradient: 4 (h) x 32 (w) columns of 16 planes
    if (items >= (1LL << 31)) return fail(RAG_E_SHAPE, "conv3d_c1_bwd: too many tiles");
    if (gin) {
        const size_t dsmem = ((size_t)4 * kWmR * kDmSlot + (size_t)C * 28) * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(conv3d_c1_bwd_data_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem);
        if (e != cudaSuccess) return fail((int)e, "conv3d_c1_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        conv3d_c1_bwd_data_march_kernel<<<(unsigned)((items + 3) / 4), 128, dsmem, st>>>(g, w, gin, B, C, D, H, W, n_w32, n_h4, n_dsd);
        if (int rc = check_launch("conv3d_c1_bwd(data)")) return rc;
    }
    if (gw) {
        const size_t wsmem = (size_t)4 * kWmR * kWmSlot * sizeof(float);
        conv3d_c1_bwd_weight_march_kernel<<<dim3(kLbGroups, (C + 1) / 2), 128, wsmem, st>>>(g, in, workspace, B, C, D, H, W, n_w32, n_h4, n_ds);
        if (int rc = check_launch("conv3d_c1_bwd(weight)")) return rc;
        conv3d_c1_bwd_weight_final_kernel<<<C, 27 * 32, 0, st>>>(workspace, gw, kLbGroups);
        if (int rc = check_launch("conv3d_c1_bwd(weight final)")) return rc;
    }
    return RAG_OK;
}

}  // namespace rag
