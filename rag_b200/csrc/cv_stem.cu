// Fused cost volume + first Matching-Net convolution.
//
// Reference: the volume built at src/models/rag_model.py:375-383 is consumed by
// `self.stem3d0[i](cost)` (rag_model.py:341) = ConvBR_3d(2C -> O, 3x3x3, stride 1, pad 1):
// Conv3d(bias=False) -> BatchNorm3d -> ReLU (src/automl/operations_3d.py:31-47).  Through cuDNN that
// layer costs 6.3 ms (TF32) / 63 ms (fp32) per 8 pairs at 480x960 on a B200 -- 15x the cost-volume kernel --
// because it reads the 2.5 GB volume and does 51 GFLOP per pair.
//
// The volume is not a general tensor: cost[c,d,h,w] = x[c,h,w]*(w>=d) and cost[C+c,d,h,w] = y[c,h,w-d]*(w>=d).
// Substituting into the convolution, with d' = d+kd-1, w' = w+kw-1:
//   out[o,d,h,w] = sum_{kd,kw} [0<=d'<Df][0<=w'<Wf][w'>=d'] ( A[kd,kw][o,h,w'] + R[kd,kw][o,h,w'-d'] )
//   A[kd,kw][o,h,w'] = sum_{c,kh} W[o,  c,kd,kh,kw] x[c,h+kh-1,w']        (nine 2-D row maps from the left features)
//   R[kd,kw][o,h,u'] = sum_{c,kh} W[o,C+c,kd,kh,kw] y[c,h+kh-1,u']        (nine from the right features)
// Away from the borders all taps are valid and the double sum collapses to
//   out[o,d,h,w] = LF[h,w] + RF[h,w-d],   LF = sum_{kd,kw} A[kd,kw][.,w+kw-1],  RF = sum_{kd,kw} R[kd,kw][.,u+kw-kd]
// i.e. the 3-D convolution of the 2.5 GB volume is two small 2-D maps broadcast-added along the
// disparity axis: ~10 FMA per output instead of 648, and the volume is never materialised.  Exceptions
// are handled exactly: d = 0 / Df-1 (class variants of LF, RF), the diagonal band -2 <= w-d <= 1 and the
// last column (evaluated tap by tap into small tables), and w-d <= -3 (all taps masked: 0).
//
// One CTA per (b, o, h): phase 1 builds the 18 row maps in shared memory (thread per column), phase 2
// the LF/RF rows (RF as 4 pre-shifted copies so the shifted read is an aligned LDS.128) and the border
// tables, phase 3 streams the Df x Wf output rows with 128-bit stores, optionally applying the folded
// eval-mode BatchNorm (scale, shift) and ReLU.  Roofline: HBM store stream of the [B,O,Df,Hf,Wf] output.
#include <cuda_pipeline.h>

#include "common.cuh"

namespace rag {

constexpr int kStemPad = 4;   // zero columns on each side of every shared-memory row
constexpr int kStemStages = 3; // depth of the per-thread cp.async ring of the packed kernel

// exact tap-by-tap value (conv only) for output (d, w); maps[m] rows are padded by kStemPad zeros
__device__ __forceinline__ float stem_taps(const float* maps, int Wp, int Df, int Wf, int d, int w) {
    float v = 0.f;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
        const int dd = d + kd - 1;
        if (dd < 0 || dd >= Df) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int ww = w + kw - 1;
            if (ww < 0 || ww >= Wf || ww < dd) continue;
            v += maps[(kd * 3 + kw) * Wp + kStemPad + ww] + maps[(9 + kd * 3 + kw) * Wp + kStemPad + ww - dd];
        }
    }
    return v;
}

// grid: x = h, y = o, z = b.  block: NT threads (multiple of 32).
// smem floats: wsm[2*C*3*12] | maps[18][Wp] | LF[3][Wp] | RF[3][4][Wp] | band[Df][4] | lastcol[Df]
template <int C>
__global__ void __launch_bounds__(512, 2)
cv_stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ w,
                   const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                   float* __restrict__ out, int O, int Df, int Hf, int Wf) {
    extern __shared__ __align__(16) float stem_smem[];
    const int Wp = Wf + 2 * kStemPad;
    float* wsm = stem_smem;                       // [side][c][kh][12]: 9 (kd,kw) weights + 3 pad
    float* maps = wsm + 2 * C * 3 * 12;
    float* LF = maps + 18 * Wp;
    float* RF = LF + 3 * Wp;
    float* band = RF + 12 * Wp;
    float* lastcol = band + 4 * Df;
    const int h = blockIdx.x, o = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, NT = blockDim.x;

    // ---- weights of this output channel, regrouped so the 9 (kd,kw) taps of one (side,c,kh) are contiguous ----
    for (int i = tid; i < 2 * C * 3 * 12; i += NT) {
        const int t = i % 12, kh = (i / 12) % 3, c = (i / 36) % C, side = i / (36 * C);
        float v = 0.f;
        if (t < 9) {
            const int kd = t / 3, kw = t % 3;
            v = __ldg(w + ((((size_t)o * 2 * C + side * C + c) * 3 + kd) * 3 + kh) * 3 + kw);
        }
        wsm[i] = v;
    }
    // zero pads of all 33 rows (maps, LF, RF) and the RF rows themselves (their shifted copies leave the first s slots untouched)
    for (int i = tid; i < 33 * 2 * kStemPad; i += NT) {
        const int row = i / (2 * kStemPad), k = i - row * (2 * kStemPad);
        maps[row * Wp + (k < kStemPad ? k : Wf + k)] = 0.f;
    }
    __syncthreads();

    // ---- phase 1: the 18 row maps, one column per thread; x/y rows walked with pointer bumps ----
    const size_t img = (size_t)Hf * Wf;
    const bool up = h > 0, dn = h < Hf - 1;
    for (int col = tid; col < Wf; col += NT) {
        float a[9], r[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { a[i] = 0.f; r[i] = 0.f; }
        const float* px = x + (size_t)b * C * img + (size_t)h * Wf + col;
        const float* py = y + (size_t)b * C * img + (size_t)h * Wf + col;
        const float4* wl = reinterpret_cast<const float4*>(wsm);
        const float4* wr = reinterpret_cast<const float4*>(wsm + C * 36);
#pragma unroll 2
        for (int c = 0; c < C; ++c) {
            // issue the (up to) six loads of this channel first, then the 54 FMAs
            const float x1 = __ldg(px), y1 = __ldg(py);
            const float x0 = up ? __ldg(px - Wf) : 0.f, y0 = up ? __ldg(py - Wf) : 0.f;
            const float x2 = dn ? __ldg(px + Wf) : 0.f, y2 = dn ? __ldg(py + Wf) : 0.f;
            px += img; py += img;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const float xv = kh == 0 ? x0 : kh == 1 ? x1 : x2;
                const float yv = kh == 0 ? y0 : kh == 1 ? y1 : y2;
                const float4 l0 = wl[kh * 3 + 0], l1 = wl[kh * 3 + 1], l2 = wl[kh * 3 + 2];
                const float4 r0 = wr[kh * 3 + 0], r1 = wr[kh * 3 + 1], r2 = wr[kh * 3 + 2];
                a[0] = __fmaf_rn(l0.x, xv, a[0]); a[1] = __fmaf_rn(l0.y, xv, a[1]); a[2] = __fmaf_rn(l0.z, xv, a[2]);
                a[3] = __fmaf_rn(l0.w, xv, a[3]); a[4] = __fmaf_rn(l1.x, xv, a[4]); a[5] = __fmaf_rn(l1.y, xv, a[5]);
                a[6] = __fmaf_rn(l1.z, xv, a[6]); a[7] = __fmaf_rn(l1.w, xv, a[7]); a[8] = __fmaf_rn(l2.x, xv, a[8]);
                r[0] = __fmaf_rn(r0.x, yv, r[0]); r[1] = __fmaf_rn(r0.y, yv, r[1]); r[2] = __fmaf_rn(r0.z, yv, r[2]);
                r[3] = __fmaf_rn(r0.w, yv, r[3]); r[4] = __fmaf_rn(r1.x, yv, r[4]); r[5] = __fmaf_rn(r1.y, yv, r[5]);
                r[6] = __fmaf_rn(r1.z, yv, r[6]); r[7] = __fmaf_rn(r1.w, yv, r[7]); r[8] = __fmaf_rn(r2.x, yv, r[8]);
            }
            wl += 9; wr += 9;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            maps[i * Wp + kStemPad + col] = a[i];
            maps[(9 + i) * Wp + kStemPad + col] = r[i];
        }
    }
    __syncthreads();

    // ---- phase 2: collapsed rows per disparity class (0: d=0 -> kd in {1,2}; 1: interior; 2: d=Df-1 -> kd in {0,1}) ----
    for (int col = tid; col < Wf; col += NT) {
        float S[3], T[3];
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
            const float* ar = maps + (kd * 3) * Wp + kStemPad + col;
            const float* rr = maps + (9 + kd * 3) * Wp + kStemPad + col - kd;
            S[kd] = ar[-1] + ar[Wp] + ar[2 * Wp + 1];
            T[kd] = rr[0] + rr[Wp + 1] + rr[2 * Wp + 2];
        }
        const float lf[3] = {S[1] + S[2], S[0] + S[1] + S[2], S[0] + S[1]};
        const float rf[3] = {T[1] + T[2], T[0] + T[1] + T[2], T[0] + T[1]};
#pragma unroll
        for (int cls = 0; cls < 3; ++cls) {
            LF[cls * Wp + kStemPad + col] = lf[cls];
#pragma unroll
            for (int s = 0; s < 4; ++s) {           // copy s holds RF shifted right by s: copy_s[i] = RF[i - s]
                if (col + s < Wf) RF[(cls * 4 + s) * Wp + kStemPad + col + s] = rf[cls];
                if (col < s) RF[(cls * 4 + s) * Wp + kStemPad + col] = 0.f;
            }
        }
    }
    // border tables, evaluated tap by tap: the diagonal band w - d in [-2, 1] and the last column
    for (int i = tid; i < 5 * Df; i += NT) {
        const int d = i / 5, j = i - d * 5;
        const int col = j < 4 ? d - 2 + j : Wf - 1;
        const float v = (col >= 0 && col < Wf) ? stem_taps(maps, Wp, Df, Wf, d, col) : 0.f;
        if (j < 4) band[d * 4 + j] = v; else lastcol[d] = v;
    }
    __syncthreads();

    // ---- phase 3: stream the Df x Wf rows of out[b,o,:,h,:] ----
    const float sc = scale ? __ldg(scale + o) : 1.f, sh = shift ? __ldg(shift + o) : 0.f;
    auto finish = [&](float v) {
        v = __fmaf_rn(v, sc, sh);
        return relu ? fmaxf(v, 0.f) : v;
    };
    float* ob = out + (((size_t)b * O + o) * Df * Hf + h) * (size_t)Wf;
    const size_t dstride = (size_t)Hf * Wf;
    if ((Wf & 3) == 0) {
        const int Wv = Wf >> 2;
        // items (d, wv) in row-major order dealt round-robin to threads; (d, wv) advanced without divisions
        const int stepd = NT / Wv, stepw = NT - stepd * Wv;
        int d = tid / Wv, wv = tid - d * Wv;
        float* op = ob + (size_t)d * dstride + 4 * wv;
        const size_t opstep = (size_t)stepd * dstride + 4 * stepw;
        while (d < Df) {
            const int w0 = wv << 2;
            const int cls = d == 0 ? 0 : (d == Df - 1 ? 2 : 1);
            const int u0 = w0 - d;
            float4 v;
            if (u0 >= 2 && w0 + 3 <= Wf - 2) {           // all four interior: aligned reads of LF and of RF copy (d & 3)
                const float4 lf = *reinterpret_cast<const float4*>(LF + cls * Wp + kStemPad + w0);
                const float4 rf = *reinterpret_cast<const float4*>(RF + (cls * 4 + (d & 3)) * Wp + kStemPad + w0 - (d & ~3));
                v = make_float4(lf.x + rf.x, lf.y + rf.y, lf.z + rf.z, lf.w + rf.w);
            } else if (u0 + 3 <= -3) {                   // fully masked: every tap sees zeros
                v = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                float e[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int col = w0 + k, u = col - d;
                    if (u <= -3) e[k] = 0.f;
                    else if (u <= 1) e[k] = band[d * 4 + u + 2];
                    else if (col == Wf - 1) e[k] = lastcol[d];
                    else e[k] = LF[cls * Wp + kStemPad + col] + RF[cls * 4 * Wp + kStemPad + u];
                }
                v = make_float4(e[0], e[1], e[2], e[3]);
            }
            v = make_float4(finish(v.x), finish(v.y), finish(v.z), finish(v.w));
            st_stream(reinterpret_cast<float4*>(op), v);
            d += stepd; wv += stepw; op += opstep;
            if (wv >= Wv) { wv -= Wv; ++d; op += dstride - Wf; }
        }
    } else {
        for (int i = tid; i < Df * Wf; i += NT) {
            const int d = i / Wf, col = i - d * Wf, u = col - d;
            const int cls = d == 0 ? 0 : (d == Df - 1 ? 2 : 1);
            float v;
            if (u <= -3) v = 0.f;
            else if (u <= 1) v = band[d * 4 + u + 2];
            else if (col == Wf - 1) v = lastcol[d];
            else v = LF[cls * Wp + kStemPad + col] + RF[cls * 4 * Wp + kStemPad + u];
            st_stream(ob + (size_t)d * dstride + col, finish(v));
        }
    }
}

// Second generation of the fused kernel (default): the first one was issue bound (profiles/: 777 M warp
// instructions per launch, 24 thread instructions per output float against a store-stream budget of ~5.5).
//   * phase 1 on packed FP32: a thread owns a PAIR of adjacent columns of one side (left: x -> A maps,
//     right: y -> R maps); data pairs come from one LDG.64, the nine tap weights are kept in shared
//     memory pre-duplicated as (w,w) so each LDS.128 feeds two FFMA2 -- 9 FFMA2 per (channel, row) for two
//     columns instead of 18 FFMA per column;
//   * phase 3 with the block size a multiple of Wf/4: a thread keeps its column group for every
//     disparity row, so the LF vector lives in registers, the RF read is one LDS.128 through a pointer
//     that moves 16 bytes per step, and no index arithmetic is left in the loop.
// Needs Wf % 4 == 0.  Same shared-memory layout otherwise.
//
// The conv weights are regrouped ONCE per launch by a tiny kernel into the layout phase 1 reads
// ([o][side][c][kh][12]: the nine (kd,kw) weights + padding) instead of by every one
// of the B*O*Hf CTAs (div/mod address arithmetic and scattered loads: 9 % of the samples).  The regrouped copy
// lives in a CALLER-OWNED workspace (rag_cv_stem_workspace_bytes): the library keeps no state, so launches with
// different weights on any streams, graph capture and replays cannot interfere.  Without a workspace every CTA
// regroups its own channel's weights from `w` (the slower path).
constexpr int kStemWPerO = 2 * 12 * 3 * 12;

__global__ void __launch_bounds__(256)
stem_regroup_kernel(const float* __restrict__ w, int C, float* __restrict__ wreg) {
    const int o = blockIdx.x;
    for (int i = threadIdx.x; i < 2 * C * 3 * 12; i += 256) {
        const int t = i % 12, kh = (i / 12) % 3, c = (i / 36) % C, side = i / (36 * C);
        float v = 0.f;
        if (t < 9) {
            const int kd = t / 3, kw = t % 3;
            v = __ldg(w + ((((size_t)o * 2 * C + side * C + c) * 3 + kd) * 3 + kh) * 3 + kw);
        }
        wreg[(size_t)o * kStemWPerO + i] = v;
    }
}

// MOMENTS: instead of streaming the output, reduce (sum z, sum z^2) of this CTA's Df x Wf conv outputs (scale/shift
// ignored) into moments[((b*Hf + h)*O + o)*2 + {0,1}] in fp64 -- the batch-statistics pass of a training-mode
// BatchNorm that never sees the volume or the conv output (oracle.stem_batch_moments_f64).
template <int C, bool RELU, bool MOMENTS = false>
__global__ void __launch_bounds__(512, 2)
cv_stem_fwd2_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ w,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    float* __restrict__ out, int O, int Df, int Hf, int Wf, const float* __restrict__ wreg, double* __restrict__ moments) {
    extern __shared__ __align__(16) float stem_smem[];
    const int Wp = Wf + 2 * kStemPad;
    float* wsm = stem_smem;                       // [side][c][kh][12]: 9 (kd,kw) weights + pad
    float* maps = wsm + 2 * C * 3 * 12;
    float* LF = maps + 18 * Wp;
    float* RF = LF + 3 * Wp;
    float* band = RF + 12 * Wp;
    float* lastcol = band + 4 * Df;
    // per-thread cp.async ring for phase 1: [NT][kStemStages][3 rows] float2, 8-byte aligned.  Thread-major, so that every slot
    // of a thread is a compile-time offset from ONE pointer (stage-major cost an address multiply-add per copy and per load:
    // 137 IMADs around the 324 packed FMAs of a thread-item); a stride of 9 float2 per thread keeps the 64-bit accesses of a
    // half-warp on distinct banks (18 t mod 32 runs through the 16 even banks).
    float2* ring = reinterpret_cast<float2*>(lastcol + ((Df + 3) & ~3));
    float2* const my_ring = ring + (size_t)threadIdx.x * (kStemStages * 3);
    const int h = blockIdx.x, o = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, NT = blockDim.x;
    const size_t img = (size_t)Hf * Wf;
    const bool up = h > 0, dn = h < Hf - 1;
    const int NP = Wf >> 1;

    // phase-1 prefetch: every thread fetches ITS OWN column pair of the three input rows of channel c into its
    // private ring slots, so no block barrier is needed between producer and consumer (same thread).  The three source
    // pointers just walk from channel to channel (no per-copy 64-bit index arithmetic), and a row outside the image is a
    // copy of zero source bytes (zero fill) instead of a predicated copy: every cp.async of the loop is unconditional.
    const float* f_mid = x;                        // row h, row h-1 (or h again), row h+1 (or h again) of the next channel to fetch
    const float* f_up = x;
    const float* f_dn = x;
    const int zf_up = up ? 0 : 8, zf_dn = dn ? 0 : 8;
    auto fetch_begin = [&](const float* src) {
        f_mid = src;
        f_up = up ? src - Wf : src;
        f_dn = dn ? src + Wf : src;
    };
    auto fetch_next = [&](int stage) {             // the next channel's three rows -> ring stage `stage`
        float2* slot = my_ring + stage * 3;
        __pipeline_memcpy_async(slot + 1, f_mid, 8);
        __pipeline_memcpy_async(slot, f_up, 8, zf_up);
        __pipeline_memcpy_async(slot + 2, f_dn, 8, zf_dn);
        f_mid += img; f_up += img; f_dn += img;
    };
    {   // first item of this thread: start its loads before the (latency-bound) weight regrouping below
        const int item = tid;
        if (item < 2 * NP) {
            const int side = item >= NP ? 1 : 0;
            fetch_begin((side ? y : x) + (size_t)b * C * img + (size_t)h * Wf + ((item - side * NP) << 1));
#pragma unroll
            for (int c = 0; c < kStemStages - 1; ++c) { fetch_next(c); __pipeline_commit(); }
        }
    }

    if (wreg != nullptr) {                         // weights already regrouped by stem_regroup_kernel: plain vector copy
        const float4* src = reinterpret_cast<const float4*>(wreg + (size_t)o * kStemWPerO);
        for (int i = tid; i < 2 * C * 3 * 12 / 4; i += NT) reinterpret_cast<float4*>(wsm)[i] = src[i];
    } else {
        for (int i = tid; i < 2 * C * 3 * 12; i += NT) {
            const int t = i % 12, kh = (i / 12) % 3, c = (i / 36) % C, side = i / (36 * C);
            float v = 0.f;
            if (t < 9) {
                const int kd = t / 3, kw = t % 3;
                v = __ldg(w + ((((size_t)o * 2 * C + side * C + c) * 3 + kd) * 3 + kh) * 3 + kw);
            }
            wsm[i] = v;
        }
    }
    for (int i = tid; i < 33 * 2 * kStemPad; i += NT) {
        const int row = i / (2 * kStemPad), k = i - row * (2 * kStemPad);
        maps[row * Wp + (k < kStemPad ? k : Wf + k)] = 0.f;
    }
    __syncthreads();

    // ---- phase 1: items = (side, column pair), input rows through the per-thread cp.async ring ----
    for (int item = tid; item < 2 * NP; item += NT) {
        const int side = item >= NP ? 1 : 0;
        const int col = (item - side * NP) << 1;
        if (item != tid) {                         // later passes (NT < Wf): restart the pipeline for this item
            fetch_begin((side ? y : x) + (size_t)b * C * img + (size_t)h * Wf + col);
#pragma unroll
            for (int c = 0; c < kStemStages - 1; ++c) { fetch_next(c); __pipeline_commit(); }
        }
        const float4* wp = reinterpret_cast<const float4*>(wsm + side * C * 36);
        float2 acc[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (c + kStemStages - 1 < C) fetch_next((c + kStemStages - 1) % kStemStages);
            __pipeline_commit();
            __pipeline_wait_prior(kStemStages - 1);        // channel c has landed
            const float2* slot = my_ring + (c % kStemStages) * 3;
            const float2 v0 = slot[0], v1 = slot[1], v2 = slot[2];        // rows outside the image were zero-filled
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const float2 v = kh == 0 ? v0 : kh == 1 ? v1 : v2;
                // nine (kd,kw) weights of (c,kh) in three vectors; the packed FMA broadcasts the scalar weight
                const float4 q0 = wp[kh * 3 + 0], q1 = wp[kh * 3 + 1], q2 = wp[kh * 3 + 2];
                acc[0] = __ffma2_rn(make_float2(q0.x, q0.x), v, acc[0]); acc[1] = __ffma2_rn(make_float2(q0.y, q0.y), v, acc[1]);
                acc[2] = __ffma2_rn(make_float2(q0.z, q0.z), v, acc[2]); acc[3] = __ffma2_rn(make_float2(q0.w, q0.w), v, acc[3]);
                acc[4] = __ffma2_rn(make_float2(q1.x, q1.x), v, acc[4]); acc[5] = __ffma2_rn(make_float2(q1.y, q1.y), v, acc[5]);
                acc[6] = __ffma2_rn(make_float2(q1.z, q1.z), v, acc[6]); acc[7] = __ffma2_rn(make_float2(q1.w, q1.w), v, acc[7]);
                acc[8] = __ffma2_rn(make_float2(q2.x, q2.x), v, acc[8]);
            }
            wp += 9;
        }
        float* dst = maps + (side * 9) * Wp + kStemPad + col;
#pragma unroll
        for (int i = 0; i < 9; ++i) *reinterpret_cast<float2*>(dst + i * Wp) = acc[i];
    }
    __syncthreads();

    // ---- phase 2 (as in the first kernel), with the folded BatchNorm applied to the rows once instead of to
    // every output: LF' = sc*LF + sh, RF' = sc*RF, so that out = relu?(LF' + RF') ----
    const float sc = (!MOMENTS && scale) ? __ldg(scale + o) : 1.f, sh = (!MOMENTS && shift) ? __ldg(shift + o) : 0.f;
    for (int col = tid; col < Wf; col += NT) {
        float S[3], T[3];
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
            const float* ar = maps + (kd * 3) * Wp + kStemPad + col;
            const float* rr = maps + (9 + kd * 3) * Wp + kStemPad + col - kd;
            S[kd] = ar[-1] + ar[Wp] + ar[2 * Wp + 1];
            T[kd] = rr[0] + rr[Wp + 1] + rr[2 * Wp + 2];
        }
        const float lf[3] = {S[1] + S[2], S[0] + S[1] + S[2], S[0] + S[1]};
        const float rf[3] = {T[1] + T[2], T[0] + T[1] + T[2], T[0] + T[1]};
#pragma unroll
        for (int cls = 0; cls < 3; ++cls) {
            LF[cls * Wp + kStemPad + col] = __fmaf_rn(lf[cls], sc, sh);
            const float rfs = rf[cls] * sc;
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2) {
                if (col + s2 < Wf) RF[(cls * 4 + s2) * Wp + kStemPad + col + s2] = rfs;
                if (col < s2) RF[(cls * 4 + s2) * Wp + kStemPad + col] = 0.f;
            }
        }
    }
    // j-major order: which of the nine taps of a band element survive the mask depends only on j = w - d + 2, so
    // with one j per warp the tap branches of stem_taps are warp-uniform (1, 3, 6, 8 taps for j = 0..3)
    for (int i = tid; i < 5 * Df; i += NT) {
        const int j = i / Df, d = i - j * Df;
        const int col = j < 4 ? d - 2 + j : Wf - 1;
        const float v = (col >= 0 && col < Wf) ? stem_taps(maps, Wp, Df, Wf, d, col) : 0.f;
        if (j < 4) band[d * 4 + j] = v; else lastcol[d] = v;
    }
    __syncthreads();

    if constexpr (MOMENTS) {
        // z(d, w) by the same rule as the output pass: 0 left of the band, the tap tables on the band and in the last
        // column, LF + RF (class by d) elsewhere.  fp64 from the first addition on: the sums feed a variance.
        double s1 = 0.0, s2 = 0.0;
        for (int col = tid; col < Wf; col += NT) {
            for (int d = 0; d < Df; ++d) {
                const int u = col - d;
                float z;
                if (u <= -3) z = 0.f;
                else if (u <= 1) z = band[d * 4 + (u + 2)];
                else if (col == Wf - 1) z = lastcol[d];
                else {
                    const int cls = d == 0 ? 0 : (d == Df - 1 ? 2 : 1);
                    z = LF[cls * Wp + kStemPad + col] + RF[(cls * 4) * Wp + kStemPad + u];
                }
                s1 += (double)z;
                s2 += (double)z * (double)z;
            }
        }
        // block reduction in a fixed order (deterministic) without warp collectives: the block size need not be a
        // multiple of 32
        __shared__ double red1[512], red2[512];
        red1[tid] = s1; red2[tid] = s2;
        __syncthreads();
        if (tid < 32) {
            double a = 0.0, b2 = 0.0;
            for (int i = tid; i < NT; i += 32) { a += red1[i]; b2 += red2[i]; }
            red1[tid] = a; red2[tid] = b2;
        }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0, b2 = 0.0;
            const int nred = NT < 32 ? NT : 32;
            for (int i = 0; i < nred; ++i) { a += red1[i]; b2 += red2[i]; }
            double* m = moments + (((size_t)b * Hf + h) * O + o) * 2;
            m[0] = a; m[1] = b2;
        }
        return;
    }

    // ---- phase 3: thread keeps its 4-column group; rows d = d0, d0 + rows_per_pass, ... ----
    // Main pass is branch-free: every element is LF + RF (masked to 0 where w - d <= -3, i.e. all taps masked);
    // the 5 border elements per disparity row (diagonal band, last column) are patched afterwards from the
    // tap-by-tap tables.
    auto finish = [&](float v) {                   // border elements (unscaled tap sums)
        v = __fmaf_rn(v, sc, sh);
        return RELU ? fmaxf(v, 0.f) : v;
    };
    auto act = [&](float v) { return RELU ? fmaxf(v, 0.f) : v; };   // main pass: rows are already scaled
    const int Wv = Wf >> 2;
    const int rpp = NT / Wv;                       // rows per pass (NT is a multiple of Wv)
    const int d0 = tid / Wv, wv = tid - d0 * Wv, w0 = wv << 2;
    float* obase = out + (((size_t)b * O + o) * Df * Hf + h) * (size_t)Wf;
    const size_t dstride = (size_t)Hf * Wf;
    if (d0 < rpp) {
        float* op = obase + (size_t)d0 * dstride + w0;
        const size_t ostep = (size_t)rpp * dstride;
        const float4 lf1 = *reinterpret_cast<const float4*>(LF + 1 * Wp + kStemPad + w0);
        const float zero_val = finish(0.f);
        // general row: class select, masks for the groups the diagonal crosses, zeros left of it
        auto slow_row = [&](int d) {
            const int cls = d == 0 ? 0 : (d == Df - 1 ? 2 : 1);
            const int u0 = w0 - d;
            float4 v;
            if (u0 + 3 <= -3) {                     // fully masked group (warp-coherent: whole left part of a row)
                v = make_float4(zero_val, zero_val, zero_val, zero_val);
            } else {
                const float4 lf = cls == 1 ? lf1 : *reinterpret_cast<const float4*>(LF + cls * Wp + kStemPad + w0);
                // RF copy (d & 3) at vector offset w0 - 4*(d >> 2); for u0 < 0 this reads the zero pad / garbage that is masked below
                const int ro = max(w0 - (d & ~3), -kStemPad);
                const float4 rf = *reinterpret_cast<const float4*>(RF + (cls * 4 + (d & 3)) * Wp + kStemPad + ro);
                v.x = u0 + 0 >= -2 ? act(lf.x + rf.x) : zero_val;
                v.y = u0 + 1 >= -2 ? act(lf.y + rf.y) : zero_val;
                v.z = u0 + 2 >= -2 ? act(lf.z + rf.z) : zero_val;
                v.w = u0 + 3 >= -2 ? act(lf.w + rf.w) : zero_val;
            }
            st_stream(reinterpret_cast<float4*>(op), v);
        };
        int d = d0;
        if (d == 0) { slow_row(0); d += rpp; op += ostep; }
        // interior rows right of the diagonal band (1 <= d <= Df-2, d <= w0+2: 88 % of all rows at 480x960): no class
        // select, no masks -- one aligned LDS.128, four adds, the activation and the store
        const int fend = min(Df - 2, w0 + 2);
        const float* rf1 = RF + 4 * Wp + kStemPad + w0;
        if ((rpp & 3) == 0) {                      // d & 3 is the same for all rows of this thread: the RF pointer just walks
            const float* rp = rf1 + (d & 3) * Wp - (d & ~3);
            for (; d <= fend; d += rpp, op += ostep, rp -= rpp) {
                const float4 rf = *reinterpret_cast<const float4*>(rp);
                st_stream(reinterpret_cast<float4*>(op), make_float4(act(lf1.x + rf.x), act(lf1.y + rf.y), act(lf1.z + rf.z), act(lf1.w + rf.w)));
            }
        } else {
            for (; d <= fend; d += rpp, op += ostep) {
                const float4 rf = *reinterpret_cast<const float4*>(rf1 + (d & 3) * Wp - (d & ~3));
                st_stream(reinterpret_cast<float4*>(op), make_float4(act(lf1.x + rf.x), act(lf1.y + rf.y), act(lf1.z + rf.z), act(lf1.w + rf.w)));
            }
        }
        for (; d < Df; d += rpp, op += ostep) slow_row(d);
    }
    __syncthreads();                               // main-pass stores of this CTA are ordered before the patches
    for (int i = tid; i < 5 * Df; i += NT) {
        const int d = i / 5, j = i - d * 5;
        const int col = j < 4 ? d - 2 + j : Wf - 1;
        if (col < 0 || col >= Wf) continue;
        if (j == 4 && col - d <= 1) continue;      // last column already covered by the band entries
        const float v = j < 4 ? band[d * 4 + j] : lastcol[d];
        obase[(size_t)d * dstride + col] = finish(v);
    }
}

// Direct evaluation (one thread per output, all 2C*27 taps): any shape incl. Df < 3; cross-check variant.
__global__ void __launch_bounds__(256)
cv_stem_direct_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ w,
                      const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                      float* __restrict__ out, int C, int O, int Df, int Hf, int Wf) {
    const size_t n = (size_t)O * Df * Hf * Wf;
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= n) return;
    const int b = blockIdx.y;
    const int col = idx % Wf, h = (idx / Wf) % Hf, d = (idx / ((size_t)Wf * Hf)) % Df, o = idx / ((size_t)Wf * Hf * Df);
    const size_t img = (size_t)Hf * Wf;
    const float* xb = x + (size_t)b * C * img;
    const float* yb = y + (size_t)b * C * img;
    float acc = 0.f;
    for (int kd = 0; kd < 3; ++kd) {
        const int dd = d + kd - 1;
        if (dd < 0 || dd >= Df) continue;
        for (int kh = 0; kh < 3; ++kh) {
            const int hh = h + kh - 1;
            if (hh < 0 || hh >= Hf) continue;
            for (int kw = 0; kw < 3; ++kw) {
                const int ww = col + kw - 1;
                if (ww < 0 || ww >= Wf || ww < dd) continue;
                for (int c = 0; c < C; ++c) {
                    const float wl = __ldg(w + ((((size_t)o * 2 * C + c) * 3 + kd) * 3 + kh) * 3 + kw);
                    const float wr = __ldg(w + ((((size_t)o * 2 * C + C + c) * 3 + kd) * 3 + kh) * 3 + kw);
                    acc = __fmaf_rn(wl, __ldg(xb + (size_t)c * img + (size_t)hh * Wf + ww), acc);
                    acc = __fmaf_rn(wr, __ldg(yb + (size_t)c * img + (size_t)hh * Wf + ww - dd), acc);
                }
            }
        }
    }
    const float sc = scale ? __ldg(scale + o) : 1.f, sh = shift ? __ldg(shift + o) : 0.f;
    float v = __fmaf_rn(acc, sc, sh);
    out[(size_t)b * n + idx] = relu ? fmaxf(v, 0.f) : v;
}

// Batch moments of the stem convolution: moments [B,Hf,O,2] fp64 = per (b,h,o) row (sum z, sum z^2) over Df x Wf.
int cv_stem_moments(const float* x, const float* y, const float* w, double* moments, int B, int C, int O, int Df, int Hf, int Wf,
                    float* workspace, cudaStream_t st) {
    if (!x || !y || !w || !moments) return fail(RAG_E_NULL, "cv_stem_moments: null pointer");
    if (B <= 0 || O <= 0 || Df <= 0 || Hf <= 0 || Wf <= 0 || B > 65535 || O > 65535)
        return fail(RAG_E_SHAPE, "cv_stem_moments: bad shape B=%d C=%d O=%d Df=%d Hf=%d Wf=%d", B, C, O, Df, Hf, Wf);
    const int Wv = Wf / 4;
    if (!(C == 12 && Df >= 3 && Wf >= 8 && Wf % 4 == 0 && Wv <= 512 && aligned(x, 8) && aligned(y, 8) && (!workspace || aligned(workspace, 16))))
        return fail(RAG_E_SHAPE, "cv_stem_moments: needs C == 12, Df >= 3, Wf >= 8, Wf %% 4 == 0, Wf <= 2048, 8-byte aligned features and a 16-byte aligned workspace");
    int k = 384 / Wv;
    if (k < 1) k = 1;
    int nt = k * Wv;
    if (nt < 128) nt = ((128 + Wv - 1) / Wv) * Wv;
    const size_t smem2 = ((size_t)2 * C * 36 + (size_t)33 * (Wf + 2 * kStemPad) + (size_t)4 * Df + ((Df + 3) & ~3)) * sizeof(float) +
                         (size_t)kStemStages * 3 * nt * sizeof(float2);
    if (smem2 > 200 * 1024) return fail(RAG_E_SHAPE, "cv_stem_moments: Wf=%d too wide for shared memory", Wf);
    auto kern = cv_stem_fwd2_kernel<12, false, true>;
    if (smem2 + 8192 > 48 * 1024) {                  // + the kernel's 8 KB of static shared memory (fp64 reduction buffers)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return fail((int)e, "cv_stem_moments: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    if (workspace) {
        stem_regroup_kernel<<<O, 256, 0, st>>>(w, C, workspace);
        if (int e = check_launch("cv_stem_moments(regroup)")) return e;
    }
    kern<<<dim3(Hf, O, B), nt, smem2, st>>>(x, y, w, nullptr, nullptr, nullptr, O, Df, Hf, Wf, workspace, moments);
    return check_launch("cv_stem_moments");
}

int cv_stem_fwd(const float* x, const float* y, const float* w, const float* scale, const float* shift, int relu,
                float* out, int B, int C, int O, int Df, int Hf, int Wf, float* workspace, int variant, cudaStream_t st) {
    if (!x || !y || !w || !out) return fail(RAG_E_NULL, "cv_stem_fwd: null pointer");
    if (workspace && !aligned(workspace, 16)) return fail(RAG_E_ALIGN, "cv_stem_fwd: workspace must be 16-byte aligned");
    if (B <= 0 || C <= 0 || O <= 0 || Df <= 0 || Hf <= 0 || Wf <= 0 || B > 65535 || O > 65535 || Hf > 2147483647 / 4)
        return fail(RAG_E_SHAPE, "cv_stem_fwd: bad shape B=%d C=%d O=%d Df=%d Hf=%d Wf=%d", B, C, O, Df, Hf, Wf);
    if ((size_t)O * Df * Hf * Wf >= ((size_t)1 << 40)) return fail(RAG_E_SHAPE, "cv_stem_fwd: output too large");
    if (variant < -1 || variant > 2) return fail(RAG_E_VARIANT, "cv_stem_fwd: unknown variant %d", variant);
    const size_t smem = ((size_t)2 * C * 36 + (size_t)33 * (Wf + 2 * kStemPad) + (size_t)5 * Df) * sizeof(float);
    const bool fast_ok = C == 12 && Df >= 3 && Wf >= 8 && smem <= 200 * 1024 && aligned(out, 16) && Hf <= 65535 * 32;
    if (variant >= 1 && !fast_ok) return fail(RAG_E_VARIANT, "cv_stem_fwd: variants 1/2 need C == 12, Df >= 3, Wf >= 8 and a 16-byte aligned output");
    const bool packed_ok = fast_ok && Wf % 4 == 0 && Wf / 4 <= 512 && aligned(x, 8) && aligned(y, 8);
    if (variant == 2 && !packed_ok) return fail(RAG_E_VARIANT, "cv_stem_fwd: variant 2 needs Wf %% 4 == 0, Wf <= 2048 and 8-byte aligned features");
    if (variant == -1) variant = packed_ok ? 2 : (fast_ok ? 1 : 0);
    if (variant == 2) {
        auto kern = relu ? cv_stem_fwd2_kernel<12, true> : cv_stem_fwd2_kernel<12, false>;
        // block = a multiple of Wf/4 (a thread keeps its column group in phase 3), as close to 384 threads as possible
        const int Wv = Wf / 4;
        int k = 384 / Wv;
        if (k < 1) k = 1;
        int nt = k * Wv;
        if (nt < 128) nt = ((128 + Wv - 1) / Wv) * Wv;
        const size_t smem2 = ((size_t)2 * C * 36 + (size_t)33 * (Wf + 2 * kStemPad) + (size_t)4 * Df + ((Df + 3) & ~3)) * sizeof(float) +
                             (size_t)kStemStages * 3 * nt * sizeof(float2);
        if (smem2 > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
            if (e != cudaSuccess) return fail((int)e, "cv_stem_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        }
        if (workspace) {
            stem_regroup_kernel<<<O, 256, 0, st>>>(w, C, workspace);
            if (int e = check_launch("cv_stem_fwd(regroup)")) return e;
        }
        kern<<<dim3(Hf, O, B), nt, smem2, st>>>(x, y, w, scale, shift, out, O, Df, Hf, Wf, workspace, nullptr);
    } else if (variant == 1) {
        auto kern = cv_stem_fwd_kernel<12>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return fail((int)e, "cv_stem_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        }
        int nt = ((Wf + 31) / 32) * 32;
        nt = nt > 512 ? 512 : (nt < 128 ? 128 : nt);
        kern<<<dim3(Hf, O, B), nt, smem, st>>>(x, y, w, scale, shift, relu, out, O, Df, Hf, Wf);
    } else {
        const size_t n = (size_t)O * Df * Hf * Wf;
        cv_stem_direct_kernel<<<dim3((unsigned)((n + 255) / 256), B), 256, 0, st>>>(x, y, w, scale, shift, relu, out, C, O, Df, Hf, Wf);
    }
    return check_launch("cv_stem_fwd");
}

// bytes of the optional regrouped-weight workspace of cv_stem_fwd / cv_stem_moments
size_t cv_stem_workspace_bytes(int C, int O) { return C == 12 && O > 0 ? (size_t)O * kStemWPerO * sizeof(float) : 0; }

}  // namespace rag
