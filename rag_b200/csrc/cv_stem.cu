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
#include "common.cuh"

namespace rag {

constexpr int kStemPad = 4;   // zero columns on each side of every shared-memory row

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
    for (int i = tid; i < 33 * Wp; i += NT) maps[i] = 0.f;   // maps, LF, RF incl. their zero pads
    __syncthreads();

    // ---- phase 1: the 18 row maps, one column per thread ----
    const size_t img = (size_t)Hf * Wf;
    const float* xb = x + (size_t)b * C * img;
    const float* yb = y + (size_t)b * C * img;
    for (int col = tid; col < Wf; col += NT) {
        float a[9], r[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { a[i] = 0.f; r[i] = 0.f; }
        for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hh = h + kh - 1;
                if (hh < 0 || hh >= Hf) continue;
                const float xv = __ldg(xb + (size_t)c * img + (size_t)hh * Wf + col);
                const float yv = __ldg(yb + (size_t)c * img + (size_t)hh * Wf + col);
                const float4* wl = reinterpret_cast<const float4*>(wsm + (c * 3 + kh) * 12);
                const float4* wr = reinterpret_cast<const float4*>(wsm + ((C + c) * 3 + kh) * 12);
                const float4 l0 = wl[0], l1 = wl[1], l2 = wl[2], r0 = wr[0], r1 = wr[1], r2 = wr[2];
                a[0] = __fmaf_rn(l0.x, xv, a[0]); a[1] = __fmaf_rn(l0.y, xv, a[1]); a[2] = __fmaf_rn(l0.z, xv, a[2]);
                a[3] = __fmaf_rn(l0.w, xv, a[3]); a[4] = __fmaf_rn(l1.x, xv, a[4]); a[5] = __fmaf_rn(l1.y, xv, a[5]);
                a[6] = __fmaf_rn(l1.z, xv, a[6]); a[7] = __fmaf_rn(l1.w, xv, a[7]); a[8] = __fmaf_rn(l2.x, xv, a[8]);
                r[0] = __fmaf_rn(r0.x, yv, r[0]); r[1] = __fmaf_rn(r0.y, yv, r[1]); r[2] = __fmaf_rn(r0.z, yv, r[2]);
                r[3] = __fmaf_rn(r0.w, yv, r[3]); r[4] = __fmaf_rn(r1.x, yv, r[4]); r[5] = __fmaf_rn(r1.y, yv, r[5]);
                r[6] = __fmaf_rn(r1.z, yv, r[6]); r[7] = __fmaf_rn(r1.w, yv, r[7]); r[8] = __fmaf_rn(r2.x, yv, r[8]);
            }
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            maps[i * Wp + kStemPad + col] = a[i];
            maps[(9 + i) * Wp + kStemPad + col] = r[i];
        }
    }
    __syncthreads();

    // ---- phase 2: collapsed rows per disparity class (0: d=0 -> kd in {1,2}; 1: interior; 2: d=Df-1 -> kd in {0,1}) ----
    for (int i = tid; i < 3 * Wf; i += NT) {
        const int cls = i / Wf, col = i - cls * Wf;
        const int kd0 = cls == 0 ? 1 : 0, kd1 = cls == 2 ? 1 : 2;
        float lf = 0.f, rf = 0.f;
        for (int kd = kd0; kd <= kd1; ++kd)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                lf += maps[(kd * 3 + kw) * Wp + kStemPad + col + kw - 1];
                rf += maps[(9 + kd * 3 + kw) * Wp + kStemPad + col + kw - kd];
            }
        LF[cls * Wp + kStemPad + col] = lf;
#pragma unroll
        for (int s = 0; s < 4; ++s)             // copy s holds RF shifted right by s: copy_s[i] = RF[i - s]
            if (col + s < Wf) RF[(cls * 4 + s) * Wp + kStemPad + col + s] = rf;
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
        for (int i = tid; i < Df * Wv; i += NT) {
            const int d = i / Wv, w0 = (i - d * Wv) << 2;
            const int cls = d == 0 ? 0 : (d == Df - 1 ? 2 : 1);
            const int u0 = w0 - d;
            float4 v;
            if (u0 >= 2 && w0 + 3 <= Wf - 2) {           // all four interior: aligned reads of LF and of RF copy (d & 3)
                const int q = d >> 2, s = d & 3;
                const float4 lf = *reinterpret_cast<const float4*>(LF + cls * Wp + kStemPad + w0);
                const float4 rf = *reinterpret_cast<const float4*>(RF + (cls * 4 + s) * Wp + kStemPad + w0 - 4 * q);
                v = make_float4(lf.x + rf.x, lf.y + rf.y, lf.z + rf.z, lf.w + rf.w);
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
            st_stream(reinterpret_cast<float4*>(ob + (size_t)d * dstride + w0), v);
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

int cv_stem_fwd(const float* x, const float* y, const float* w, const float* scale, const float* shift, int relu,
                float* out, int B, int C, int O, int Df, int Hf, int Wf, int variant, cudaStream_t st) {
    if (!x || !y || !w || !out) return fail(RAG_E_NULL, "cv_stem_fwd: null pointer");
    if (B <= 0 || C <= 0 || O <= 0 || Df <= 0 || Hf <= 0 || Wf <= 0 || B > 65535 || O > 65535 || Hf > 2147483647 / 4)
        return fail(RAG_E_SHAPE, "cv_stem_fwd: bad shape B=%d C=%d O=%d Df=%d Hf=%d Wf=%d", B, C, O, Df, Hf, Wf);
    if ((size_t)O * Df * Hf * Wf >= ((size_t)1 << 40)) return fail(RAG_E_SHAPE, "cv_stem_fwd: output too large");
    if (variant < -1 || variant > 1) return fail(RAG_E_VARIANT, "cv_stem_fwd: unknown variant %d", variant);
    const size_t smem = ((size_t)2 * C * 36 + (size_t)33 * (Wf + 2 * kStemPad) + (size_t)5 * Df) * sizeof(float);
    const bool fast_ok = C == 12 && Df >= 3 && Wf >= 8 && smem <= 200 * 1024 && aligned(out, 16) && Hf <= 65535 * 32;
    if (variant == 1 && !fast_ok) return fail(RAG_E_VARIANT, "cv_stem_fwd: variant 1 needs C == 12, Df >= 3, Wf >= 8 and a 16-byte aligned output");
    if (variant == -1) variant = fast_ok ? 1 : 0;
    if (variant == 1) {
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

}  // namespace rag
