// Trilinear resize of a [B*C, Di, Hi, Wi] volume, forward and deterministic backward:
// the `upsample_6` / `upsample_12` branches of the Matching Net's tail, src/models/rag_model.py:356-366 and :675-685:
//     upsample_6  = nn.Upsample(size=x.size()[2:],        mode='trilinear', align_corners=True)
//     upsample_12 = nn.Upsample(size=[d//2, h//2, w//2],  mode='trilinear', align_corners=True)
// applied between last_12_3d / last_6_3d / last_3_3d, i.e. they produce the input of the layer that feeds the head.
//
// Arithmetic = PyTorch's (ATen/native/cuda/UpSample.cuh:96-130 + UpSampleTrilinear3d.cu), all fp32:
//     align_corners:  scale = (in-1)/(out-1) (0 if out == 1),  src = scale*dst
//     otherwise:      scale = in/out,                         src = max(scale*(dst+0.5)-0.5, 0)
//     i0 = (int)src, i1 = i0 + (i0 < in-1), l1 = src - i0, l0 = 1 - l1
//     out = t0*(h0*(w0*a000 + w1*a001) + h1*(w0*a010 + w1*a011)) + t1*(h0*(w0*a100 + w1*a101) + h1*(w0*a110 + w1*a111))
// Forward: one thread per 4 consecutive output columns (128-bit streaming stores; the input is 8x smaller and stays in
// L1/L2).  Backward: PyTorch scatters with atomicAdd (arrival order: not reproducible); here every INPUT voxel gathers the
// output voxels that reference it, in a fixed order -- the per-axis (i0, l1) tables and the first-destination index of
// every source index are built in shared memory by each CTA -- so the result is bitwise repeatable.
#include "common.cuh"

namespace rag {

template <bool AC>
__device__ __forceinline__ void tri_src(float scale, int dst, int n_in, int& i0, int& i1, float& l0, float& l1) {
    if (AC) {
        const float s = scale * (float)dst;
        i0 = min((int)s, n_in - 1);
        i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
        l1 = s - (float)i0;
        l0 = 1.f - l1;
    } else {
        src_index<true>(scale, dst, n_in, i0, i1, l0, l1);
    }
}

__host__ __device__ inline float tri_scale(int n_in, int n_out, bool ac) {
    if (ac) return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f;
    return (float)n_in / (float)n_out;
}

// grid-stride over (bc, d, h, w-quad); Wo % 4 == 0 not required (scalar tail).
template <bool AC>
__global__ void __launch_bounds__(256)
trilinear_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int BC, int Di, int Hi, int Wi, int Do, int Ho, int Wo,
                     float sd, float sh, float sw, bool vec) {
    const int Wq = (Wo + 3) >> 2;
    const size_t n = (size_t)BC * Do * Ho * Wq;
    const size_t in_plane = (size_t)Hi * Wi, in_vol = (size_t)Di * in_plane;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const int q = (int)(i % Wq);
        size_t r = i / Wq;
        const int h = (int)(r % Ho); r /= Ho;
        const int d = (int)(r % Do);
        const size_t bc = r / Do;
        int d0, d1, h0, h1;
        float t0, t1, e0, e1;
        tri_src<AC>(sd, d, Di, d0, d1, t0, t1);
        tri_src<AC>(sh, h, Hi, h0, h1, e0, e1);
        const float* p = in + bc * in_vol;
        const float* r00 = p + (size_t)d0 * in_plane + (size_t)h0 * Wi;
        const float* r01 = p + (size_t)d0 * in_plane + (size_t)h1 * Wi;
        const float* r10 = p + (size_t)d1 * in_plane + (size_t)h0 * Wi;
        const float* r11 = p + (size_t)d1 * in_plane + (size_t)h1 * Wi;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int w = 4 * q + k;
            int w0, w1;
            float f0, f1;
            tri_src<AC>(sw, min(w, Wo - 1), Wi, w0, w1, f0, f1);
            v[k] = t0 * (e0 * (f0 * __ldg(r00 + w0) + f1 * __ldg(r00 + w1)) + e1 * (f0 * __ldg(r01 + w0) + f1 * __ldg(r01 + w1))) +
                   t1 * (e0 * (f0 * __ldg(r10 + w0) + f1 * __ldg(r10 + w1)) + e1 * (f0 * __ldg(r11 + w0) + f1 * __ldg(r11 + w1)));
        }
        float* o = out + ((bc * Do + d) * Ho + h) * (size_t)Wo + 4 * q;
        if (vec) {
            st_stream(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (4 * q + k < Wo) o[k] = v[k];
        }
    }
}

// Per-axis tables in shared memory: i0[dst], l1[dst] for dst < n_out and first[i] = smallest dst with i0[dst] >= i for i <= n_in.
template <bool AC>
__device__ void tri_tables(float scale, int n_in, int n_out, int* i0t, float* l1t, int* first) {
    for (int dst = threadIdx.x; dst < n_out; dst += blockDim.x) {
        int a, b;
        float l0, l1;
        tri_src<AC>(scale, dst, n_in, a, b, l0, l1);
        i0t[dst] = a;
        l1t[dst] = l1;
    }
    __syncthreads();
    for (int dst = threadIdx.x; dst <= n_out; dst += blockDim.x) {
        const int prev = dst == 0 ? -1 : i0t[dst - 1];
        const int cur = dst == n_out ? n_in : i0t[dst];      // sentinel: every remaining source index starts past the end
        for (int i = prev + 1; i <= cur; ++i) first[i] = dst;
    }
    __syncthreads();
}

// gin[bc, id, ih, iw] = sum over the output voxels that read it of (weight * gout).  One thread per input voxel.
// smem: per axis  int i0[n_out] | float l1[n_out] | int first[n_in + 1]
template <bool AC>
__global__ void __launch_bounds__(256)
trilinear_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int BC, int Di, int Hi, int Wi, int Do, int Ho, int Wo,
                     float sd, float sh, float sw) {
    extern __shared__ int tri_smem[];
    int* d_i0 = tri_smem;            float* d_l1 = reinterpret_cast<float*>(d_i0 + Do); int* d_first = reinterpret_cast<int*>(d_l1 + Do);
    int* h_i0 = d_first + Di + 1;    float* h_l1 = reinterpret_cast<float*>(h_i0 + Ho); int* h_first = reinterpret_cast<int*>(h_l1 + Ho);
    int* w_i0 = h_first + Hi + 1;    float* w_l1 = reinterpret_cast<float*>(w_i0 + Wo); int* w_first = reinterpret_cast<int*>(w_l1 + Wo);
    tri_tables<AC>(sd, Di, Do, d_i0, d_l1, d_first);
    tri_tables<AC>(sh, Hi, Ho, h_i0, h_l1, h_first);
    tri_tables<AC>(sw, Wi, Wo, w_i0, w_l1, w_first);
    const size_t n = (size_t)BC * Di * Hi * Wi;
    const size_t out_plane = (size_t)Ho * Wo, out_vol = (size_t)Do * out_plane;
    // weight of destination `dst` on source index i along an axis with tables (i0, l1), n_in sources
    auto wgt = [](int i, int a, float l1, int n_in) -> float {
        const int b = a + (a < n_in - 1 ? 1 : 0);
        return (a == i ? 1.f - l1 : 0.f) + (b == i ? l1 : 0.f);
    };
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < n; idx += (size_t)gridDim.x * 256) {
        const int iw = (int)(idx % Wi);
        size_t r = idx / Wi;
        const int ih = (int)(r % Hi); r /= Hi;
        const int id = (int)(r % Di);
        const size_t bc = r / Di;
        const int dl = d_first[max(id - 1, 0)], dh = d_first[id + 1];      // destinations with i0 in {id-1, id}
        const int hl = h_first[max(ih - 1, 0)], hh = h_first[ih + 1];
        const int wl = w_first[max(iw - 1, 0)], wh = w_first[iw + 1];
        const float* gp = gout + bc * out_vol;
        float acc = 0.f;
        if (wh - wl <= 5) {
            // common case (up-sampling by ~2-3): at most five destinations per axis.  The w weights live in registers and
            // the five loads per row are unconditional (clamped column, zero weight) -- no table look-ups in the inner loop.
            float ww[5];
            int wc[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int w = wl + k;
                wc[k] = min(w, Wo - 1);
                ww[k] = w < wh ? wgt(iw, w_i0[wc[k]], w_l1[wc[k]], Wi) : 0.f;
            }
            for (int d = dl; d < dh; ++d) {
                const float wd = wgt(id, d_i0[d], d_l1[d], Di);
                float accd = 0.f;
                for (int h = hl; h < hh; ++h) {
                    const float whh = wgt(ih, h_i0[h], h_l1[h], Hi);
                    const float* row = gp + (size_t)d * out_plane + (size_t)h * Wo;
                    float acch = ww[0] * __ldg(row + wc[0]);
#pragma unroll
                    for (int k = 1; k < 5; ++k) acch = __fmaf_rn(ww[k], __ldg(row + wc[k]), acch);
                    accd = __fmaf_rn(whh, acch, accd);
                }
                acc = __fmaf_rn(wd, accd, acc);
            }
        } else {
            for (int d = dl; d < dh; ++d) {
                const float wd = wgt(id, d_i0[d], d_l1[d], Di);
                if (wd == 0.f) continue;
                float accd = 0.f;
                for (int h = hl; h < hh; ++h) {
                    const float whh = wgt(ih, h_i0[h], h_l1[h], Hi);
                    if (whh == 0.f) continue;
                    const float* row = gp + (size_t)d * out_plane + (size_t)h * Wo;
                    float acch = 0.f;
                    for (int w = wl; w < wh; ++w) acch = __fmaf_rn(wgt(iw, w_i0[w], w_l1[w], Wi), __ldg(row + w), acch);
                    accd = __fmaf_rn(whh, acch, accd);
                }
                acc = __fmaf_rn(wd, accd, acc);
            }
        }
        gin[idx] = acc;
    }
}

static int check_tri(const char* who, const void* a, const void* b, int BC, int Di, int Hi, int Wi, int Do, int Ho, int Wo) {
    if (!a || !b) return fail(RAG_E_NULL, "%s: null pointer", who);
    if (BC <= 0 || Di <= 0 || Hi <= 0 || Wi <= 0 || Do <= 0 || Ho <= 0 || Wo <= 0) return fail(RAG_E_SHAPE, "%s: non-positive dimension", who);
    if ((size_t)Di * Hi * Wi >= ((size_t)1 << 31) || (size_t)Do * Ho * Wo >= ((size_t)1 << 31))
        return fail(RAG_E_SHAPE, "%s: a volume must have fewer than 2^31 voxels", who);
    if ((size_t)(Do + Ho + Wo) * 8 + (size_t)(Di + Hi + Wi + 3) * 4 > 160 * 1024) return fail(RAG_E_SHAPE, "%s: axes too long for the shared-memory tables", who);
    if (!aligned(a, 4) || !aligned(b, 4)) return fail(RAG_E_ALIGN, "%s: pointers must be 4-byte aligned", who);
    return RAG_OK;
}

int trilinear_resize_fwd(const float* in, float* out, int BC, int Di, int Hi, int Wi, int Do, int Ho, int Wo, int align_corners, cudaStream_t st) {
    if (int e = check_tri("trilinear_resize_fwd", in, out, BC, Di, Hi, Wi, Do, Ho, Wo)) return e;
    const bool ac = align_corners != 0;
    const float sd = tri_scale(Di, Do, ac), sh = tri_scale(Hi, Ho, ac), sw = tri_scale(Wi, Wo, ac);
    const size_t n = (size_t)BC * Do * Ho * ((Wo + 3) / 4);
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)num_sms() * 32);
    const bool vec = Wo % 4 == 0 && aligned(out, 16);
    if (ac) trilinear_fwd_kernel<true><<<grid, 256, 0, st>>>(in, out, BC, Di, Hi, Wi, Do, Ho, Wo, sd, sh, sw, vec);
    else trilinear_fwd_kernel<false><<<grid, 256, 0, st>>>(in, out, BC, Di, Hi, Wi, Do, Ho, Wo, sd, sh, sw, vec);
    return check_launch("trilinear_resize_fwd");
}

int trilinear_resize_bwd(const float* gout, float* gin, int BC, int Di, int Hi, int Wi, int Do, int Ho, int Wo, int align_corners, cudaStream_t st) {
    if (int e = check_tri("trilinear_resize_bwd", gout, gin, BC, Di, Hi, Wi, Do, Ho, Wo)) return e;
    const bool ac = align_corners != 0;
    const float sd = tri_scale(Di, Do, ac), sh = tri_scale(Hi, Ho, ac), sw = tri_scale(Wi, Wo, ac);
    const size_t n = (size_t)BC * Di * Hi * Wi;
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)num_sms() * 16);
    const size_t smem = (size_t)(Do + Ho + Wo) * 8 + (size_t)(Di + Hi + Wi + 3) * 4;
    auto kern = ac ? trilinear_bwd_kernel<true> : trilinear_bwd_kernel<false>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail((int)e, "trilinear_resize_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    kern<<<grid, 256, smem, st>>>(gout, gin, BC, Di, Hi, Wi, Do, Ho, Wo, sd, sh, sw);
    return check_launch("trilinear_resize_bwd");
}

}  // namespace rag
