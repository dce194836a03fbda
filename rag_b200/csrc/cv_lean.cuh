// Cost-volume forward, lean thread-stationary kernel (default when Wf % 4 == 0).
//
// cv_fwd_kernel reaches the HBM store roofline but spends ~33 warp instructions per 512-byte warp store
// (items dealt to threads, table look-ups, per-item mask logic: 165 M warp instructions at B=8 480x960,
// 37 % of the issue slots).  Here a thread OWNS output vector positions (row r, vector v) of a tile of R
// feature rows and walks the Df disparities:
//   * right half  out[d][r][4v..4v+3] = E_r[4v-d .. 4v-d+3]  (E = y row behind a zero prefix): one aligned
//     LDS.128 from image (d & 3) of the row -- the row pre-shifted by 0..3 elements -- and one STG.128;
//   * left half   out[d][r][4v..4v+3] = x masked by w >= d: the x vector sits in registers; while
//     d <= min over the warp of 4v the store is unmasked (warp-uniform loop bound), afterwards four
//     compare/select pairs build the mask.
// ~4 instructions per warp store (21 M in total), so the kernel needs ~5 % of the issue slots and few
// resident warps; as a persistent grid with a fixed number of CTAs per SM it leaves the SM to the
// FP32-bound disparity head running on a second stream (rag_b200.pipeline).
// Each feature row is read from HBM once per full disparity sweep; the next tile's rows are prefetched
// into registers while the current tile streams out.
#pragma once
#include "common.cuh"

namespace rag {

// grid: persistent, blockIdx.x strides over items (b*C + c, row tile).  NT threads, VPT vectors per thread.
// smem: Y[R][4][YS], YS = Df4 + Wf + 4 with Df4 = Df rounded up to a multiple of 4.
// Work counter of the persistent form (DYN): ONE unsigned int "next item" in a CALLER-OWNED workspace, zeroed by the
// launcher with cudaMemsetAsync on the launch stream right before the kernel -- the library keeps no state of its
// own, so concurrent launches, graph capture and replays cannot interfere (each launch has its own counter).

// DYN: the CTAs of a persistent grid take items from an atomic counter IN ORDER, so the resident CTAs always
// work on neighbouring rows / disparity chunks (a static stride lets them drift apart and costs 12 % of
// the HBM write bandwidth); !DYN: item = blockIdx.x + k * gridDim.x (one item per CTA when the grid covers
// all items).
// ST: store flavour of the output stream (0 = st.global.cs, 1 = plain st.global, 2 = st.global.wt)
template <int ST>
__device__ __forceinline__ void cv_store(float4* p, float4 v) {
    if (ST == 0) __stcs(p, v);
    else if (ST == 1) *p = v;
    else __stwt(p, v);
}

template <int NT, int VPT, bool DYN, int ST = 0>
__global__ void __launch_bounds__(NT)
cv_fwd_lean_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ cost,
                   int BC, int C, int Df, int Hf, int Wf, int R, int n_tiles, int dchunk, int n_dchunks, unsigned int* __restrict__ ctr) {
    extern __shared__ __align__(16) float cvl_smem[];
    __shared__ int s_item[2];
    const int Df4 = (Df + 3) & ~3;
    const int YS = Df4 + Wf + 4;
    const int Wv = Wf >> 2;
    const int tid = threadIdx.x;
    const size_t plane = (size_t)Hf * Wf;
    const int n_items = BC * n_tiles * n_dchunks;    // item = ((b*C + c) * n_tiles + tile) * n_dchunks + d-chunk

    // zero prefixes [0, Df4 + s) of every image: written once, never overwritten
    for (int i = tid; i < R * 4 * (Df4 + 4); i += NT) {
        const int img = i / (Df4 + 4), k = i - img * (Df4 + 4);
        cvl_smem[img * YS + k] = 0.f;
    }

    int rv[VPT], vv[VPT];                           // row / vector inside the tile of the owned positions
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int p = tid + i * NT;
        rv[i] = p / Wv;
        vv[i] = p - rv[i] * Wv;
    }
    float4 xn[VPT], yn[VPT];
    auto prefetch = [&](int item) {
        const int it = item / n_dchunks;
        const int tile = it % n_tiles, bc = it / n_tiles;
        const int h0 = tile * R, rows = min(R, Hf - h0);
        const size_t off = (size_t)bc * plane + (size_t)h0 * Wf;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            if (rv[i] < rows) {
                xn[i] = __ldg(reinterpret_cast<const float4*>(x + off) + tid + i * NT);
                yn[i] = __ldg(reinterpret_cast<const float4*>(y + off) + tid + i * NT);
            }
        }
    };
    int item = blockIdx.x;
    if (DYN) {
        if (tid == 0) s_item[0] = (int)atomicAdd(ctr, 1u);
        __syncthreads();
        item = s_item[0];
    }
    if (item < n_items) prefetch(item);

    for (int k = 0; item < n_items; ++k) {
        if (DYN && tid == 0) s_item[(k + 1) & 1] = (int)atomicAdd(ctr, 1u);   // published by the barrier below
        const int it = item / n_dchunks, dc = item - it * n_dchunks;
        const int tile = it % n_tiles, bc = it / n_tiles;
        const int d_beg = dc * dchunk, d_end = min(d_beg + dchunk, Df);   // dchunk % 4 == 0
        const int b = bc / C, c = bc - b * C;
        const int h0 = tile * R, rows = min(R, Hf - h0);
        float4 xv[VPT];
        __syncthreads();                             // the previous tile's LDS are done (and the zero prefixes written)
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            xv[i] = xn[i];
            if (rv[i] < rows) {
                float* yr = cvl_smem + rv[i] * 4 * YS + Df4 + 4 * vv[i];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    float* p = yr + s * YS + s;      // image s = the row shifted right by s
                    p[0] = yn[i].x; p[1] = yn[i].y; p[2] = yn[i].z; p[3] = yn[i].w;
                }
            }
        }
        const int next = DYN ? s_item[(k + 1) & 1] : item + (int)gridDim.x;
        if (next < n_items) prefetch(next);
        __syncthreads();

        float* outL = cost + ((size_t)(b * 2 * C + c) * Df) * plane + (size_t)h0 * Wf;
        float* outR = outL + (size_t)C * Df * plane;
        // d is the OUTER loop: at any moment the CTA's warps write neighbouring vectors of the same few
        // disparity planes (R*Wf*4 contiguous bytes per plane), which is what keeps HBM pages open
        bool on[VPT];
        int t0[VPT], dA[VPT];
        const float* src[VPT];
        float* pL[VPT];
        float* pR[VPT];
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            on[i] = rv[i] < rows;
            const int e = rv[i] * Wf + 4 * vv[i];    // element offset of the owned vector inside a plane tile
            src[i] = cvl_smem + rv[i] * 4 * YS + Df4 + 4 * vv[i];
            pL[i] = outL + e + (size_t)d_beg * plane;
            pR[i] = outR + e + (size_t)d_beg * plane;
            t0[i] = on[i] ? 4 * vv[i] : 0x7fffffff;
            dA[i] = __reduce_min_sync(0xffffffffu, t0[i]);   // warp-uniform: unmasked left stores while d <= dA
        }
        auto emit = [&](int d4, int s) {             // disparity d4 + s, d4 % 4 == 0
            const int d = d4 + s;
#pragma unroll
            for (int i = 0; i < VPT; ++i) {
                if (on[i]) {
                    cv_store<ST>(reinterpret_cast<float4*>(pR[i]), *reinterpret_cast<const float4*>(src[i] + s * YS - d4));
                    float4 m = xv[i];
                    if (d > dA[i]) {
                        m.x = t0[i] + 0 >= d ? m.x : 0.f;
                        m.y = t0[i] + 1 >= d ? m.y : 0.f;
                        m.z = t0[i] + 2 >= d ? m.z : 0.f;
                        m.w = t0[i] + 3 >= d ? m.w : 0.f;
                    }
                    cv_store<ST>(reinterpret_cast<float4*>(pL[i]), m);
                }
                pL[i] += plane;
                pR[i] += plane;
            }
        };
        int d4 = d_beg;
        for (; d4 + 3 < d_end; d4 += 4) {
#pragma unroll
            for (int s = 0; s < 4; ++s) emit(d4, s);
        }
        for (int s = 0; d4 + s < d_end; ++s) emit(d4, s);
        item = next;
    }
}

}  // namespace rag
