// Cost-volume forward with the TMA copy engine doing the HBM stores (cp.async.bulk shared -> global).
//
// The volume is 2*Df shifted / masked copies of every feature row (rag_model.py:375-383):
//     cost[b,   c, d, h, w] = (w >= d) ? x[b,c,h,w]   : 0
//     cost[b, C+c, d, h, w] = (w >= d) ? y[b,c,h,w-d] : 0
// so a CTA stages R rows of one (b,c) in shared memory ONCE (each feature row is read from HBM once per
// full disparity sweep) and every output row then leaves as bulk copies issued by a single thread each:
//   * right half, d = 4q+s: one bulk copy of Wf floats from image s of the row.  Image s holds
//     [Df+s zeros | y row], i.e. the row pre-shifted by s elements behind a zero prefix, so the source of
//     out[0..Wf) = E[-d .. Wf-d) starts at element Df-4q of image s: 16-byte aligned on both sides;
//   * left half: bulk copy of 4q zeros, ONE ordinary 16-byte store for the vector that straddles the
//     mask edge, bulk copy of x[4(q+1) ..] from the plain row.
// The SM issues ~4 instructions per 1.25 KB output row instead of ~33 per 512 B (cv_fwd_kernel), which is
// what lets this kernel share SMs with the FP32-bound disparity head (overlap).
// Two stages: while the copy engine drains stage k, the CTA fills stage k^1 with the next rows.
//
// Requirements (checked by the host, otherwise cv_fwd_kernel runs): Wf % 4 == 0, Df % 4 == 0, Df <= Wf,
// 16-byte aligned x, y, cost.
#pragma once
#include "common.cuh"

namespace rag {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// grid: persistent, blockIdx.x strides over items (b, c, row tile).  128 threads.
// smem: zeros[Df] | 2 stages x ( X[R][Wf] | Y[R][4][Df+Wf+4] )

// Items = (b*C + c, row tile, d-chunk), handed out IN ORDER from an atomic counter (caller-owned workspace, zeroed by the launcher; see cv_lean.cuh) so that the resident CTAs
// write neighbouring rows of the same few disparity planes (HBM page locality).
template <int NT>
__global__ void __launch_bounds__(NT)
cv_fwd_tma_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ cost,
                  int B, int C, int Df, int Hf, int Wf, int R, int n_tiles, int dchunk, int n_dchunks, unsigned int* __restrict__ ctr) {
    extern __shared__ __align__(128) float cvt_smem[];
    __shared__ int s_item[2];
    const int YS = Df + Wf + 4;                    // one image of one row
    const int stage_floats = R * Wf + R * 4 * YS;
    float* zeros = cvt_smem;
    float* stage0 = cvt_smem + Df;
    const int tid = threadIdx.x;
    const size_t plane = (size_t)Hf * Wf;
    const int n_items = B * C * n_tiles * n_dchunks;
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

    // zero buffer + the zero prefixes of every image (never overwritten afterwards)
    for (int i = tid; i < Df; i += NT) zeros[i] = 0.f;
    for (int st = 0; st < 2; ++st) {
        float* Y = stage0 + st * stage_floats + R * Wf;
        for (int i = tid; i < R * 4 * (Df + 4); i += NT) {
            const int img = i / (Df + 4), k = i - img * (Df + 4);
            Y[img * YS + k] = 0.f;              // covers [0, Df+s) for every s <= 3
        }
    }
    if (tid == 0) s_item[0] = (int)atomicAdd(ctr, 1u);
    __syncthreads();

    int item = s_item[0];
    for (int it_local = 0; item < n_items; ++it_local) {
        if (tid == 0) s_item[(it_local + 1) & 1] = (int)atomicAdd(ctr, 1u);   // published by the barrier below
        const int it = item / n_dchunks, dc = item - it * n_dchunks;
        const int d_beg = dc * dchunk, nd = min(dchunk, Df - d_beg);
        const int tile = it % n_tiles;
        const int bc = it / n_tiles;               // b*C + c
        const int b = bc / C, c = bc - b * C;
        const int h0 = tile * R, rows = min(R, Hf - h0);
        float* X = stage0 + (it_local & 1) * stage_floats;
        float* Y = X + R * Wf;
        // the bulk copies issued from this stage two items ago must have finished READING shared memory
        bulk_wait_read<1>();
        __syncthreads();
        // ---- fill: x rows as they are, y rows as four images shifted by s = 0..3 behind the zero prefix ----
        const float4* xs = reinterpret_cast<const float4*>(x + (size_t)bc * plane + (size_t)h0 * Wf);
        const float4* ys = reinterpret_cast<const float4*>(y + (size_t)bc * plane + (size_t)h0 * Wf);
        const int nv = rows * (Wf >> 2);
        for (int i = tid; i < nv; i += NT) {
            reinterpret_cast<float4*>(X)[i] = __ldg(xs + i);
            const float4 v = __ldg(ys + i);
            const int r = i / (Wf >> 2), w = (i - r * (Wf >> 2)) << 2;
            float* yr = Y + r * 4 * YS + Df + w;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                float* p = yr + s * YS + s;
                p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
            }
        }
        fence_async_smem();                          // generic-proxy writes -> visible to the async proxy
        __syncthreads();
        // ---- drain: one (row, disparity) per thread per pass ----
        float* outL = cost + ((size_t)(b * 2 * C + c) * Df) * plane + (size_t)h0 * Wf;
        float* outR = cost + ((size_t)(b * 2 * C + C + c) * Df) * plane + (size_t)h0 * Wf;
        for (int p = tid; p < rows * nd; p += NT) {
            const int dl = p / rows, r = p - dl * rows;   // r fastest: neighbouring copies land in the same plane
            const int d = d_beg + dl;
            const int q = d >> 2, s = d & 3;
            const size_t o = (size_t)d * plane + (size_t)r * Wf;
            bulk_s2g(outR + o, Y + (r * 4 + s) * YS + (Df - 4 * q), (uint32_t)Wf * 4u, policy);
            const float* xr = X + r * Wf;
            if (q > 0) bulk_s2g(outL + o, zeros, (uint32_t)q * 16u, policy);
            float4 m = *reinterpret_cast<const float4*>(xr + 4 * q);
            if (s > 0) m.x = 0.f;
            if (s > 1) m.y = 0.f;
            if (s > 2) m.z = 0.f;
            __stcs(reinterpret_cast<float4*>(outL + o + 4 * q), m);
            const int tail = Wf - 4 * (q + 1);
            if (tail > 0) bulk_s2g(outL + o + 4 * (q + 1), xr + 4 * (q + 1), (uint32_t)tail * 4u, policy);
        }
        bulk_commit();
        item = s_item[(it_local + 1) & 1];
    }
    bulk_wait_read<0>();                             // shared memory must outlive the copies that read it
}

}  // namespace rag
