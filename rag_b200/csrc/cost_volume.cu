// Concatenation cost volume, forward and backward, for sm_100a.
//
// Reference semantics: src/models/rag_model.py:375-383 (and its two verbatim copies):
//   cost[b,   c, d,h,w] = x[b,c,h,w]   (w >= d) else 0
//   cost[b, C+c, d,h,w] = y[b,c,h,w-d] (w >= d) else 0
// The reference does this with one zero-fill + 2*Df strided slice copies (129 launches, the
// volume written twice) and its autograd replays 2*Df CopySlices nodes that each clone the full
// gradient volume.  Here: one launch per direction, every byte of the volume touched once.
//
// Forward  (HBM store stream): a CTA stages R feature rows of one (b,c) -- the left row once and
//   the right row as V copies pre-shifted by 0..V-1 elements -- in shared memory, so that for every
//   disparity d = V*q + s the shifted right row is an ALIGNED 16-byte shared-memory read of copy s
//   at vector offset -q, and every global store is a full, aligned, coalesced 128-bit store.
//   Each input row is read from HBM/L2 once per disparity sweep (d-chunk), not once per disparity.
// Backward (HBM load stream): one thread owns V adjacent w of one (b,c,h) row and walks d from
//   Df-1 down to 0 accumulating in fp32 -- exactly the order in which autograd accumulates the
//   CopySlices gradients -- so the result is bit-identical to the reference; loads are 128-bit,
//   coalesced, and issued U disparities ahead of the dependent adds.  No atomics, no reduction tree.
#include <algorithm>

#include "common.cuh"
#include "cv_tma.cuh"
#include "cv_lean.cuh"

namespace rag {

template <int V> struct Vec;
template <> struct Vec<4> { using T = float4; };
template <> struct Vec<2> { using T = float2; };
template <> struct Vec<1> { using T = float; };

__device__ __forceinline__ float get(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ float get(const float2& v, int i) { return i == 0 ? v.x : v.y; }
__device__ __forceinline__ float get(const float& v, int) { return v; }
__device__ __forceinline__ void set(float4& v, int i, float a) { if (i == 0) v.x = a; else if (i == 1) v.y = a; else if (i == 2) v.z = a; else v.w = a; }
__device__ __forceinline__ void set(float2& v, int i, float a) { if (i == 0) v.x = a; else v.y = a; }
__device__ __forceinline__ void set(float& v, int, float a) { v = a; }
template <typename T> __device__ __forceinline__ T zero_vec();
template <> __device__ __forceinline__ float4 zero_vec<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <> __device__ __forceinline__ float2 zero_vec<float2>() { return make_float2(0.f, 0.f); }
template <> __device__ __forceinline__ float zero_vec<float>() { return 0.f; }

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// One work unit = R rows of one (b,c) x one sweep of `dchunk` disparities.
// smem: [V+1][R*Wf] floats + u16 table[R*Wv]
template <int V, int NT>
__device__ __forceinline__ void
cv_fwd_unit(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ cost,
            int C, int Df, int Hf, int Wf, int R, int dchunk, int tile, int dci, int c, int b, float* smem) {
    using VT = typename Vec<V>::T;
    const int Wv = Wf / V;
    const int h0 = tile * R;
    const int rows = min(R, Hf - h0);
    const int d0 = dci * dchunk;                 // multiple of V by construction
    const int nd = min(dchunk, Df - d0);
    const int P = rows * Wv;                     // vectors in the tile
    const int tile_elems = R * Wf;

    float* sx = smem;                            // left rows
    float* sy = smem + tile_elems;               // V shifted copies of the right rows
    unsigned short* wvtab = reinterpret_cast<unsigned short*>(smem + (V + 1) * tile_elems);

    // ---- stage the rows (each input element read once per CTA) ----
    const size_t in_off = ((size_t)(b * C + c) * Hf + h0) * Wf;
    const VT* xg = reinterpret_cast<const VT*>(x + in_off);
    const VT* yg = reinterpret_cast<const VT*>(y + in_off);
    for (int p = threadIdx.x; p < P; p += NT) {
        const int r = p / Wv;
        const int wv = p - r * Wv;
        wvtab[p] = (unsigned short)wv;
        reinterpret_cast<VT*>(sx)[p] = __ldg(xg + p);
        const VT yv = __ldg(yg + p);
        const int w0 = wv * V;
#pragma unroll
        for (int s = 0; s < V; ++s) {
            float* row = sy + s * tile_elems + r * Wf;
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const int i = w0 + k + s;        // copy_s[i] = y[i - s]
                if (i < Wf) row[i] = get(yv, k);
            }
            if (wv == 0) {
#pragma unroll
                for (int k = 0; k < V; ++k)
                    if (k < s) row[k] = 0.f;     // copy_s[i < s] = 0
            }
        }
    }
    __syncthreads();

    // ---- stream the volume: items = (d_local, p), consecutive threads -> consecutive vectors ----
    const size_t plane = (size_t)Hf * Wf;
    float* left = cost + (((size_t)(b * 2 * C + c) * Df + d0) * Hf + h0) * Wf;
    float* right = left + (size_t)C * Df * plane;
    const VT* sxv = reinterpret_cast<const VT*>(sx);

    int p = threadIdx.x, dl = 0;
    while (p >= P) { p -= P; ++dl; }
    while (dl < nd) {
        const int d = d0 + dl;
        const int wv = wvtab[p];
        const int w0 = wv * V;
        const size_t off = (size_t)dl * plane + (size_t)p * V;
        // left half: x masked by w >= d
        VT xv = sxv[p];
        if (w0 < d) {
#pragma unroll
            for (int k = 0; k < V; ++k)
                if (w0 + k < d) set(xv, k, 0.f);
        }
        st_stream(reinterpret_cast<VT*>(left + off), xv);
        // right half: y shifted by d = V*q + s  ->  copy s, vector (wv - q)
        const int q = d / V, s = d - q * V;
        VT yv = zero_vec<VT>();
        if (wv >= q) yv = reinterpret_cast<const VT*>(sy + s * tile_elems)[p - q];
        st_stream(reinterpret_cast<VT*>(right + off), yv);
        p += NT;
        while (p >= P) { p -= P; ++dl; }
    }
}

// grid: x = h-tile * n_dchunks + d-chunk, y = c, z = b.
template <int V, int NT>
__global__ void __launch_bounds__(NT)
cv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ cost,
              int C, int Df, int Hf, int Wf, int R, int dchunk, int n_dchunks) {
    extern __shared__ __align__(16) float smem[];
    const int tile = blockIdx.x / n_dchunks;
    cv_fwd_unit<V, NT>(x, y, cost, C, Df, Hf, Wf, R, dchunk, tile, blockIdx.x - tile * n_dchunks, blockIdx.y, blockIdx.z, smem);
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// Vector path (Wf % 4 == 0): one thread owns vector p of the (b, c) plane.
template <int U>
__device__ __forceinline__ void cv_bwd_v4_unit(const float* __restrict__ g, float* __restrict__ gx, float* __restrict__ gy,
                                               int C, int Df, int Wv, int PV, int p, int c, int b) {
    const int wv = p % Wv;
    const int w0 = wv << 2;
    const size_t planeV = (size_t)PV;
    const float4* gl = reinterpret_cast<const float4*>(g) + (size_t)(b * 2 * C + c) * Df * planeV + p;
    const float4* gr = gl + (size_t)C * Df * planeV;

    float4 ax = make_float4(0.f, 0.f, 0.f, 0.f), ay = ax;
    // d runs from the first multiple-of-4 boundary above Df-1 down to 0 in blocks of U (4 or 2) so that
    // s = d & 3 is a compile-time constant inside the unrolled body
    for (int dq = ((Df + 3) & ~3) - 1; dq >= 0; dq -= U) {
        float4 tl[U], ta[U], tb[U];
        const int q = dq >> 2;                      // same q for the U disparities of this block
        const bool inA = (wv + q) < Wv;
        const bool inB = (wv + q + 1) < Wv;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int d = dq - u;
            const bool dv = d < Df;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            tl[u] = (dv && (w0 + 3 >= d)) ? ld_stream(gl + (size_t)d * planeV) : z;
            ta[u] = (dv && inA) ? __ldg(gr + (size_t)d * planeV + q) : z;
            tb[u] = (dv && inB && ((U == 4 ? 3 - u : (d & 3)) > 0)) ? __ldg(gr + (size_t)d * planeV + q + 1) : z;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int d = dq - u;
            if (d < Df) {
                // gx[w] += g[d][w] for w >= d
                if (w0 + 0 >= d) ax.x = ax.x + tl[u].x;
                if (w0 + 1 >= d) ax.y = ax.y + tl[u].y;
                if (w0 + 2 >= d) ax.z = ax.z + tl[u].z;
                if (w0 + 3 >= d) ax.w = ax.w + tl[u].w;
                // gy[w'] += g[d][w'+d], elements s..s+3 of the 8-float window (A,B), s = 3-u
                if (inA) {
                    const int s = (U == 4) ? 3 - u : (d & 3);
                    const float e0 = s == 0 ? ta[u].x : s == 1 ? ta[u].y : s == 2 ? ta[u].z : ta[u].w;
                    const float e1 = s == 0 ? ta[u].y : s == 1 ? ta[u].z : s == 2 ? ta[u].w : tb[u].x;
                    const float e2 = s == 0 ? ta[u].z : s == 1 ? ta[u].w : s == 2 ? tb[u].x : tb[u].y;
                    const float e3 = s == 0 ? ta[u].w : s == 1 ? tb[u].x : s == 2 ? tb[u].y : tb[u].z;
                    // validity: w'+d < Wf  <=>  element index (s+k) stays inside A, or B exists
                    ay.x = ay.x + e0;                                   // s+0 <= 3: always inside A
                    if (s + 1 <= 3 || inB) ay.y = ay.y + e1;
                    if (s + 2 <= 3 || inB) ay.z = ay.z + e2;
                    if (s + 3 <= 3 || inB) ay.w = ay.w + e3;
                }
            }
        }
    }
    const size_t o = (size_t)(b * C + c) * planeV + p;
    reinterpret_cast<float4*>(gx)[o] = ax;
    reinterpret_cast<float4*>(gy)[o] = ay;
}

// grid: x over Hf*Wv vectors, y = c, z = b.
template <int NT, int U, int MINB>
__global__ void __launch_bounds__(NT, MINB)
cv_bwd_v4_kernel(const float* __restrict__ g, float* __restrict__ gx, float* __restrict__ gy,
                 int C, int Df, int Hf, int Wf) {
    const int Wv = Wf >> 2;
    const int PV = Hf * Wv;
    const int p = blockIdx.x * NT + threadIdx.x;
    if (p >= PV) return;
    cv_bwd_v4_unit<U>(g, gx, gy, C, Df, Wv, PV, p, blockIdx.y, blockIdx.z);
}

// Persistent form: a grid of `per_sm` CTAs per SM that is resident at once walks the items (b*C + c, block of NT vectors)
// in order.  Used when the kernel has to SHARE the SMs with the FP32-bound head backward on a second stream
// (rag_b200.pipeline.OverlappedTrainPath): launched first, its small footprint (NT threads x ~48 registers per CTA)
// leaves the register file to the head's CTAs, which the block scheduler can then dispatch beside it.
template <int NT, int U>
__global__ void __launch_bounds__(NT)
cv_bwd_v4_persistent_kernel(const float* __restrict__ g, float* __restrict__ gx, float* __restrict__ gy,
                            int BC, int C, int Df, int Hf, int Wf, int n_blk) {
    const int Wv = Wf >> 2;
    const int PV = Hf * Wv;
    const long long n_items = (long long)BC * n_blk;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int bc = (int)(item / n_blk), blk = (int)(item - (long long)bc * n_blk);
        const int p = blk * NT + threadIdx.x;
        if (p < PV) cv_bwd_v4_unit<U>(g, gx, gy, C, Df, Wv, PV, p, bc % C, bc / C);
    }
}

// Scalar path (any Wf).  grid: x over Hf*Wf elements, y = c, z = b.
template <int NT>
__global__ void __launch_bounds__(NT)
cv_bwd_scalar_kernel(const float* __restrict__ g, float* __restrict__ gx, float* __restrict__ gy,
                     int C, int Df, int Hf, int Wf) {
    const int PE = Hf * Wf;
    const int p = blockIdx.x * NT + threadIdx.x;
    if (p >= PE) return;
    const int c = blockIdx.y, b = blockIdx.z;
    const int w = p % Wf;
    const float* gl = g + (size_t)(b * 2 * C + c) * Df * PE + p;
    const float* gr = gl + (size_t)C * Df * PE;
    float ax = 0.f, ay = 0.f;
    for (int d = Df - 1; d >= 0; --d) {
        if (w >= d) ax = ax + ld_stream(gl + (size_t)d * PE);
        if (w + d < Wf) ay = ay + ld_stream(gr + (size_t)d * PE + d);
    }
    const size_t o = (size_t)(b * C + c) * PE + p;
    gx[o] = ax;
    gy[o] = ay;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static int check_cv_args(const void* a, const void* b_, const void* c, int B, int C, int Df, int Hf, int Wf) {
    if (!a || !b_ || !c) return fail(RAG_E_NULL, "cost_volume: null pointer");
    if (B <= 0 || C <= 0 || Df <= 0 || Hf <= 0 || Wf <= 0)
        return fail(RAG_E_SHAPE, "cost_volume: non-positive dimension B=%d C=%d Df=%d Hf=%d Wf=%d", B, C, Df, Hf, Wf);
    if (B > 65535 || C > 65535) return fail(RAG_E_SHAPE, "cost_volume: B and C must be <= 65535");
    // the kernels index (b*2C + c) and the item count B*C*tiles*chunks in 32-bit int
    if ((long long)2 * B * C >= (1LL << 31) || (long long)B * C * Hf * ((Df + 15) / 16) >= (1LL << 31))
        return fail(RAG_E_SHAPE, "cost_volume: 2*B*C and B*C*Hf*ceil(Df/16) must be < 2^31");
    if ((size_t)Df * Hf * Wf >= ((size_t)1 << 31) || Wf > 65535)
        return fail(RAG_E_SHAPE, "cost_volume: Df*Hf*Wf must be < 2^31 and Wf <= 65535");
    if (!aligned(a, 4) || !aligned(b_, 4) || !aligned(c, 4)) return fail(RAG_E_ALIGN, "cost_volume: pointers must be 4-byte aligned");
    return RAG_OK;
}

// any width / alignment: items (d, vector) dealt to threads, 16-disparity chunks, 36 KB tiles
template <int V, int NT>
static int launch_cv_fwd(const float* x, const float* y, float* cost, int B, int C, int Df, int Hf, int Wf, cudaStream_t st) {
    int dchunk = ((16 + V - 1) / V) * V;
    const int n_dchunks = (Df + dchunk - 1) / dchunk;
    const size_t budget = 36 * 1024;
    int R = (int)(budget / ((size_t)(V + 1) * 4 * Wf + (size_t)2 * (Wf / V)));
    R = R < 1 ? 1 : (R > Hf ? Hf : R);
    // prefer an R that divides Hf (no ragged last tile) when one is close
    for (int r = R; r >= (R * 3) / 4 && r >= 1; --r)
        if (Hf % r == 0) { R = r; break; }
    const size_t smem = (size_t)(V + 1) * R * Wf * 4 + (size_t)R * (Wf / V) * 2;
    if (smem > 200 * 1024) return fail(RAG_E_SHAPE, "cost_volume_fwd: Wf=%d too wide for one shared-memory row tile", Wf);
    auto kern = cv_fwd_kernel<V, NT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail((int)e, "cost_volume_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    const int n_tiles = (Hf + R - 1) / R;
    if ((long long)n_tiles * n_dchunks > 2147483647LL) return fail(RAG_E_SHAPE, "cost_volume_fwd: grid too large");
    dim3 grid(n_tiles * n_dchunks, C, B);
    kern<<<grid, NT, smem, st>>>(x, y, cost, C, Df, Hf, Wf, R, dchunk, n_dchunks);
    return check_launch("cost_volume_fwd");
}

static bool cv_lean_ok(const float* x, const float* y, const float* cost, int Wf) {
    return Wf % 4 == 0 && Wf <= 2048 && aligned(x, 16) && aligned(y, 16) && aligned(cost, 16);
}

// zero the caller's work counter on the launch stream (stream-ordered before the kernel that consumes it)
static int arm_counter(void* workspace, cudaStream_t st, const char* who) {
    if (!aligned(workspace, 8)) return fail(RAG_E_ALIGN, "%s: workspace must be 8-byte aligned", who);
    cudaError_t e = cudaMemsetAsync(workspace, 0, RAG_CV_FWD_WORKSPACE_BYTES, st);
    if (e != cudaSuccess) return fail((int)e, "%s: cudaMemsetAsync(workspace): %s", who, cudaGetErrorString(e));
    return RAG_OK;
}

// lean thread-stationary kernel.  ctr == nullptr: one CTA per item (hardware dispatch order keeps the resident
// CTAs on neighbouring rows; stateless).  ctr != nullptr: persistent grid of `per_sm` CTAs per SM that takes its
// items IN ORDER from the caller's counter (3 % faster at B=8 480x960, and resident at once, which is what lets the
// disparity head share the SMs on a second stream).
template <int NT, int VPT>
static int launch_cv_fwd_lean(const float* x, const float* y, float* cost, int B, int C, int Df, int Hf, int Wf,
                              int per_sm, unsigned int* ctr, cudaStream_t st) {
    const int Wv = Wf / 4, Df4 = (Df + 3) & ~3;
    const size_t row_bytes = (size_t)16 * (Df4 + Wf + 4);          // four images of one row
    int R = std::min(Hf, (NT * VPT) / Wv);
    if (R < 1) return fail(RAG_E_SHAPE, "cost_volume_fwd(lean): Wf=%d too wide for a %d-vector tile", Wf, NT * VPT);
    while (R > 1 && R * row_bytes > 64 * 1024) --R;
    const size_t smem = R * row_bytes;
    if (smem > 200 * 1024) return fail(RAG_E_SHAPE, "cost_volume_fwd(lean): Df=%d Wf=%d do not fit shared memory", Df, Wf);
    const bool dyn = ctr != nullptr;
    auto kern = dyn ? cv_fwd_lean_kernel<NT, VPT, true> : cv_fwd_lean_kernel<NT, VPT, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // the L1/shared split of an SM cannot change while CTAs are resident: ask for the largest shared-memory
    // carve-out so that this kernel and the disparity head (which does the same) can share an SM
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return fail((int)e, "cost_volume_fwd(lean): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const int n_tiles = (Hf + R - 1) / R;
    const int dchunk = std::min(16, (Df + 3) & ~3);
    const int n_dchunks = (Df + dchunk - 1) / dchunk;
    const long long n_items = (long long)B * C * n_tiles * n_dchunks;
    if (n_items >= (1LL << 31)) return fail(RAG_E_SHAPE, "cost_volume_fwd(lean): too many work items");
    const int grid = (int)(dyn ? std::min<long long>(n_items, (long long)num_sms() * per_sm) : n_items);
    if (dyn) if (int rc = arm_counter(ctr, st, "cost_volume_fwd(lean)")) return rc;
    kern<<<grid, NT, smem, st>>>(x, y, cost, B * C, C, Df, Hf, Wf, R, n_tiles, dchunk, n_dchunks, ctr);
    return check_launch("cost_volume_fwd(lean)");
}

static bool cv_tma_ok(const float* x, const float* y, const float* cost, int Df, int Wf) {
    return Wf % 4 == 0 && Df % 4 == 0 && Df <= Wf && aligned(x, 16) && aligned(y, 16) && aligned(cost, 16);
}

static int launch_cv_fwd_tma(const float* x, const float* y, float* cost, int B, int C, int Df, int Hf, int Wf,
                             unsigned int* ctr, cudaStream_t st) {
    constexpr int NT = 128;
    int R = std::min(4, Hf);
    const int per_sm = 2;
    const size_t row_floats = (size_t)Wf + 4 * (size_t)(Df + Wf + 4);
    while (R > 1 && 4 * (Df + 2 * R * row_floats) > 200 * 1024) --R;
    const size_t smem = 4 * (Df + 2 * R * row_floats);
    if (smem > 200 * 1024) return fail(RAG_E_SHAPE, "cost_volume_fwd(tma): Wf=%d too wide for shared memory", Wf);
    auto kern = cv_fwd_tma_kernel<NT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "cost_volume_fwd(tma): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const int n_tiles = (Hf + R - 1) / R;
    const int dchunk = std::min(16, Df), n_dchunks = (Df + dchunk - 1) / dchunk;
    const long long n_items = (long long)B * C * n_tiles * n_dchunks;
    if (n_items >= (1LL << 31)) return fail(RAG_E_SHAPE, "cost_volume_fwd(tma): too many work items");
    const int grid = (int)std::min<long long>(n_items, (long long)num_sms() * per_sm);
    if (int rc = arm_counter(ctr, st, "cost_volume_fwd(tma)")) return rc;
    kern<<<grid, NT, smem, st>>>(x, y, cost, B, C, Df, Hf, Wf, R, n_tiles, dchunk, n_dchunks, ctr);
    return check_launch("cost_volume_fwd(tma)");
}

// variant: -1 = default (lean persistent when a workspace is given, lean one-CTA-per-item without, generic when the
// lean preconditions fail); 0 = generic; 1 = lean one CTA per item; 2 = lean persistent 256 threads (needs workspace);
// 3 = lean persistent 512 threads, RAG_CV_FWD_SHARED (needs workspace); 4 = TMA bulk-store kernel (needs workspace)
int cost_volume_fwd(const float* x, const float* y, float* cost, int B, int C, int Df, int Hf, int Wf,
                    void* workspace, int variant, cudaStream_t st) {
    if (int e = check_cv_args(x, y, cost, B, C, Df, Hf, Wf)) return e;
    if (variant < -1 || variant > 4) return fail(RAG_E_VARIANT, "cost_volume_fwd: unknown variant %d", variant);
    unsigned int* ctr = static_cast<unsigned int*>(workspace);
    const bool lean = cv_lean_ok(x, y, cost, Wf) && (Wf / 4) <= 512;
    if (variant == -1) variant = lean ? (ctr ? RAG_CV_FWD_LEAN : 1) : 0;
    if (variant >= 2 && !ctr) return fail(RAG_E_NULL, "cost_volume_fwd: variant %d needs a workspace of RAG_CV_FWD_WORKSPACE_BYTES", variant);
    if (variant >= 1 && variant <= 3 && !lean)
        return fail(RAG_E_VARIANT, "cost_volume_fwd: lean variants need Wf %% 4 == 0, Wf <= 2048 and 16-byte aligned pointers");
    if (variant == 1) return launch_cv_fwd_lean<256, 2>(x, y, cost, B, C, Df, Hf, Wf, 0, nullptr, st);
    if (variant == 2) return launch_cv_fwd_lean<256, 2>(x, y, cost, B, C, Df, Hf, Wf, 1, ctr, st);
    if (variant == 3) return launch_cv_fwd_lean<512, 1>(x, y, cost, B, C, Df, Hf, Wf, 1, ctr, st);
    if (variant == 4) {
        if (!cv_tma_ok(x, y, cost, Df, Wf))
            return fail(RAG_E_VARIANT, "cost_volume_fwd: the TMA variant needs Wf %% 4 == 0, Df %% 4 == 0, Df <= Wf and 16-byte aligned pointers");
        return launch_cv_fwd_tma(x, y, cost, B, C, Df, Hf, Wf, ctr, st);
    }
    const bool a16 = aligned(x, 16) && aligned(y, 16) && aligned(cost, 16);
    const bool a8 = aligned(x, 8) && aligned(y, 8) && aligned(cost, 8);
    if (Wf % 4 == 0 && a16) return launch_cv_fwd<4, 256>(x, y, cost, B, C, Df, Hf, Wf, st);
    if (Wf % 2 == 0 && a8) return launch_cv_fwd<2, 256>(x, y, cost, B, C, Df, Hf, Wf, st);
    return launch_cv_fwd<1, 256>(x, y, cost, B, C, Df, Hf, Wf, st);
}

// variant: 0 = default (128-bit vector kernel when Wf % 4 == 0 and aligned, else scalar); 1 = scalar;
// 2 = RAG_CV_BWD_SHARED: the vector kernel as a persistent grid of 4 x 128-thread CTAs per SM (SM sharing)
int cost_volume_bwd(const float* g, float* gx, float* gy, int B, int C, int Df, int Hf, int Wf,
                    int variant, cudaStream_t st) {
    if (int e = check_cv_args(g, gx, gy, B, C, Df, Hf, Wf)) return e;
    if (variant < 0 || variant > 2) return fail(RAG_E_VARIANT, "cost_volume_bwd: unknown variant %d", variant);
    const bool a16 = aligned(g, 16) && aligned(gx, 16) && aligned(gy, 16);
    if (variant == 2) {
        if (!(Wf % 4 == 0 && a16)) return fail(RAG_E_VARIANT, "cost_volume_bwd: variant 2 needs Wf %% 4 == 0 and 16-byte aligned pointers");
        constexpr int NT = 128;
        const int PV = Hf * (Wf / 4), n_blk = (PV + NT - 1) / NT;
        const long long n_items = (long long)B * C * n_blk;
        const int grid = (int)std::min<long long>(n_items, (long long)num_sms() * 4);
        cv_bwd_v4_persistent_kernel<NT, 4><<<grid, NT, 0, st>>>(g, gx, gy, B * C, C, Df, Hf, Wf, n_blk);
    } else if (variant != 1 && Wf % 4 == 0 && a16) {
        constexpr int NT = 128;
        const int PV = Hf * (Wf / 4);
        dim3 grid((PV + NT - 1) / NT, C, B);
        cv_bwd_v4_kernel<NT, 4, 1><<<grid, NT, 0, st>>>(g, gx, gy, C, Df, Hf, Wf);
    } else {
        constexpr int NT = 256;
        const int PE = Hf * Wf;
        dim3 grid((PE + NT - 1) / NT, C, B);
        cv_bwd_scalar_kernel<NT><<<grid, NT, 0, st>>>(g, gx, gy, C, Df, Hf, Wf);
    }
    return check_launch("cost_volume_bwd");
}

}  // namespace rag
