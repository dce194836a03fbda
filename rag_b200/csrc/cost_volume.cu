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

// Persistent form: a fixed number of CTAs (a multiple of the SM count) walks the work units.  Used when the
// kernel has to SHARE every SM with the disparity-head kernel on a second stream (rag_b200.pipeline): a
// resident footprint of exactly k CTAs per SM leaves the rest of the register file / shared memory free.
template <int V, int NT>
__global__ void __launch_bounds__(NT)
cv_fwd_persistent_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ cost,
                         int B, int C, int Df, int Hf, int Wf, int R, int dchunk, int n_dchunks, int n_tiles) {
    extern __shared__ __align__(16) float smem[];
    const int per_bc = n_tiles * n_dchunks;
    const long long total = (long long)B * C * per_bc;
    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
        const int bc = (int)(u / per_bc);
        const int rem = (int)(u - (long long)bc * per_bc);
        const int tile = rem / n_dchunks;
        cv_fwd_unit<V, NT>(x, y, cost, C, Df, Hf, Wf, R, dchunk, tile, rem - tile * n_dchunks, bc % C, bc / C, smem);
        __syncthreads();   // the staging buffers are reused by the next unit
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// Vector path (Wf % 4 == 0).  grid: x over Hf*Wv vectors, y = c, z = b.
template <int NT, int U, int MINB>
__global__ void __launch_bounds__(NT, MINB)
cv_bwd_v4_kernel(const float* __restrict__ g, float* __restrict__ gx, float* __restrict__ gy,
                 int C, int Df, int Hf, int Wf) {
    const int Wv = Wf >> 2;
    const int PV = Hf * Wv;
    const int p = blockIdx.x * NT + threadIdx.x;
    if (p >= PV) return;
    const int c = blockIdx.y, b = blockIdx.z;
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
    if ((size_t)Df * Hf * Wf >= ((size_t)1 << 31) || Wf > 65535)
        return fail(RAG_E_SHAPE, "cost_volume: Df*Hf*Wf must be < 2^31 and Wf <= 65535");
    if (!aligned(a, 4) || !aligned(b_, 4) || !aligned(c, 4)) return fail(RAG_E_ALIGN, "cost_volume: pointers must be 4-byte aligned");
    return RAG_OK;
}

template <int V, int NT>
static int launch_cv_fwd(const float* x, const float* y, float* cost, int B, int C, int Df, int Hf, int Wf,
                         int variant, size_t smem_floor, cudaStream_t st) {
    // variant: 0 -> 16-disparity chunks, 36 KB tiles; 1 -> 32-disparity chunks; 2 -> whole sweep per CTA;
    //          3 -> 16-disparity chunks, 72 KB tiles
    int dchunk = variant == 1 ? 32 : variant == 2 ? Df : 16;
    dchunk = ((dchunk + V - 1) / V) * V;
    const int n_dchunks = (Df + dchunk - 1) / dchunk;
    const size_t budget = variant == 3 ? 72 * 1024 : 36 * 1024;
    int R = (int)(budget / ((size_t)(V + 1) * 4 * Wf + (size_t)2 * (Wf / V)));
    R = R < 1 ? 1 : (R > Hf ? Hf : R);
    // prefer an R that divides Hf (no ragged last tile) when one is close
    for (int r = R; r >= (R * 3) / 4 && r >= 1; --r)
        if (Hf % r == 0) { R = r; break; }
    size_t smem = (size_t)(V + 1) * R * Wf * 4 + (size_t)R * (Wf / V) * 2;
    if (smem < smem_floor) smem = smem_floor;   // occupancy experiments only
    if (smem > 200 * 1024) return fail(RAG_E_SHAPE, "cost_volume_fwd: Wf=%d too wide for one shared-memory row tile", Wf);
    auto kern = cv_fwd_kernel<V, NT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail((int)e, "cost_volume_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    const int n_tiles = (Hf + R - 1) / R;
    dim3 grid(n_tiles * n_dchunks, C, B);
    kern<<<grid, NT, smem, st>>>(x, y, cost, C, Df, Hf, Wf, R, dchunk, n_dchunks);
    return check_launch("cost_volume_fwd");
}

static bool cv_lean_ok(const float* x, const float* y, const float* cost, int Wf) {
    return Wf % 4 == 0 && Wf <= 2048 && aligned(x, 16) && aligned(y, 16) && aligned(cost, 16);
}

template <int NT = 256, int VPT = 2, int ST = 0>
static int launch_cv_fwd_lean(const float* x, const float* y, float* cost, int B, int C, int Df, int Hf, int Wf,
                              int per_sm, int dchunk, cudaStream_t st, size_t smem_floor = 0, bool dyn = false) {
    const int Wv = Wf / 4, Df4 = (Df + 3) & ~3;
    const size_t row_bytes = (size_t)16 * (Df4 + Wf + 4);          // four images of one row
    int R = std::min(Hf, (NT * VPT) / Wv);
    while (R > 1 && R * row_bytes > (NT * VPT > 512 ? 100 : 64) * 1024) --R;
    const size_t smem = std::max(R * row_bytes, smem_floor);
    if (smem > 200 * 1024) return fail(RAG_E_SHAPE, "cost_volume_fwd(lean): Df=%d Wf=%d do not fit shared memory", Df, Wf);
    static std::atomic<unsigned> ticket{0};
    auto kern = dyn ? cv_fwd_lean_kernel<NT, VPT, true, ST> : cv_fwd_lean_kernel<NT, VPT, false, ST>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // the L1/shared split of an SM cannot change while CTAs are resident: ask for the largest shared-memory
    // carve-out so that this kernel and the disparity head (which does the same) can share an SM
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return fail((int)e, "cost_volume_fwd(lean): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const int n_tiles = (Hf + R - 1) / R;
    if (dchunk <= 0 || dchunk > Df) dchunk = (Df + 3) & ~3;
    const int n_dchunks = (Df + dchunk - 1) / dchunk;
    const long long n_items = (long long)B * C * n_tiles * n_dchunks;
    const int grid = (int)(per_sm > 0 ? std::min<long long>(n_items, (long long)kNumSMs * per_sm) : n_items);
    const int slot = dyn ? (int)(ticket.fetch_add(1) % 64u) : 0;
    kern<<<grid, NT, smem, st>>>(x, y, cost, B * C, C, Df, Hf, Wf, R, n_tiles, dchunk, n_dchunks, slot);
    return check_launch("cost_volume_fwd(lean)");
}

static bool cv_tma_ok(const float* x, const float* y, const float* cost, int Df, int Wf) {
    return Wf % 4 == 0 && Df % 4 == 0 && Df <= Wf && aligned(x, 16) && aligned(y, 16) && aligned(cost, 16);
}

static int launch_cv_fwd_tma(const float* x, const float* y, float* cost, int B, int C, int Df, int Hf, int Wf,
                             int R, int per_sm, int mode, cudaStream_t st) {
    constexpr int NT = 128;
    R = R > Hf ? Hf : R;
    const size_t row_floats = (size_t)Wf + 4 * (size_t)(Df + Wf + 4);
    while (R > 1 && 4 * (Df + 2 * R * row_floats) > 200 * 1024) --R;
    const size_t smem = 4 * (Df + 2 * R * row_floats);
    if (smem > 200 * 1024) return fail(RAG_E_SHAPE, "cost_volume_fwd(tma): Wf=%d too wide for shared memory", Wf);
    auto kern = mode == 1 ? cv_fwd_tma_kernel<NT, 1> : mode == 2 ? cv_fwd_tma_kernel<NT, 2> : cv_fwd_tma_kernel<NT, 0>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "cost_volume_fwd(tma): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    static std::atomic<unsigned> ticket{0};
    const int n_tiles = (Hf + R - 1) / R;
    const int dchunk = std::min(16, Df), n_dchunks = (Df + dchunk - 1) / dchunk;
    const long long n_items = (long long)B * C * n_tiles * n_dchunks;
    const int grid = (int)std::min<long long>(n_items, (long long)kNumSMs * per_sm);
    kern<<<grid, NT, smem, st>>>(x, y, cost, B, C, Df, Hf, Wf, R, n_tiles, dchunk, n_dchunks, (int)(ticket.fetch_add(1) % 64u));
    return check_launch("cost_volume_fwd(tma)");
}

int cost_volume_fwd(const float* x, const float* y, float* cost, int B, int C, int Df, int Hf, int Wf,
                    int variant, cudaStream_t st) {
    if (int e = check_cv_args(x, y, cost, B, C, Df, Hf, Wf)) return e;
    if (variant < -1 || variant > 36) return fail(RAG_E_VARIANT, "cost_volume_fwd: unknown variant %d", variant);
    if (variant == -1) variant = cv_lean_ok(x, y, cost, Wf) ? RAG_CV_FWD_LEAN : 0;
    if (variant >= 18) {   // lean thread-stationary kernel, persistent: 18 = 2 CTAs/SM, 19 = 1, 20 = 3, 21 = 4
        if (!cv_lean_ok(x, y, cost, Wf))
            return fail(RAG_E_VARIANT, "cost_volume_fwd: lean variants need Wf %% 4 == 0, Wf <= 2048 and 16-byte aligned pointers");
        // 22-25: 16-disparity chunks per item; 22 = 2 CTAs/SM, 23 = 1, 24 = 4, 25 = one CTA per item (not persistent)
        // 31 / 32: as 29 with 128 threads x 4 vectors / 512 threads x 1 vector per thread
        // 33 / 34: as 29 with plain / write-through stores instead of st.global.cs (A/B of the store flavour)
        // 35 / 36: as 29 with twice the rows per tile (512 x 2 / 256 x 4 vectors per thread)
        if (variant == 35) return launch_cv_fwd_lean<512, 2>(x, y, cost, B, C, Df, Hf, Wf, 1, 16, st, 0, true);
        if (variant == 36) return launch_cv_fwd_lean<256, 4>(x, y, cost, B, C, Df, Hf, Wf, 1, 16, st, 0, true);
        if (variant == 33) return launch_cv_fwd_lean<256, 2, 1>(x, y, cost, B, C, Df, Hf, Wf, 1, 16, st, 0, true);
        if (variant == 34) return launch_cv_fwd_lean<256, 2, 2>(x, y, cost, B, C, Df, Hf, Wf, 1, 16, st, 0, true);
        if (variant == 31) return launch_cv_fwd_lean<128, 4>(x, y, cost, B, C, Df, Hf, Wf, 1, 16, st, 0, true);
        if (variant == 32) return launch_cv_fwd_lean<512, 1>(x, y, cost, B, C, Df, Hf, Wf, 1, 16, st, 0, true);
        // 28-30: persistent with in-order (atomic counter) item hand-out, 16-disparity chunks: 2 / 1 / 4 CTAs per SM
        if (variant >= 28) return launch_cv_fwd_lean(x, y, cost, B, C, Df, Hf, Wf, variant == 28 ? 2 : variant == 29 ? 1 : 4, 16, st, 0, true);
        // 26 / 27: as 25 with the shared-memory request padded so that only 1 / 2 CTAs fit an SM
        if (variant >= 26) return launch_cv_fwd_lean(x, y, cost, B, C, Df, Hf, Wf, 0, 16, st, variant == 26 ? 120 * 1024 : 76 * 1024);
        if (variant >= 22) return launch_cv_fwd_lean(x, y, cost, B, C, Df, Hf, Wf, variant == 22 ? 2 : variant == 23 ? 1 : variant == 24 ? 4 : 0, 16, st);
        return launch_cv_fwd_lean(x, y, cost, B, C, Df, Hf, Wf, variant == 18 ? 2 : variant == 19 ? 1 : variant == 20 ? 3 : 4, 0, st);
    }
    if (variant >= 11) {   // TMA bulk-store kernel: 11 = R4 x 1 CTA/SM, 12 = R4 x 2, 13 = R2 x 2, 14 = R8 x 1
        if (!cv_tma_ok(x, y, cost, Df, Wf))
            return fail(RAG_E_VARIANT, "cost_volume_fwd: TMA variants need Wf %% 4 == 0, Df %% 4 == 0, Df <= Wf and 16-byte aligned pointers");
        // 15 = R2 x 4 CTAs/SM; 16 / 17 = EXPERIMENTS writing only the right / left half (output incomplete)
        const int R = (variant == 13 || variant == 15) ? 2 : variant == 14 ? 8 : 4;
        const int per_sm = variant == 15 ? 4 : (variant == 11 || variant == 14) ? 1 : 2;
        return launch_cv_fwd_tma(x, y, cost, B, C, Df, Hf, Wf, R, per_sm, variant == 16 ? 1 : variant == 17 ? 2 : 0, st);
    }
    if (variant >= 8) {   // persistent: exactly 3 (variant 8), 4 (9) or 2 (10) 128-thread CTAs per SM, ~20 KB tiles
        if (!(Wf % 4 == 0 && aligned(x, 16) && aligned(y, 16) && aligned(cost, 16)))
            return fail(RAG_E_VARIANT, "cost_volume_fwd: persistent variants need Wf %% 4 == 0 and 16-byte aligned pointers");
        constexpr int V = 4, NT = 128;
        const int per_sm = variant == 8 ? 3 : variant == 9 ? 4 : 2;
        const int dchunk = 16, n_dchunks = (Df + dchunk - 1) / dchunk;
        int R = (int)((20 * 1024) / ((size_t)(V + 1) * 4 * Wf + (size_t)2 * (Wf / V)));
        R = R < 1 ? 1 : (R > Hf ? Hf : R);
        for (int r = R; r >= (R * 3) / 4 && r >= 1; --r)
            if (Hf % r == 0) { R = r; break; }
        const size_t smem = (size_t)(V + 1) * R * Wf * 4 + (size_t)R * (Wf / V) * 2;
        if (smem > 48 * 1024) return fail(RAG_E_SHAPE, "cost_volume_fwd: Wf=%d too wide for the persistent variant", Wf);
        const int n_tiles = (Hf + R - 1) / R;
        cv_fwd_persistent_kernel<V, NT><<<kNumSMs * per_sm, NT, smem, st>>>(x, y, cost, B, C, Df, Hf, Wf, R, dchunk, n_dchunks, n_tiles);
        return check_launch("cost_volume_fwd(persistent)");
    }
    const bool a16 = aligned(x, 16) && aligned(y, 16) && aligned(cost, 16);
    const bool a8 = aligned(x, 8) && aligned(y, 8) && aligned(cost, 8);
    if (variant >= 4) {   // occupancy experiments: 128-thread CTAs, 4: unconstrained, 5: <=4, 6: <=2, 7: 1 CTA per SM
        if (!(Wf % 4 == 0 && a16)) return fail(RAG_E_VARIANT, "cost_volume_fwd: variants 4-7 need Wf %% 4 == 0");
        const size_t floor_b = variant == 5 ? 50 * 1024 : variant == 6 ? 100 * 1024 : variant == 7 ? 190 * 1024 : 0;
        return launch_cv_fwd<4, 128>(x, y, cost, B, C, Df, Hf, Wf, 0, floor_b, st);
    }
    if (Wf % 4 == 0 && a16) return launch_cv_fwd<4, 256>(x, y, cost, B, C, Df, Hf, Wf, variant, 0, st);
    if (Wf % 2 == 0 && a8) return launch_cv_fwd<2, 256>(x, y, cost, B, C, Df, Hf, Wf, variant, 0, st);
    return launch_cv_fwd<1, 256>(x, y, cost, B, C, Df, Hf, Wf, variant, 0, st);
}

int cost_volume_bwd(const float* g, float* gx, float* gy, int B, int C, int Df, int Hf, int Wf,
                    int variant, cudaStream_t st) {
    if (int e = check_cv_args(g, gx, gy, B, C, Df, Hf, Wf)) return e;
    if (variant < 0 || variant > 3) return fail(RAG_E_VARIANT, "cost_volume_bwd: unknown variant %d", variant);
    const bool a16 = aligned(g, 16) && aligned(gx, 16) && aligned(gy, 16);
    if (variant != 1 && Wf % 4 == 0 && a16) {
        constexpr int NT = 128;
        const int PV = Hf * (Wf / 4);
        dim3 grid((PV + NT - 1) / NT, C, B);
        // variant 3 = 12-CTA/SM build (2 disparities in flight, 40 registers): measured SLOWER than the 4-deep
        // build at every size (profiles/r1_kbench_cv_bwd.jsonl), kept for A/B only
        const bool hi_occ = variant == 3;
        if (hi_occ) cv_bwd_v4_kernel<NT, 2, 12><<<grid, NT, 0, st>>>(g, gx, gy, C, Df, Hf, Wf);
        else cv_bwd_v4_kernel<NT, 4, 1><<<grid, NT, 0, st>>>(g, gx, gy, C, Df, Hf, Wf);
    } else {
        constexpr int NT = 256;
        const int PE = Hf * Wf;
        dim3 grid((PE + NT - 1) / NT, C, B);
        cv_bwd_scalar_kernel<NT><<<grid, NT, 0, st>>>(g, gx, gy, C, Df, Hf, Wf);
    }
    return check_launch("cost_volume_bwd");
}

}  // namespace rag
