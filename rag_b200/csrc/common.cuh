// Shared device/host helpers for librag_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/rag_b200.h"

namespace rag {

// ---- host side: error text + launch accounting ----------------------------------------------
extern thread_local char g_last_error[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

inline int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
    return RAG_OK;
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// SM count of the current device (148 on a B200; smaller under MIG), queried once per device
inline int num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// ---- device side ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// streaming (evict-first) 16/8/4-byte global stores: the volume is written once and is far larger
// than L2, so it should not displace the feature rows other CTAs are about to read
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float2* p, float2 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

// streaming loads (read once): bypass L1 allocation where it does not help
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ float2 ld_stream(const float2* p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }

// PyTorch's linear-upsample source index (align_corners=False), fp32:
// ATen/native/cuda/UpSample.cuh:115-130.  FMA=true is what nvcc emits for the ATen kernel.
template <bool FMA>
__device__ __forceinline__ void src_index(float scale, int dst, int n_in, int& i0, int& i1, float& l0, float& l1) {
    float s;
    if (FMA)
        s = __fmaf_rn(scale, (float)dst + 0.5f, -0.5f);
    else
        s = __fsub_rn(__fmul_rn(scale, (float)dst + 0.5f), 0.5f);
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    l1 = s - (float)i0;
    l0 = 1.f - l1;
}

#endif  // __CUDACC__

}  // namespace rag
