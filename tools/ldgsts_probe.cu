// Probe: shared-memory wavefronts of a 16-byte cp.async.cg (LDGSTS.E.BYPASS.128) per warp instruction as a function of how
// the 32 lanes' global addresses sit relative to 32-byte sectors and 128-byte lines.  Read with ncu (source page, column
// "L1 Wavefronts Shared" of the LDGSTS line).   nvcc -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a tools/ldgsts_probe.cu -o tools/ldgsts_probe
#include <cstdio>
#include <cuda_pipeline.h>

template <int MODE>
__global__ void __launch_bounds__(128) probe(const float* __restrict__ g, float* __restrict__ out, int iters) {
    __shared__ __align__(128) float buf[4][4][128];          // per warp: 4 stages x 512 B
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t wbase = ((size_t)blockIdx.x * 4 + warp) * (size_t)iters * 256;     // floats; 1 KB per iteration per warp
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        const float* src = g + wbase + (size_t)it * 256;
        int off;                                                // float offset of the lane's 16 bytes
        if (MODE == 0) off = lane * 4;                          // 512 B contiguous, line aligned
        else if (MODE == 1) off = lane * 4 + 4;                 // contiguous, shifted by 16 B (half a sector)
        else if (MODE == 2) off = lane * 4 + 8;                 // contiguous, shifted by 32 B (sector aligned, not line aligned)
        else off = (lane < 18 ? lane * 4 + 4 : 100 + (lane - 18) * 4 + 4);   // two runs (288 B + 224 B), both 16 B shifted
        float* dst = &buf[warp][it & 3][lane * 4];
        __pipeline_memcpy_async(dst, src + off, 16);
        __pipeline_commit();
        __pipeline_wait_prior(2);
        __syncwarp();
        acc += buf[warp][(it + 2) & 3][lane * 4];
    }
    __pipeline_wait_prior(0);
    if (acc == 123.456f) out[threadIdx.x] = acc;
}

int main() {
    const int iters = 256, blocks = 148 * 4;
    float *g, *out;
    cudaMalloc(&g, (size_t)blocks * 4 * iters * 1024 + 4096);
    cudaMemset(g, 0, (size_t)blocks * 4 * iters * 1024 + 4096);
    cudaMalloc(&out, 4096);
    probe<0><<<blocks, 128>>>(g, out, iters);
    probe<1><<<blocks, 128>>>(g, out, iters);
    probe<2><<<blocks, 128>>>(g, out, iters);
    probe<3><<<blocks, 128>>>(g, out, iters);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
