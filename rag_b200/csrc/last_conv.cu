// `last_3_3d`: the Matching Net's last layer, the producer of the disparity head's input (SURVEY.md
// section 8f rank 2).  Reference: src/models/rag_model.py:269 `ConvBR_3d(initial_fm, 1, 3, 1, 1, bn=False,
// relu=False)` applied at :361-365 -> a bias-free Conv3d C -> 1, 3x3x3, stride 1, zero padding 1:
//     out[b,0,d,h,w] = sum_{c,kd,kh,kw} W[0,c,kd,kh,kw] * in[b,c,d+kd-1,h+kh-1,w+kw-1]
// Through cuDNN this layer costs 6.6 ms (TF32) / 32 ms (fp32) per 8 pairs at 480x960 on a B200 (one output
// channel wastes a tensor-core tile); its floors are 0.19 ms of HBM (read the 1.26 GB input once) and 0.23 ms
// of FP32 (8.5 GFMA at 128 lanes/clk/SM), so it is written here as a register-tiled fp32 direct convolution.
//
// A CTA owns a 16 x 8 x 64 (d,h,w) output tile of one pair and walks the C input channels; the channel's
// 18 x 10 x 72 input tile (halo included, zero-filled outside the volume by cp.async zfill) is staged in
// shared memory (one stage; the three resident CTAs of an SM cover each other's copies).  A thread owns 4 (d) x 8 (w) outputs of one h: per input row (d', h') it gets the ten
// inputs it needs from two conflict-free LDS.128 plus two warp shuffles (the neighbours' edge elements) and
// does up to 72 FFMAs with the channel's 27 weights held in registers: 864 FFMA per channel.
// Fusing this layer INTO the head kernel was rejected: the head's CTA window has a (6/4)x(40/32) halo, so the
// 324-MAC convolution would be recomputed 1.9x; instead the 105 MB result stays L2-resident (126 MB L2) for
// the head that runs next.
#include <cuda_pipeline.h>

#include <mutex>

#include "common.cuh"

namespace rag {

constexpr int kLcDT = 16, kLcHT = 8, kLcWT = 64;            // output tile
constexpr int kLcSD = kLcDT + 2, kLcSH = kLcHT + 2;        // staged tile: d in [d0-1, d0+17), h in [h0-1, h0+9)
constexpr int kLcRowVecs = (kLcWT + 8) / 4;                 // ... w in [w0-4, w0+68): 18 vectors per row
// Shared-memory rows are SKEWED: 4 floats of padding after every 32, so that the eight 8-float segments of
// a row start in distinct bank groups and the threads' LDS.128 are conflict-free (unskewed: 2-way).
constexpr int kLcSW = kLcWT + 8 + 8;                        // row stride (80 floats)
__host__ __device__ constexpr int lc_skew(int col) { return col + 4 * (col >> 5); }
constexpr int kLcRows = kLcSD * kLcSH;                      // 180 rows per stage
constexpr int kLcStage = kLcRows * kLcSW;                   // floats per stage (14400)


// Weights live in CONSTANT memory (copied there device-to-device, stream-ordered, before the launch) so that
// the FFMAs take them as uniform-register operands instead of 27 vector registers per thread (168 registers at
// three CTAs per SM leave no room for them).  Constant memory cannot be caller-owned, so this is the ONE piece
// of per-device state the library keeps, and it is made race-free instead of hidden: eight slots are used
// round-robin, every launch records an event after its kernel, and the next user of a slot -- on whatever
// stream -- first makes its stream wait for that event, so a slot is never rewritten under a kernel that reads
// it (a cross-stream dependency appears only when a slot comes round again, eight launches later).  The
// event chain cannot be expressed inside a stream capture: a capturing stream is refused with RAG_E_CAPTURE
// (callers fall back to the reference's own nn.Conv3d there).
constexpr int kLcMaxC = 64, kLcSlotsW = 8;
__constant__ float c_lcw[kLcSlotsW][kLcMaxC * 27];
struct LcRing {
    std::mutex mu;
    unsigned next = 0;
    cudaEvent_t ev[kLcSlotsW] = {};
    bool have[kLcSlotsW] = {};
};
static LcRing g_lc_ring[64];   // per device

// grid: x = ceil(W/64), y = ceil(H/8), z = B * ceil(D/16); 256 threads.
// smem: STAGES x kLcStage floats
// TH = output rows (h) per thread: 1 -> 256 threads, 2 -> 128 threads (twice the FMAs per shared-memory load)
template <int STAGES, int TH>
__global__ void __launch_bounds__(256 / TH, STAGES == 1 ? 3 : 1)
conv3d_c1_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out,
                 int C, int D, int H, int W, int n_dt, int wslot) {
    extern __shared__ __align__(16) float lc_smem[];
    const int tid = threadIdx.x;
    const int b = blockIdx.z / n_dt, dt = blockIdx.z - b * n_dt;
    const int d0 = dt * kLcDT, h0 = blockIdx.y * kLcHT, w0 = blockIdx.x * kLcWT;
    const size_t chan = (size_t)D * H * W;
    const float* inb = in + (size_t)b * C * chan;

    // staging: a thread copies vector `svec` of rows srow0, srow0 + kLcRowGroups, ... (source offset inside a channel, or -1)
    constexpr int NT = 256 / TH;
    constexpr int kLcRowGroups = NT / kLcRowVecs;                            // rows staged per pass (14 or 7)
    constexpr int kLcSlots = (kLcRows + kLcRowGroups - 1) / kLcRowGroups;   // passes (13 or 26)
    const int srow0 = tid / kLcRowVecs, svec = tid - srow0 * kLcRowVecs;
    const bool stager = srow0 < kLcRowGroups;
    int goff[kLcSlots];
#pragma unroll
    for (int i = 0; i < kLcSlots; ++i) {
        const int row = srow0 + i * kLcRowGroups;
        const int dz = row / kLcSH, hy = row - dz * kLcSH;
        const int gd = d0 - 1 + dz, gh = h0 - 1 + hy, gw = w0 - 4 + 4 * svec;
        const bool ok = stager && row < kLcRows && gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W;   // W % 4 == 0
        goff[i] = ok ? (gd * H + gh) * W + gw : -1;
    }
    auto issue = [&](int c, int stage) {
        if (c < C && stager) {
            const float* src = inb + (size_t)c * chan;
            float* dst = lc_smem + stage * kLcStage + srow0 * kLcSW + lc_skew(4 * svec);
#pragma unroll
            for (int i = 0; i < kLcSlots; ++i) {
                if (srow0 + i * kLcRowGroups < kLcRows) {
                    const bool ok = goff[i] >= 0;
                    __pipeline_memcpy_async(dst + i * kLcRowGroups * kLcSW, ok ? src + goff[i] : src, 16, ok ? 0 : 16);   // zero fill outside
                }
            }
        }
        __pipeline_commit();
    };
    if (STAGES == 2) issue(0, 0);
    // thread geometry: 4 d-groups x (8 / TH) h groups x 8 w-groups
    const int tw = tid & 7, th = ((tid >> 3) & (8 / TH - 1)) * TH, td = tid / (64 / TH);
    float acc[4][TH][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int r = 0; r < TH; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[a][r][i] = 0.f;
    // the thread's ten inputs of a row are tile columns 8tw+3 .. 8tw+12: two aligned vectors (8tw+4, 8tw+8), the
    // neighbours' edge elements by shuffle, and for the two lanes at the ends of the 64-column strip one scalar
    const int toff = ((4 * td) * kLcSH + th) * kLcSW;
    const int c0 = lc_skew(8 * tw + 4), c1 = lc_skew(8 * tw + 8);
    const bool edge = tw == 0 || tw == 7;
    const int ce = tw == 0 ? lc_skew(3) : lc_skew(68);

    for (int c = 0; c < C; ++c) {
        if (STAGES == 2) {
            issue(c + 1, (c + 1) & 1);
            __pipeline_wait_prior(1);
        } else {
            issue(c, 0);                                  // single stage: the other CTAs of the SM cover the copy
            __pipeline_wait_prior(0);
        }
        __syncthreads();
        const float* st = lc_smem + (STAGES == 2 ? (c & 1) : 0) * kLcStage + toff;
        float wr[3][3][3];
#pragma unroll
        for (int kd = 0; kd < 3; ++kd)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) wr[kd][kh][kw] = c_lcw[wslot][c * 27 + (kd * 3 + kh) * 3 + kw];
#pragma unroll
        for (int dz = 0; dz < 6; ++dz) {
#pragma unroll
            for (int hy = 0; hy < TH + 2; ++hy) {          // input row th + hy feeds output rows hy - kh
                const float* p = st + (dz * kLcSH + hy) * kLcSW;
                float v[10];
                const float4 m0 = *reinterpret_cast<const float4*>(p + c0);
                const float4 m1 = *reinterpret_cast<const float4*>(p + c1);
                const float e = edge ? p[ce] : 0.f;
                const float lft = __shfl_up_sync(0xffffffffu, m1.w, 1, 8);     // column 8tw+3 = the left neighbour's last
                const float rgt = __shfl_down_sync(0xffffffffu, m0.x, 1, 8);   // column 8tw+12 = the right neighbour's first
                v[0] = tw == 0 ? e : lft;
                v[1] = m0.x; v[2] = m0.y; v[3] = m0.z; v[4] = m0.w;
                v[5] = m1.x; v[6] = m1.y; v[7] = m1.z; v[8] = m1.w;
                v[9] = tw == 7 ? e : rgt;
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    const int od = dz - kd;                  // output d (tile-local, within the thread's 4) fed by this row
                    if (od < 0 || od > 3) continue;
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const int oh = hy - kh;              // output row (within the thread's TH)
                        if (oh < 0 || oh >= TH) continue;
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            acc[od][oh][i] = __fmaf_rn(wr[kd][kh][2], v[i + 2], __fmaf_rn(wr[kd][kh][1], v[i + 1], __fmaf_rn(wr[kd][kh][0], v[i], acc[od][oh][i])));
                    }
                }
            }
        }
        __syncthreads();
    }

    const int wo = w0 + 8 * tw;
#pragma unroll
    for (int r = 0; r < TH; ++r) {
        const int h = h0 + th + r;
        if (h < H && wo < W) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int d = d0 + 4 * td + a;
                if (d < D) {
                    float* o = out + (((size_t)b * D + d) * H + h) * W + wo;
                    reinterpret_cast<float4*>(o)[0] = make_float4(acc[a][r][0], acc[a][r][1], acc[a][r][2], acc[a][r][3]);
                    if (wo + 4 < W) reinterpret_cast<float4*>(o)[1] = make_float4(acc[a][r][4], acc[a][r][5], acc[a][r][6], acc[a][r][7]);
                }
            }
        }
    }
}

int conv3d_c1_fwd(const float* in, const float* w, float* out, int B, int C, int D, int H, int W, cudaStream_t st) {
    if (!in || !w || !out) return fail(RAG_E_NULL, "conv3d_c1_fwd: null pointer");
    if (B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return fail(RAG_E_SHAPE, "conv3d_c1_fwd: non-positive dimension");
    if (W % 4 != 0) return fail(RAG_E_SHAPE, "conv3d_c1_fwd: W=%d must be a multiple of 4", W);
    if ((size_t)D * H * W >= ((size_t)1 << 31)) return fail(RAG_E_SHAPE, "conv3d_c1_fwd: D*H*W must be < 2^31");
    if (C > kLcMaxC) return fail(RAG_E_SHAPE, "conv3d_c1_fwd: C=%d exceeds the constant-memory weight table (%d)", C, kLcMaxC);
    if (!aligned(in, 16) || !aligned(out, 16) || !aligned(w, 4)) return fail(RAG_E_ALIGN, "conv3d_c1_fwd: in/out must be 16-byte aligned");
    const int n_dt = (D + kLcDT - 1) / kLcDT;
    if ((long long)B * n_dt > 65535) return fail(RAG_E_SHAPE, "conv3d_c1_fwd: B*ceil(D/16) must be <= 65535");
    // one 57.6 KB stage: 3 CTAs per SM cover each other's copies (0.565 ms at B=8 480x960); a double-buffered
    // CTA fits only once per SM and is slower (0.735 ms)
    constexpr int stages = 1;
    const size_t smem = (size_t)stages * kLcStage * sizeof(float);
    // two output rows per thread (128-thread CTAs, 18 instead of 12 FMAs per shared-memory load) pay off once
    // the grid is several waves deep: 0.501 vs 0.554 ms at B=8 480x960, 0.126 vs 0.122 ms at B=4 288x576
    const long long ctas = (long long)((W + kLcWT - 1) / kLcWT) * ((H + kLcHT - 1) / kLcHT) * B * n_dt;
    const int TH = ctas >= 4LL * 3 * num_sms() ? 2 : 1;
    auto kern = TH == 2 ? conv3d_c1_kernel<stages, 2> : conv3d_c1_kernel<stages, 1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "conv3d_c1_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    e = cudaStreamIsCapturing(st, &cap);
    if (e != cudaSuccess) return fail((int)e, "conv3d_c1_fwd: cudaStreamIsCapturing: %s", cudaGetErrorString(e));
    if (cap != cudaStreamCaptureStatusNone)
        return fail(RAG_E_CAPTURE, "conv3d_c1_fwd: the stream is capturing; the constant-memory weight slots are ordered with events and cannot be captured");
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return fail(RAG_E_SHAPE, "conv3d_c1_fwd: unsupported device index");
    LcRing& ring = g_lc_ring[dev];
    std::lock_guard<std::mutex> lock(ring.mu);
    const int wslot = (int)(ring.next++ % (unsigned)kLcSlotsW);
    if (ring.have[wslot]) {
        e = cudaStreamWaitEvent(st, ring.ev[wslot], 0);      // the previous reader of this slot has finished
        if (e != cudaSuccess) return fail((int)e, "conv3d_c1_fwd: cudaStreamWaitEvent: %s", cudaGetErrorString(e));
    } else {
        e = cudaEventCreateWithFlags(&ring.ev[wslot], cudaEventDisableTiming);
        if (e != cudaSuccess) return fail((int)e, "conv3d_c1_fwd: cudaEventCreate: %s", cudaGetErrorString(e));
        ring.have[wslot] = true;
    }
    e = cudaMemcpyToSymbolAsync(c_lcw, w, (size_t)C * 27 * sizeof(float), (size_t)wslot * kLcMaxC * 27 * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return fail((int)e, "conv3d_c1_fwd: weight copy: %s", cudaGetErrorString(e));
    dim3 grid((W + kLcWT - 1) / kLcWT, (H + kLcHT - 1) / kLcHT, B * n_dt);
    kern<<<grid, 256 / TH, smem, st>>>(in, w, out, C, D, H, W, n_dt, wslot);
    if (int rc = check_launch("conv3d_c1_fwd")) return rc;
    e = cudaEventRecord(ring.ev[wslot], st);
    if (e != cudaSuccess) return fail((int)e, "conv3d_c1_fwd: cudaEventRecord: %s", cudaGetErrorString(e));
    return RAG_OK;
}

}  // namespace rag
