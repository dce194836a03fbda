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
#include <type_traits>

#include "common.cuh"

namespace rag {

constexpr int kLcDT = 16, kLcHT = 8, kLcWT = 64;            // output tile
constexpr int kLcSD = kLcDT + 2, kLcSH = kLcHT + 2;        // staged tile: d in [d0-1, d0+17), h in [h0-1, h0+9)
constexpr int kLcRowVecs = (kLcWT + 8) / 4;                 // ... w in [w0-4, w0+68): 18 vectors per row
// Shared-memory rows are SKEWED: 4 floats of padding after every 32, so that the eight 8-float segments of
// a row start in distinct bank groups and the threads' LDS.128 are conflict-free (unskewed: 2-way).
constexpr int kLcSW = kLcWT + 8 + 8;                        // row stride (80 floats)
// (columns 64-67, the ninth vector a thread row reads, sit IN the first gap: after the plain skew they would share bank
// group 2 with columns 8-11, a 2-way conflict on every second LDS.128)
__host__ __device__ constexpr int lc_skew(int col) { return col < 64 ? col + 4 * (col >> 5) : (col < 68 ? col - 32 : col + 8); }
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
__constant__ __align__(16) float c_lcw[kLcSlotsW][kLcMaxC * 28];   // 27 weights per channel, padded to 28 (16-byte rows)
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
                for (int kw = 0; kw < 3; ++kw) wr[kd][kh][kw] = c_lcw[wslot][c * 28 + (kd * 3 + kh) * 3 + kw];
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

// The march kernel: ONE warp per CTA owns an 8 (h) x 64 (w) column of DT output planes and walks the input planes d0-1 ..
// d0+DT in order; a thread keeps THREE rotating accumulator sets (2 h x 8 w each) for the output planes d-1, d, d+1 that the
// current input plane feeds (kd = 2, 1, 0), loops over the channels inside the plane, and stores the oldest set when the
// plane is done.  A unit of work is one channel of one plane (10 x 72 inputs, 432 FFMA per thread); the warp's units stream
// through a private ring of R shared-memory slots filled by cp.async R-1 units ahead -- no block barrier anywhere, one
// __syncwarp per unit.  The inner loop (one unit) is ~9 KB of code run C times in a row, where the tile kernel's unrolled
// channel body is 47 KB that twelve warps walk at different places (no_instruction stalls).
//
// Shared-memory layout: a cp.async lands in shared memory one wavefront per (global 128-byte line, shared 128-byte chunk)
// pair it touches (tools/ldgsts_probe.cu: 4 wavefronts per warp instruction when lines map onto chunks, 12 when the run is
// shifted by 16 bytes), and with the tile kernel's layout those wavefronts -- not issue slots -- are the bound.  So a staged
// row is three chunks that mirror global lines: columns w0..w0+31, w0+32..w0+63, and an edge chunk with w0+64..w0+67 at its
// front and w0-4..w0-1 further back.  The vector pairs of the second chunk are swapped (lm_swz), so that the eight lanes of a
// row, which own columns 8tw..8tw+7, hit eight different bank groups with each of their LDS.128.
constexpr int kLmSW = 96, kLmPlane = kLcSH * kLmSW;           // row stride (floats), floats per staged plane
__host__ __device__ constexpr int lm_swz(int v) { return v < 8 || v > 15 ? v : v ^ 1; }
// the layout's bank claim, checked at compile time: the eight lanes of a row (tw = 0..7) read three vectors each -- left
// neighbour (the edge chunk's vector 6 for lane 0), own first, own second -- and every one of the three warp-wide loads must
// touch eight different 16-byte bank groups
constexpr int lm_vec_left(int tw) { return tw == 0 ? 16 + 6 : lm_swz(2 * tw - 1); }
constexpr bool lm_groups_distinct(int which) {
    int seen = 0;
    for (int tw = 0; tw < 8; ++tw) {
        const int v = which == 0 ? lm_vec_left(tw) : (which == 1 ? lm_swz(2 * tw) : lm_swz(2 * tw + 1));
        const int grp = v & 7;
        if (seen & (1 << grp)) return false;
        seen |= 1 << grp;
    }
    return true;
}
static_assert(lm_groups_distinct(0) && lm_groups_distinct(1) && lm_groups_distinct(2), "march kernel: LDS.128 bank conflict in the row layout");
__device__ __forceinline__ void lm_cp16(float* smem, const float* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}

template <int R>
__global__ void __launch_bounds__(32, 12)
conv3d_c1_march_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int D, int H, int W, int DT, int n_dt, int wslot) {
    extern __shared__ __align__(128) float lc_smem[];
    const int lane = threadIdx.x;
    const int b = blockIdx.z / n_dt, dt = blockIdx.z - b * n_dt;
    const int d0 = dt * DT, h0 = blockIdx.y * kLcHT, w0 = blockIdx.x * kLcWT;
    const int dend = min(d0 + DT, D), nplanes = dend - d0 + 2;
    const size_t chan = (size_t)D * H * W, plane = (size_t)H * W;
    const float* inb = in + (size_t)b * C * chan;

    // Copy slots of a lane: i < 5 -> the two full chunks of rows 2i and 2i+1 (16 lanes per row: shared and global offsets
    // are a per-lane base plus i times a constant); slot 5 -> the two edge vectors of the ten rows (lanes 0-19).  Vectors
    // outside the image are zeroed once in every ring slot and never copied; a plane outside the volume is not copied at
    // all (its units are skipped).
    const int mrow = lane >> 4, mvec = lane & 15, erow = lane >> 1, eright = lane & 1;
    int m_s = mrow * kLmSW + 4 * lm_swz(mvec), e_s = erow * kLmSW + 64 + (eright ? 0 : 24);
    int m_g = (h0 - 1 + mrow) * W + w0 + 4 * mvec, e_g = (h0 - 1 + erow) * W + w0 + (eright ? 64 : -4);
    unsigned okmask = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int row = i < 5 ? 2 * i + mrow : erow;
        const int gh = h0 - 1 + row, gw = i < 5 ? w0 + 4 * mvec : w0 + (eright ? 64 : -4);
        const bool mine = i < 5 || lane < 20;
        const bool ok = mine && gh >= 0 && gh < H && gw >= 0 && gw < W;                  // W % 4 == 0
        okmask |= ok ? 1u << i : 0u;
        if (mine && !ok) {
#pragma unroll
            for (int q = 0; q < R; ++q)
                *reinterpret_cast<float4*>(lc_smem + q * kLmPlane + (i < 5 ? m_s + i * 2 * kLmSW : e_s)) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncwarp();
    asm volatile("" : "+r"(m_s), "+r"(e_s), "+r"(m_g), "+r"(e_g));   // keep them in registers (the compiler would re-derive them per unit)
    const bool ok0 = okmask & 1u, ok1 = okmask & 2u, ok2 = okmask & 4u, ok3 = okmask & 8u, ok4 = okmask & 16u, ok5 = okmask & 32u;
    const int W2 = 2 * W;
    // producer state: the next unit to copy is channel pc of plane pj (d = d0 - 1 + pj), into ring slot ps
    int pj = 0, pc = 0, ps = 0;
    const float* psrc = inb + (ptrdiff_t)(d0 - 1) * (ptrdiff_t)plane;       // points before the volume while the plane is outside (not read)
    // a unit's copy is issued in three pairs of vectors, spread over the consumer's row blocks (six LDGSTS back to back fill
    // the shared-memory queue the row loads wait in)
    bool cv = false;
    const float* cg = nullptr;
    float* cdst = nullptr;
    auto copy_prep = [&]() {
        const int gd = d0 - 1 + pj;
        cv = pj < nplanes && gd >= 0 && gd < D;
        cg = psrc + m_g;
        cdst = lc_smem + ps * kLmPlane + m_s;
    };
    auto copy_pair = [&](int k) {
        if (!cv) return;
        if (k == 0) {
            if (ok0) lm_cp16(cdst, cg);
            if (ok1) lm_cp16(cdst + 2 * kLmSW, cg + W2);
        } else if (k == 1) {
            if (ok2) lm_cp16(cdst + 4 * kLmSW, cg + 2 * W2);
            if (ok3) lm_cp16(cdst + 6 * kLmSW, cg + 3 * W2);
        } else {
            if (ok4) lm_cp16(cdst + 8 * kLmSW, cg + 4 * W2);
            if (ok5) lm_cp16(lc_smem + ps * kLmPlane + e_s, psrc + e_g);
        }
    };
    auto copy_done = [&]() {
        if (pj < nplanes) {
            psrc += chan;
            if (++pc == C) { pc = 0; ++pj; psrc += (ptrdiff_t)plane - (ptrdiff_t)C * (ptrdiff_t)chan; }
            if (++ps == R) ps = 0;
        }
        __pipeline_commit();
    };
    auto copy_next = [&]() { copy_prep(); copy_pair(0); copy_pair(1); copy_pair(2); copy_done(); };
#pragma unroll
    for (int i = 0; i < R - 1; ++i) copy_next();

    // thread geometry: 8 w groups x 4 h groups; a thread owns columns 8tw..8tw+7 of rows th, th+1 and reads, per input row,
    // the vector left of them (only its last element is used), its own two vectors and the element right of them -- four
    // loads at per-lane offsets, no shuffles and no selects (the out-of-tile neighbours of lanes 0 and 7 are the edge chunk)
    const int tw = lane & 7, th = ((lane >> 3) & 3) * 2;
    int c0 = th * kLmSW + 4 * lm_swz(2 * tw), c1 = th * kLmSW + 4 * lm_swz(2 * tw + 1);
    int cl = th * kLmSW + (tw == 0 ? 64 + 24 : 4 * lm_swz(2 * tw - 1)), cr = th * kLmSW + (tw == 7 ? 64 : 4 * lm_swz(2 * tw + 2));
    asm volatile("" : "+r"(c0), "+r"(c1), "+r"(cl), "+r"(cr));
    int cs = 0;                                                   // consumer ring slot

    float a2[2][8], a1[2][8], a0[2][8];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int i = 0; i < 8; ++i) a2[q][i] = a1[q][i] = a0[q][i] = 0.f;

    // input plane j feeds output planes d0+j-2 (kd = 2, set a2, complete afterwards), d0+j-1 (kd = 1, a1), d0+j (kd = 0, a0);
    // MASK (bit kd) says which of the three lie inside the tile -- a compile-time constant, so a unit is one basic block
    auto do_plane = [&](auto mask_tag, bool live) {
        constexpr int MASK = decltype(mask_tag)::value;
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
            __pipeline_wait_prior(R - 2);
            __syncwarp();                                         // the unit has landed for all lanes; all lanes are past the previous one
            const float* st = lc_smem + cs * kLmPlane;
            if (++cs == R) cs = 0;
            if (!live) { copy_next(); continue; }                 // a plane outside the volume is all zeros: nothing to add
            copy_prep();
            float wr[3][3][3];
#pragma unroll
            for (int kd = 0; kd < 3; ++kd)
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) wr[kd][kh][kw] = c_lcw[wslot][c * 28 + (kd * 3 + kh) * 3 + kw];
#pragma unroll
            for (int hy = 0; hy < 4; ++hy) {
                const float* p = st + hy * kLmSW;
                const float4 ml = *reinterpret_cast<const float4*>(p + cl);
                const float4 m0 = *reinterpret_cast<const float4*>(p + c0);
                const float4 m1 = *reinterpret_cast<const float4*>(p + c1);
                float v[10];
                v[0] = ml.w;
                v[1] = m0.x; v[2] = m0.y; v[3] = m0.z; v[4] = m0.w;
                v[5] = m1.x; v[6] = m1.y; v[7] = m1.z; v[8] = m1.w;
                v[9] = p[cr];
                if (hy < 3) copy_pair(hy);
                else copy_done();
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int oh = hy - kh;
                    if (oh < 0 || oh >= 2) continue;
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd) {
                        if (!(MASK & (1 << kd))) continue;
                        float (&a)[2][8] = kd == 2 ? a2 : (kd == 1 ? a1 : a0);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            a[oh][i] = __fmaf_rn(wr[kd][kh][2], v[i + 2], __fmaf_rn(wr[kd][kh][1], v[i + 1], __fmaf_rn(wr[kd][kh][0], v[i], a[oh][i])));
                    }
                }
            }
        }
    };
    const int n = dend - d0;
    const int wo = w0 + 8 * tw;
#pragma unroll 1
    for (int j = 0; j < nplanes; ++j) {
        const int gd = d0 - 1 + j;
        const bool live = gd >= 0 && gd < D;
        const int mask = (j >= 2 ? 4 : 0) | (j >= 1 && j <= n ? 2 : 0) | (j < n ? 1 : 0);
        switch (mask) {
            case 7: do_plane(std::integral_constant<int, 7>{}, live); break;
            case 1: do_plane(std::integral_constant<int, 1>{}, live); break;
            case 3: do_plane(std::integral_constant<int, 3>{}, live); break;
            case 6: do_plane(std::integral_constant<int, 6>{}, live); break;
            case 4: do_plane(std::integral_constant<int, 4>{}, live); break;
            default: do_plane(std::integral_constant<int, 2>{}, live); break;
        }
        if (j >= 2) {                                             // output plane d0 + j - 2 is complete
            const int d = d0 + j - 2;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int h = h0 + th + q;
                if (h < H && wo < W) {
                    float* o = out + (((size_t)b * D + d) * H + h) * W + wo;
                    reinterpret_cast<float4*>(o)[0] = make_float4(a2[q][0], a2[q][1], a2[q][2], a2[q][3]);
                    if (wo + 4 < W) reinterpret_cast<float4*>(o)[1] = make_float4(a2[q][4], a2[q][5], a2[q][6], a2[q][7]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int i = 0; i < 8; ++i) { a2[q][i] = a1[q][i]; a1[q][i] = a0[q][i]; a0[q][i] = 0.f; }
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
    void* ctab = nullptr;
    e = cudaGetSymbolAddress(&ctab, c_lcw);
    if (e != cudaSuccess) return fail((int)e, "conv3d_c1_fwd: cudaGetSymbolAddress: %s", cudaGetErrorString(e));
    e = cudaMemcpy2DAsync(static_cast<float*>(ctab) + (size_t)wslot * kLcMaxC * 28, 28 * sizeof(float), w, 27 * sizeof(float), 27 * sizeof(float), C,
                          cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return fail((int)e, "conv3d_c1_fwd: weight copy: %s", cudaGetErrorString(e));
    // The march kernel (one warp per CTA, 8 output planes each) wins wherever its grid fills the machine: 0.366 vs 0.494 ms
    // at B=8 480x960, 0.082 vs 0.111 ms at B=4 288x576; the crossover is at 600-800 warps (one pair at 480x960, 800 warps: 0.081 vs
    // 0.085 ms; two pairs at 288x576, 576 warps: 0.064 vs 0.062 ms), below it the tile kernel's 4-warp CTAs are faster.
    constexpr int kMarchDT = 8, kMarchRing = 3;
    const int n_dtm = (D + kMarchDT - 1) / kMarchDT;
    const long long warps = (long long)((W + kLcWT - 1) / kLcWT) * ((H + kLcHT - 1) / kLcHT) * B * n_dtm;
    if (warps >= 5LL * num_sms() && (long long)B * n_dtm <= 65535) {
        auto mk = conv3d_c1_march_kernel<kMarchRing>;
        e = cudaFuncSetAttribute(mk, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return fail((int)e, "conv3d_c1_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        dim3 grid((W + kLcWT - 1) / kLcWT, (H + kLcHT - 1) / kLcHT, B * n_dtm);
        mk<<<grid, 32, kMarchRing * kLmPlane * sizeof(float), st>>>(in, out, C, D, H, W, kMarchDT, n_dtm, wslot);
    } else {
        dim3 grid((W + kLcWT - 1) / kLcWT, (H + kLcHT - 1) / kLcHT, B * n_dt);
        kern<<<grid, 256 / TH, smem, st>>>(in, w, out, C, D, H, W, n_dt, wslot);
    }
    if (int rc = check_launch("conv3d_c1_fwd")) return rc;
    e = cudaEventRecord(ring.ev[wslot], st);
    if (e != cudaSuccess) return fail((int)e, "conv3d_c1_fwd: cudaEventRecord: %s", cudaGetErrorString(e));
    return RAG_OK;
}

}  // namespace rag
