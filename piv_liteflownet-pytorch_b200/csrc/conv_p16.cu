// Convolution (+ bias + LeakyReLU) on P16 activations: implicit GEMM on the 5th-generation tensor cores with NO operand
// preparation inside the kernel.
//
// Split-operand arithmetic (p16.cuh): a = a_hi + 2^-11 a_lo, w = W_hi + W_lo (weights pre-scaled by a per-layer power of two),
//     D = a_hi*W_hi                               one kind::f16 MMA, K = 16
//       + [a_lo8 | a_hi8] * [W 2^-11 ; W_lo]      ONE kind::f8f6f4 MMA (A e5m2, B e4m3), K = 32: both correction products
// into one fp32 accumulator in TMEM: two MMA slots per 16 input channels (three for an fp16 split of the corrections).
// The activations arrive from HBM already in operand form: per 32-channel chunk ONE 4-D TMA box
//     [16*NT + KH - 1 rows][8 + KW - 1 pixels][128 bytes = hi0 | c0 | hi1 | c1]         (128B swizzle, zero fill outside)
// lands in shared memory as the MMA-ready K-major halo tile; every filter tap's A operand is a shifted window into it and
// the four K steps of a tap are 32-byte advances of the descriptor start address.  Compared with conv_tc.cu's f16c modes
// the 8 operand-split warps, the raw-tile slot and ~15-23 % of the shared-memory traffic are gone, and the freed
// warps double the epilogue: 16 warps (4 per TMEM lane quarter) turn accumulators into P16 (or fp32) rows.
//
//   warp 0        TMA producer of the activation tiles          warp 3   producer of the weight ring (one bulk copy / stage)
//   warps 1, 2    MMA issuers (every second stacked tile each)  warps 4..19   epilogue (12 of them gather when a backwarp is fused)
//
// Replaces torch.nn.Conv2d (+LeakyReLU(0.1)) of src/models.py:77-106 (NetC), :124 (NetC_ext), :154-163 (conv_M),
// :197-207 (conv_S), :229-272 (moduleFeat, conv_R, conv_dist_R).
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "p16.cuh"
#include "tc_ptx.cuh"

namespace {
using namespace tcptx;

constexpr int HT_W = 8, HT_H = 16;       // accumulator tile: 16 rows of 8 pixels = 128 GEMM rows
constexpr int MAX_A = 3, MAX_B = 12;
constexpr int P16_THREADS = 640;
constexpr int EPI_WARP0 = 4, EPI_WARPS = 16;
constexpr int SMEM_BUDGET = 226 * 1024;  // 227 KB opt-in minus the static part (padded to 1 KB by the 1024-byte alignment)

enum { OUT_P16 = 0, OUT_F32 = 1, OUT_PLANES = 2, OUT_TAIL = 3 };

// OUT_TAIL: the K*K output channels are the Regularization distances (src/models.py:279-300); the epilogue turns them into the
// regularised flow while they are still in TMEM (negative square, softmax over the K*K neighbours, weighted unfold of the flow,
// the two 1x1 ScaleX / ScaleY convolutions and the division) -- the distance tensor never exists in HBM.
struct TailArgs {
    const float2* flow;      // dense [N,H,W,2] flow that is unfolded (the Subpixel output)
    const float *wx, *wy;    // moduleScaleX / moduleScaleY weights [K*K]
    const float *bx, *by;    // their biases [1]
    float2* out;             // dense [N,H,W,2] regularised flow
    float* nchw;             // optional [N,2,H,W] copy scaled by `scale` (the network output), or NULL
    float scale;
    int K;
};

struct ConvP16Args {
    const float* bias;
    uint32_t* y;             // output words (P16 words or fp32 bits)
    int y_ld;                // output pixel pitch in words
    int N, H, W;             // OUTPUT size
    int Cw;                  // input channel words per pixel that exist (16 * groups)
    int Cout, CoutP;
    int KH, KW;
    int tiles_x, tiles_y, total;
    int lrelu, out_fmt;
    int quad;                // rows 16-byte aligned and W % 4 == 0: quad-transposed 64-byte runs
    int cout_st;             // OUT_F32: channels that may be stored (Cout, or Cout rounded up to 4 when the rows are that wide)
    long long planar;        // OUT_PLANES: floats per channel-pair plane
    int NT, nA, nB, tps, nsets;   // nA = nT + nG activation slots
    int tap_off;             // fused backwarp: byte offset of the tap table (24 B per halo-tile pixel) in dynamic shared memory
    int nT, nG;              // slots filled by TMA / by the gather warps (fused backwarp): every slot has ONE producer, so each
                             // producer sees every phase of the barriers it waits on (mbarrier parity cannot tell 0 from 2)
    int s2, cpp;             // stride-2 restatement over the four input parities (see conv_tc.cu): cpp chunks per parity
    int x_shift;
    const uint8_t* w_img;    // ring-stage image of the fp16 weights (pivlfn.model.stage_image)
    int* range_flag;         // raised when an OUT_P16 result is not finite in fp16 (|x| >= 65520 or NaN); may be NULL
    // fused backwarp (src/models.py:20-35, the Subpixel consumer :209-217): the wnc 32-channel chunks [wc0, wc0 + wnc) of the GEMM K
    // range are NOT in the input buffer; they are backwarp(wsrc, wscale * wflow), gathered, blended, split into fp16 pairs and
    // written into the swizzled activation slot by 12 of the 16 epilogue warps.  The warped features never exist in HBM.
    const uint8_t* wsrc;     // NHWC source of the warp: fp32 (wsrc_p16 = 0) or P16, pixel pitch wsrc_ld words
    const float2* wflow;     // dense [N,H,W,2]
    float wscale;
    int wsrc_ld, wsrc_p16, wc0, wnc;
    int one_issuer;          // a single thread issues the MMAs of both stacked tiles (experiment switch)
    TailArgs tl;             // OUT_TAIL only
};

__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// OUT_TAIL epilogue of one output pixel (same arithmetic, in the same order, as reg_tail_pixel in misc.cu: the fused and the
// two-kernel forms agree to rounding): pass 1 finds min_k d_k^2 over the K*K accumulator columns, pass 2 re-reads them from
// TMEM (cheaper than 49 live registers next to the 32 staging ones) and accumulates the softmax-weighted neighbour flows.
// Fully unrolled over the K*K neighbours so that the K*K flow loads are independent of each other.
// tcol = TMEM address of this thread's accumulator row (CoutP columns, scaled by the layer's weight scale 1 / osc); sw = [wx 64 | wy 64] in shared memory.
template <int K>
__device__ __forceinline__ void tail_pixel(const ConvP16Args& a, uint32_t tcol, float osc, const float* bias_s, const float* sw, int n, int x, int y) {
    constexpr float LOG2E = 1.4426950408889634f;
    constexpr int KK = K * K, P = K / 2, NCG = (KK + 15) / 16;
    uint32_t v[16];
    float mn = INFINITY;
#pragma unroll
    for (int cg = 0; cg < NCG; ++cg) {
        tmem_ld16_nowait(tcol + (uint32_t)(cg * 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (cg * 16 + j < KK) {
                const float d = fmaf(__uint_as_float(v[j]), osc, bias_s[cg * 16 + j]);
                mn = fminf(mn, __fmul_rn(d, d));
            }
        }
    }
    unsigned cm = 0, rm = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        cm |= (unsigned)(x + j - P >= 0 && x + j - P < a.W) << j;
        rm |= (unsigned)(y + j - P >= 0 && y + j - P < a.H) << j;
    }
    const float2* c0 = a.tl.flow + ((long long)n * a.H + (y - P)) * a.W + (x - P);
    float sum = 0.f, au = 0.f, av = 0.f;
#pragma unroll
    for (int cg = 0; cg < NCG; ++cg) {
        tmem_ld16_nowait(tcol + (uint32_t)(cg * 16), v);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            // the neighbour flows of this half column group: 8 independent loads in flight (the first batch ahead of the TMEM wait)
            float2 f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = cg * 16 + hh * 8 + j;
                f[j] = make_float2(0.f, 0.f);
                if (k < KK) {
                    const int ky = k / K, kx = k - ky * K;
                    if (((rm >> ky) & 1u) && ((cm >> kx) & 1u)) f[j] = __ldg(c0 + (long long)ky * a.W + kx);
                }
            }
            if (hh == 0) tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = cg * 16 + hh * 8 + j;
                if (k < KK) {
                    const float d = fmaf(__uint_as_float(v[hh * 8 + j]), osc, bias_s[k]);
                    const float e = ex2_ftz((mn - __fmul_rn(d, d)) * LOG2E);   // d*d rounded on its own, as in pass 1
                    sum += e;
                    au = fmaf(sw[k], e * f[j].x, au);
                    av = fmaf(sw[64 + k], e * f[j].y, av);
                }
            }
        }
    }
    if (x < a.W && y < a.H) {
        const float r = 1.f / sum;
        const float fu = (au + __ldg(a.tl.bx)) * r, fv = (av + __ldg(a.tl.by)) * r;
        const long long HW = (long long)a.H * a.W, q = (long long)y * a.W + x;
        a.tl.out[n * HW + q] = make_float2(fu, fv);
        if (a.tl.nchw) {
            a.tl.nchw[(n * 2LL + 0) * HW + q] = fu * a.tl.scale;
            a.tl.nchw[(n * 2LL + 1) * HW + q] = fv * a.tl.scale;
        }
    }
}

__device__ __forceinline__ bool s2_tap_used(int t, int par) {
    return ((t >> 1) >= 1 - (par >> 1)) && ((t & 1) >= 1 - (par & 1));
}

// Order in which the K chunks are processed when a backwarp is fused in (wc0 == wnc: the concat is [f1 | warp(f2) | rest]): TMA
// and gathered chunks alternate, so that the single gather slot is refilled while the MMAs of a TMA chunk run.
__device__ __forceinline__ int chunk_at(int i, int wc0, int wnc) {
    if (wnc && wc0 == wnc && i < 2 * wnc) return (i & 1) ? wc0 + (i >> 1) : (i >> 1);
    return i;
}

__device__ __forceinline__ void st_global_v4(uint32_t* p, const uint4& v) {
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(P16_THREADS, 1)
conv_p16_kernel(const __grid_constant__ CUtensorMap tmA, const ConvP16Args a) {
    static_assert(MODE == 6, "one product scheme: f16 main + fp8 corrections, single accumulator");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int pitch = HT_W + a.KW - 1;
    const int halo_rows = HT_H * a.NT + a.KH - 1;
    const int halo_bytes = halo_rows * pitch * 128;
    const int slot_bytes = (halo_bytes + 1023) & ~1023;
    const int part_bytes = a.CoutP * 64;                        // one fp16 weight tile: CoutP rows of 32 channels
    const int b_stage = 2 * part_bytes;                         // per tap: [W_hi f16 | (W 2^-11 ; W_lo) e4m3]
    uint8_t* smemB = smem + (size_t)a.nA * slot_bytes;
    __shared__ __align__(8) uint64_t a_full[MAX_A], a_free[MAX_A], b_full[MAX_B], b_empty[MAX_B], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[128];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunk = a.s2 ? 4 * a.cpp : (a.Cw + 31) / 32 + a.wnc;              // GEMM K range in 32-channel chunks
    const int ntaps = a.KH * a.KW;
    const int G = gridDim.x;
    const int n_iss = (a.NT >= 2 && !a.one_issuer) ? 2 : 1;
    const int tile_cols = a.CoutP;
    const int set_cols = a.NT * tile_cols;
    const uint32_t ncols = tmem_cols_for(a.nsets * set_cols);
    const int negr = a.wnc ? 1 : 4;                 // epilogue warp groups (of 4 warps); the other 12 warps gather when a backwarp is fused
                                                    // (that layer is MMA-bound: 4 warps drain its accumulators with time to spare)

    if (threadIdx.x >= 128 && threadIdx.x < 256) {
        const int i = threadIdx.x - 128;
        bias_s[i] = (a.bias && i < a.Cout) ? a.bias[i] : 0.f;
    }
    if (a.out_fmt == OUT_TAIL && threadIdx.x >= 256 && threadIdx.x < 384) {
        const int i = threadIdx.x - 256, k = i & 63;
        reinterpret_cast<float*>(smem + a.tap_off)[i] = k < a.Cout ? __ldg((i < 64 ? a.tl.wx : a.tl.wy) + k) : 0.f;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < a.nA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_free[i], n_iss); }
        for (int i = 0; i < a.nB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], n_iss); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], n_iss); mbar_init(&acc_empty[i], 4 * negr); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    // Programmatic dependent launch (the launch carries cudaLaunchAttributeProgrammaticStreamSerialization): everything above --
    // barrier init, TMEM allocation, bias / tail weights (model parameters, never written by a kernel of the forward) -- ran
    // while the previous kernel of the stream was still draining.  launch_dependents lets the NEXT kernel do the same as soon
    // as this CTA's SM frees up; wait blocks until the previous kernel has completed and its writes are visible.  Every thread
    // waits: every role reads activations (TMA, gathers, the tail's flow loads) or writes buffers the predecessor may read.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ================================ activation tiles ==============================
        if (elect_one()) {
            int slot = 0;                                     // TMA slots 0 .. nT-1, in the order of the TMA chunks
            uint32_t use = 0;
            for (int w = blockIdx.x; w < a.total; w += G) {
                const int tx = w % a.tiles_x, ty = (w / a.tiles_x) % a.tiles_y, n = w / (a.tiles_x * a.tiles_y);
                for (int i = 0; i < nchunk; ++i) {
                    const int c = chunk_at(i, a.wc0, a.wnc);
                    if (a.wnc && c >= a.wc0 && c < a.wc0 + a.wnc) continue;              // filled by the gather warps
                    mbar_wait(&a_free[slot], (use & 1) ^ 1);  // the MMAs of the slot's previous tenant have retired
                    mbar_expect_tx(&a_full[slot], halo_bytes);
                    int c0, c1, c2;
                    if (a.s2) {
                        const int par = c / a.cpp, cc = c - par * a.cpp;
                        c0 = cc * 32; c1 = 2 * (tx * HT_W - 1) + (par & 1); c2 = 2 * (ty * HT_H * a.NT - 1) + (par >> 1);
                    } else {
                        c0 = ((a.wnc && c >= a.wc0 + a.wnc) ? c - a.wnc : c) * 32;       // chunk index inside the input buffer
                        c1 = tx * HT_W + a.x_shift; c2 = ty * HT_H * a.NT - a.KH / 2;
                    }
                    tma_load_4d(smem + (size_t)slot * slot_bytes, &tmA, &a_full[slot], c0, c1, c2, n);
                    if (++slot == a.nT) { slot = 0; ++use; }
                }
            }
        }
    } else if (warp == 3) {
        // ================================ weight ring ====================================
        if (elect_one()) {
            int bs = 0;
            uint32_t bphase = 0;
            for (int w = blockIdx.x; w < a.total; w += G) {
                for (int i = 0; i < nchunk; ++i) {
                    const int c = chunk_at(i, a.wc0, a.wnc);
                    for (int t = 0; t < ntaps; t += a.tps) {
                        if (a.s2 && !s2_tap_used(t, c / a.cpp)) continue;
                        mbar_wait(&b_empty[bs], bphase ^ 1);
                        mbar_expect_tx(&b_full[bs], a.tps * b_stage);
                        bulk_load(smemB + (size_t)bs * a.tps * b_stage, a.w_img + ((size_t)c * ntaps + t) * b_stage,
                                  (uint32_t)(a.tps * b_stage), &b_full[bs]);
                        if (++bs == a.nB) { bs = 0; bphase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ================================ MMA issuers ====================================
        // One thread per stacked tile.  Per tap: one add for the A window, one for the weight tile, the MMAs (every scalar
        // instruction between two tcgen05.mma of a single in-order thread is on the critical path for Cout <= 64).
        const int issuer = warp - 1;
        if (issuer < n_iss && elect_one()) {
            const uint32_t lbo_bits = 1u << 16;
            // A: 128-byte rows, SWIZZLE_128B, 8-row group stride = one halo row (pitch * 128 B); B: 64-byte rows, SWIZZLE_64B
            const uint32_t hiA = (uint32_t)((((uint64_t)((pitch * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61)) >> 32);
            const uint32_t hiB = (uint32_t)((((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61)) >> 32);
            const uint32_t idesc = make_idesc_f16(a.CoutP);
            const uint32_t idesc8 = make_idesc_f8(a.CoutP);
            const uint32_t part16 = (uint32_t)(part_bytes >> 4);
            const uint32_t stage16 = (uint32_t)(b_stage >> 4);
            const uint32_t ring16 = stage16 * (uint32_t)a.tps;
            const uint32_t bbase16 = (smem_u32(smemB) >> 4) | lbo_bits;
            const uint32_t tile16 = (uint32_t)(HT_H * pitch * 8);
            const uint32_t tileA = (uint32_t)issuer * tile16;
            const int ntl = a.NT / n_iss;                                       // stacked tiles fed by this thread: issuer, issuer + n_iss, ..
            const uint32_t row_step = (uint32_t)((pitch - a.KW) * 8);
            int bs = 0;
            uint32_t bphase = 0, bcur = bbase16;
            int wl = 0;
            int tslot = 0, gslot = 0;                                           // next TMA / gather slot
            uint32_t tuse = 0, guse = 0;
            for (int w = blockIdx.x; w < a.total; w += G, ++wl) {
                const int as = wl % a.nsets;
                const uint32_t use = (uint32_t)(wl / a.nsets);
                mbar_wait(&acc_empty[as], (use & 1) ^ 1);     // the epilogue drained this accumulator set
                tc_fence_after();
                const uint32_t t_first = tmem_base + (uint32_t)(as * set_cols) + (uint32_t)(issuer * tile_cols);
                uint32_t acc = 0;
                for (int ci = 0; ci < nchunk; ++ci) {
                    const int c = chunk_at(ci, a.wc0, a.wnc);
                    // second 16-channel K step present?  (only the last chunk of the input buffer can be half empty)
                    const bool two = (a.s2 || (a.wnc && c < a.wc0 + a.wnc)) ? true : (a.Cw - (c - a.wnc) * 32 > 16);
                    const bool gathered = a.wnc && c >= a.wc0 && c < a.wc0 + a.wnc;
                    const int slot = gathered ? a.nT + gslot : tslot;
                    mbar_wait(&a_full[slot], (gathered ? guse : tuse) & 1);
                    tc_fence_after();
                    uint32_t A0 = ((smem_u32(smem + (size_t)slot * slot_bytes) >> 4) | lbo_bits) + tileA;
                    int kx = 0;
                    for (int t = 0; t < ntaps;) {
                        if (a.s2 && !s2_tap_used(t, c / a.cpp)) {
                            ++t; A0 += 8;
                            if (++kx == a.KW) { kx = 0; A0 += row_step; }
                            continue;
                        }
                        mbar_wait(&b_full[bs], bphase);
                        tc_fence_after();
                        uint32_t b = bcur;
                        for (int sub = 0; sub < a.tps; ++sub, ++t, b += stage16) {
                            uint32_t A = A0, t_main = t_first;
                            for (int i = 0; i < ntl; ++i, A += tile16 * (uint32_t)n_iss, t_main += (uint32_t)(tile_cols * n_iss)) {
                                // per 16-channel K step: a_hi * W_hi (f16, K = 16) and [a_lo8 | a_hi8] * [W 2^-11 ; W_lo] (fp8, K = 32)
                                umma_bf16_lohi(t_main, A, hiA, b, hiB, idesc, acc);
                                umma_f8_lohi(t_main, A + 2, hiA, b + part16, hiB, idesc8, 1);
                                if (two) {
                                    umma_bf16_lohi(t_main, A + 4, hiA, b + 2, hiB, idesc, 1);
                                    umma_f8_lohi(t_main, A + 6, hiA, b + part16 + 2, hiB, idesc8, 1);
                                }
                            }
                            acc = 1;
                            A0 += 8;                                            // next tap of the filter row (one pixel = 128 B)
                            if (++kx == a.KW) { kx = 0; A0 += row_step; }
                        }
                        umma_commit(&b_empty[bs]);
                        if (++bs == a.nB) { bs = 0; bphase ^= 1; bcur = bbase16; } else bcur += ring16;
                    }
                    umma_commit(&a_free[slot]);
                    if (gathered) { if (++gslot == a.nG) { gslot = 0; ++guse; } }
                    else if (++tslot == a.nT) { tslot = 0; ++tuse; }
                }
                umma_commit(&acc_full[as]);
            }
        }
    } else {
        // ================================ epilogue =======================================
        // 16 warps: TMEM lane quarter q = warp % 4 (the hardware's rule), group eg = (warp - 4) / 4 takes the (tile, 16-column
        // group) units eg, eg + 4, ...  A thread owns one pixel: 16 accumulator columns -> 16 output words (fp32 bits, or 8
        // words of f16x2 hi + 8 of lo' = one P16 group), quad-transposed so that every store instruction writes 64 contiguous
        // bytes per pixel.
        const int eg = (warp - EPI_WARP0) >> 2, q = warp & 3;
        if (eg >= negr) {
            // ============================ fused-backwarp gather warps ============================
            // 384 threads; item = (pixel of the halo tile, 8-channel unit): 4 bilinear taps x 2 x 16 bytes gathered, blended in
            // fp32, split into (hi, lo') and stored where TMA + the 128B swizzle would have put them (16-byte unit index XOR
            // the low 3 bits of the 128-byte row index; slots are 1024-byte aligned).
            const int gt = threadIdx.x - (EPI_WARP0 + 4 * negr) * 32;
            constexpr int NG = 384;
            const int npx = halo_rows * pitch;
            uint32_t bad = 0;
            int gslot = 0;
            uint32_t guse = 0;
            const int npx4 = npx * 4;
            float4* const tapw = reinterpret_cast<float4*>(smem + a.tap_off);     // bilinear weights of the halo-tile pixels
            int2* const tapxy = reinterpret_cast<int2*>(tapw + npx);              // top-left tap (x0, y0)
            for (int w = blockIdx.x; w < a.total; w += G) {
                const int tx = w % a.tiles_x, ty = (w / a.tiles_x) % a.tiles_y, n = w / (a.tiles_x * a.tiles_y);
                const int xs = tx * HT_W + a.x_shift, ys = ty * HT_H * a.NT - a.KH / 2;
                const long long img = (long long)n * a.H * a.W;
                // taps of the tile's pixels, once per work item (they do not depend on the channel): one flow fetch per pixel here
                // instead of one per (pixel, 8 channels), and the gathers below no longer wait behind it
                for (int p = gt; p < npx; p += NG) {
                    const int r = p / pitch, cx = p - r * pitch;
                    const int y = ys + r, x = xs + cx;
                    float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
                    int2 xy = make_int2(0, 0);
                    if (y >= 0 && y < a.H && x >= 0 && x < a.W) {
                        const float2 fl = __ldg(a.wflow + img + (long long)y * a.W + x);
                        const BilinearTaps tp = make_taps((float)x + fl.x * a.wscale, (float)y + fl.y * a.wscale, a.H, a.W);
                        wv = make_float4(tp.w00, tp.w01, tp.w10, tp.w11);
                        xy = make_int2(tp.x0, tp.y0);
                    }
                    tapw[p] = wv;
                    tapxy[p] = xy;
                }
                asm volatile("bar.sync 2, 384;" ::: "memory");
                for (int k = 0; k < a.wnc; ++k) {                            // (the gathered chunks are processed in increasing order)
                    const int slot = a.nT + gslot;
                    // ONE warp polls the mbarrier, the other eleven sleep in the named barrier: twelve warps spinning on
                    // try_wait would compete with the MMA issuers for issue slots and with the operand reads for shared memory
                    if (gt < 32) mbar_wait(&a_free[slot], (guse & 1) ^ 1);
                    asm volatile("bar.sync 2, 384;" ::: "memory");
                    uint8_t* const dst = smem + (size_t)slot * slot_bytes;
                    constexpr int UNR = 2;                                   // two items = 16 x 16-byte gathers in flight per thread
                    for (int base = gt; base < npx4; base += UNR * NG) {
                        uint4 va[UNR][4], vb[UNR][4];
#pragma unroll
                        for (int e = 0; e < UNR; ++e) {
                            const int item = base + e * NG;
                            const bool ok = item < npx4;
                            const int p = ok ? item >> 2 : 0, u = item & 3;
                            const float4 wv = tapw[p];
                            const int2 xy = tapxy[p];
                            const float wgt[4] = {wv.x, wv.y, wv.z, wv.w};
                            const int cu = k * 4 + u;                        // 8-channel unit of the warp source
                            const int off = a.wsrc_p16 ? p16::unit_off_bytes(cu) : cu * 32;
                            const int off2 = a.wsrc_p16 ? p16::unit_lo8_bytes(cu) : off + 16;
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                va[e][t] = vb[e][t] = make_uint4(0u, 0u, 0u, 0u);
                                if (ok && wgt[t] != 0.f) {                   // taps outside the frame are never dereferenced
                                    const uint8_t* src = a.wsrc + (img + (long long)(xy.y + (t >> 1)) * a.W + (xy.x + (t & 1))) * (long long)a.wsrc_ld * 4;
                                    va[e][t] = __ldg(reinterpret_cast<const uint4*>(src + off));
                                    if (a.wsrc_p16) {                        // 8 lo8 bytes
                                        const uint2 l8 = __ldg(reinterpret_cast<const uint2*>(src + off2));
                                        vb[e][t].x = l8.x; vb[e][t].y = l8.y;
                                    } else {
                                        vb[e][t] = __ldg(reinterpret_cast<const uint4*>(src + off2));
                                    }
                                }
                            }
                        }
#pragma unroll
                        for (int e = 0; e < UNR; ++e) {
                            const int item = base + e * NG;
                            if (item < npx4) {
                                const int p = item >> 2, u = item & 3;
                                const float4 wv = tapw[p];
                                const float wgt[4] = {wv.x, wv.y, wv.z, wv.w};
                                float v[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = 0.f;
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    float f[8];
                                    if (a.wsrc_p16) p16::decode8(va[e][t], make_uint2(vb[e][t].x, vb[e][t].y), f);
                                    else {
                                        f[0] = __uint_as_float(va[e][t].x); f[1] = __uint_as_float(va[e][t].y); f[2] = __uint_as_float(va[e][t].z); f[3] = __uint_as_float(va[e][t].w);
                                        f[4] = __uint_as_float(vb[e][t].x); f[5] = __uint_as_float(vb[e][t].y); f[6] = __uint_as_float(vb[e][t].z); f[7] = __uint_as_float(vb[e][t].w);
                                    }
#pragma unroll
                                    for (int j = 0; j < 8; ++j) v[j] = fmaf(wgt[t], f[j], v[j]);
                                }
                                uint4 h;
                                uint2 l, g;
                                p16::encode8(v, h, l, g);
                                bad |= p16::nonfinite_bits(h);
                                // 16-byte units of the 128-byte row: group base gb = 4 * (u / 2); hi of this half group at gb + (u & 1),
                                // the lo8 bytes in unit gb + 2 and the hi8 bytes in unit gb + 3, each at byte 8 * (u & 1)
                                const uint32_t gb = (uint32_t)(u >> 1) * 4u, hf = (uint32_t)(u & 1);
                                const uint32_t sw = (uint32_t)(p & 7);
                                uint8_t* rowp = dst + (size_t)p * 128;
                                *reinterpret_cast<uint4*>(rowp + (((gb + hf) ^ sw) << 4)) = h;
                                *reinterpret_cast<uint2*>(rowp + (((gb + 2) ^ sw) << 4) + hf * 8) = l;
                                *reinterpret_cast<uint2*>(rowp + (((gb + 3) ^ sw) << 4) + hf * 8) = g;
                            }
                        }
                    }
                    fence_proxy_async();                      // generic-proxy writes -> visible to the tensor core (async proxy)
                    asm volatile("bar.sync 2, 384;" ::: "memory");
                    if (gt == 0) mbar_arrive(&a_full[slot]);
                    if (++gslot == a.nG) { gslot = 0; ++guse; }
                }
            }
            if (a.range_flag && p16::any_nonfinite(bad)) *a.range_flag = 1;
        } else {
        const int row = q * 32 + lane;
        const int ncg = a.CoutP >> 4;
        const int nunits = a.NT * ncg;
        const int u_i0 = eg / ncg, u_c0 = eg - u_i0 * ncg, u_di = negr / ncg, u_dc = negr - u_di * ncg;
        const float slope = a.lrelu ? PIVLFN_LRELU_SLOPE : 1.f;
        // the weights are packed times a per-layer power of two S (pivlfn.model._pack_f8); the image ends with [1 / S, S, 0, 0]
        const float osc = __ldg(reinterpret_cast<const float*>(a.w_img + (size_t)nchunk * ntaps * b_stage));
        uint32_t bad = 0;
        int wl = 0;
        for (int w = blockIdx.x; w < a.total; w += G, ++wl) {
            const int as = wl % a.nsets;
            const uint32_t use = (uint32_t)(wl / a.nsets);
            const int tx = w % a.tiles_x, ty = (w / a.tiles_x) % a.tiles_y, n = w / (a.tiles_x * a.tiles_y);
            const int x = tx * HT_W + (row & (HT_W - 1));
            mbar_wait(&acc_full[as], use & 1);
            tc_fence_after();
            const uint32_t trow = tmem_base + (uint32_t)(as * set_cols) + ((uint32_t)(q * 32) << 16);
            if (a.out_fmt == OUT_TAIL) {
                {
                    const float* sw = reinterpret_cast<const float*>(smem + a.tap_off);
                    for (int i = eg; i < a.NT; i += negr) {
                        const uint32_t tc = trow + (uint32_t)(i * tile_cols);
                        const int yy = (ty * a.NT + i) * HT_H + (row >> 3);
                        if (a.tl.K == 7) tail_pixel<7>(a, tc, osc, bias_s, sw, n, x, yy);
                        else if (a.tl.K == 5) tail_pixel<5>(a, tc, osc, bias_s, sw, n, x, yy);
                        else tail_pixel<3>(a, tc, osc, bias_s, sw, n, x, yy);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[as]);
                continue;
            }
            uint32_t v[16];
            // unit = (stacked tile i, 16-column group cg), walked with two counters (a division per unit cost ~40 instructions of
            // the ~400 this loop body takes)
            int i = u_i0, cg = u_c0;
            if (eg < nunits) tmem_ld16_nowait(trow + (uint32_t)(i * tile_cols + cg * 16), v);
            for (int unit = eg; unit < nunits; unit += negr) {
                const int cb = cg * 16;
                int ni = i + u_di, ncg2 = cg + u_dc;                       // the next unit of this warp
                if (ncg2 >= ncg) { ncg2 -= ncg; ++ni; }
                const int i_cur = i;
                tmem_ld_wait();
                float r[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 b4 = *reinterpret_cast<const float4*>(&bias_s[cb + 4 * j]);
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float t = fmaf(__uint_as_float(v[4 * j + k]), osc, bb[k]);
                        r[4 * j + k] = fmaxf(t, t * slope);                 // LeakyReLU(0.1) = max(t, 0.1 t); slope 1: identity
                    }
                }
                // v / u are consumed: the TMEM loads of this warp's next unit hide behind the conversion and the stores
                if (unit + negr < nunits) tmem_ld16_nowait(trow + (uint32_t)(ni * tile_cols + ncg2 * 16), v);
                i = ni; cg = ncg2;
                const int yy = (ty * a.NT + i_cur) * HT_H + (row >> 3);
                const size_t pix = ((size_t)n * a.H + yy) * a.W + x;
                if (a.out_fmt == OUT_PLANES) {
                    if (x < a.W && yy < a.H) {
#pragma unroll
                        for (int j = 0; j < 16; j += 2)
                            if (cb + j < a.Cout)
                                *reinterpret_cast<float2*>(reinterpret_cast<float*>(a.y) + (size_t)((cb + j) >> 1) * a.planar + pix * 2) =
                                    make_float2(r[j], r[j + 1]);
                    }
                    continue;
                }
                uint4 t4[4];
                if (a.out_fmt == OUT_P16) {
                    p16::encode16(r, t4[0], t4[1], t4[2], t4[3]);
                    bad |= p16::nonfinite_bits(t4[0]) | p16::nonfinite_bits(t4[1]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        t4[j] = make_uint4(__float_as_uint(r[4 * j]), __float_as_uint(r[4 * j + 1]), __float_as_uint(r[4 * j + 2]),
                                           __float_as_uint(r[4 * j + 3]));
                }
                const bool full = a.out_fmt == OUT_P16 || cb + 16 <= a.cout_st;
                if (a.quad && full) {
                    // 4x4 transpose of 16-byte chunks inside each lane quad (4 consecutive pixels of a tile row): afterwards lane
                    // j of the quad holds chunk j of all four pixels
#pragma unroll
                    for (int sft = 2; sft >= 1; sft >>= 1) {
                        const bool up = (lane & sft) != 0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k & sft) continue;
                            const uint4 snd = up ? t4[k] : t4[k ^ sft];
                            uint4 rcv;
                            rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, sft);
                            rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, sft);
                            rcv.z = __shfl_xor_sync(0xffffffffu, snd.z, sft);
                            rcv.w = __shfl_xor_sync(0xffffffffu, snd.w, sft);
                            if (up) t4[k] = rcv; else t4[k ^ sft] = rcv;
                        }
                    }
                    const int lq = lane & 3;
                    if (yy < a.H && x - lq < a.W) {          // W % 4 == 0: a quad is live or dead as a whole
                        uint32_t* qb = a.y + (pix - lq) * a.y_ld + cb + lq * 4;
#pragma unroll
                        for (int k = 0; k < 4; ++k) st_global_v4(qb + (size_t)k * a.y_ld, t4[k]);
                    }
                } else if (x < a.W && yy < a.H) {
                    uint32_t* dst = a.y + pix * a.y_ld + cb;
                    if (full && !(a.y_ld & 3)) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) st_global_v4(dst + 4 * k, t4[k]);
                    } else {
                        const uint32_t o[16] = {t4[0].x, t4[0].y, t4[0].z, t4[0].w, t4[1].x, t4[1].y, t4[1].z, t4[1].w,
                                                t4[2].x, t4[2].y, t4[2].z, t4[2].w, t4[3].x, t4[3].y, t4[3].z, t4[3].w};
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (cb + j < a.cout_st) dst[j] = o[j];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
        if (a.range_flag && p16::any_nonfinite(bad)) *a.range_flag = 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// NT / slots / ring depth; returns the dynamic shared memory size or 0
int configure(ConvP16Args& h, int mode) {
    const int pitch = HT_W + h.KW - 1;
    const int b_stage = 2 * h.CoutP * 64;
    const int tile_cols = h.CoutP;
    // stacked tiles per work item: the whole weight tensor streams from L2 once per item, so NT sets the L2 -> SM traffic per
    // pixel (Cout = 128: 590 KB per item); PIVLFN_P16_NT = 1 | 2 | 4 overrides the choice (experiments)
    int NT = h.CoutP <= 64 ? 4 : 2;       // Cout <= 64: four tiles still leave two accumulator sets (epilogue overlapped with the next item)
    static int nt_env = -1;
    if (nt_env < 0) { const char* v = getenv("PIVLFN_P16_NT"); nt_env = v ? atoi(v) : 0; }
    if (nt_env == 1 || nt_env == 2 || nt_env == 4) NT = nt_env;
    if (h.wnc && NT > 2) NT = 2;
    while (NT > 1 && (NT * tile_cols > 512 || HT_H * (NT - 1) >= h.H)) NT >>= 1;
    for (; NT >= 1; NT >>= 1) {
        const int halo_rows = HT_H * NT + h.KH - 1;
        if (halo_rows > 256 || pitch > 256) continue;
        const int slot = (halo_rows * pitch * 128 + 1023) & ~1023;
        // fused backwarp: tap table; OUT_TAIL: the ScaleX / ScaleY weights
        const int taps = h.wnc ? ((halo_rows * pitch * 24 + 1023) & ~1023) : (h.out_fmt == OUT_TAIL ? 1024 : 0);
        const int budget = SMEM_BUDGET - taps;
        int tps = 1;
        if (!h.s2 && h.CoutP <= 64 && (h.KH * h.KW) % 3 == 0 && (budget - 2 * slot) / (3 * b_stage) >= 2) tps = 3;
        // a third activation slot when it still leaves a deep weight ring (fused backwarp: always, one slot is the gather's)
        int nA = ((budget - 3 * slot) / (tps * b_stage) >= (h.wnc ? 2 : 4)) ? 3 : 2;
        {
            static int na_env = -1;                       // experiment switch: PIVLFN_P16_NA = 2 | 3
            if (na_env < 0) { const char* v = getenv("PIVLFN_P16_NA"); na_env = v ? atoi(v) : 0; }
            if (!h.wnc && (na_env == 2 || (na_env == 3 && 3 * slot + 2 * tps * b_stage <= budget))) nA = na_env;
        }
        int nB = (budget - nA * slot) / (tps * b_stage);
        if (nB > MAX_B) nB = MAX_B;
        if (nB < 2) continue;
        h.NT = NT; h.nA = nA; h.nB = nB; h.tps = tps;
        h.nG = h.wnc ? 1 : 0; h.nT = nA - h.nG;
        h.nsets = (2 * NT * tile_cols <= 512) ? 2 : 1;
        h.tiles_x = cdiv(h.W, HT_W); h.tiles_y = cdiv(h.H, HT_H * NT);
        const long long total = (long long)h.tiles_x * h.tiles_y * h.N;
        if (total > 0x7FFFFFFFLL) return 0;
        h.total = (int)total;
        h.tap_off = nA * slot + nB * tps * b_stage;
        return h.tap_off + taps;
    }
    return 0;
}



struct WarpSrc { const void* src; int ld, p16; const float* flow; float scale; int c0, n; };

int conv_p16_impl(const void* x, int x_ld, int N, int H, int W, int Cin, const void* w_img, int mode,
                  const float* bias, void* y, int y_ld, int Cout, int KH, int KW, int stride, int lrelu,
                  int out_fmt, long long plane_stride, int* range_flag, const WarpSrc* ws, const TailArgs* tail, void* stream) {
    if (!x || !w_img || (!y && out_fmt != OUT_TAIL) || N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return PIVLFN_EINVAL;
    if (mode != 6) return PIVLFN_EINVAL;                  // 6 = f16 main product + e5m2 corrections (pivlfn.model._pack_f8)
    if (stride != 1 && stride != 2) return PIVLFN_EINVAL;
    if (out_fmt < OUT_P16 || out_fmt > OUT_TAIL || (out_fmt == OUT_TAIL) != (tail != nullptr)) return PIVLFN_EINVAL;
    const int CoutP = (Cout + 15) & ~15;
    if (CoutP > 128) return PIVLFN_EUNSUPPORTED;
    ConvP16Args h;
    h.wsrc = nullptr; h.wflow = nullptr; h.wscale = 0.f; h.wsrc_ld = 0; h.wsrc_p16 = 0; h.wc0 = 0; h.wnc = 0;
    int Cbuf = Cin;                                         // channels that live in the input buffer
    if (ws) {
        if (!ws->src || !ws->flow || stride != 1 || (ws->c0 & 31) || (ws->n & 31) || ws->n <= 0 || ws->c0 + ws->n > Cin) return PIVLFN_EINVAL;
        if (ws->p16 ? (((uintptr_t)ws->src & 63) || (ws->ld & 15)) : (((uintptr_t)ws->src & 15) || (ws->ld & 3))) return PIVLFN_EINVAL;
        if (ws->ld < ws->n || ((uintptr_t)ws->flow & 7)) return PIVLFN_EINVAL;
        h.wsrc = reinterpret_cast<const uint8_t*>(ws->src); h.wflow = reinterpret_cast<const float2*>(ws->flow); h.wscale = ws->scale;
        h.wsrc_ld = ws->ld; h.wsrc_p16 = ws->p16; h.wc0 = ws->c0 / 32; h.wnc = ws->n / 32;
        Cbuf = Cin - ws->n;
    }
    const int Cw = (Cbuf + 15) & ~15;                       // the P16 buffer holds whole 16-channel groups
    if (((uintptr_t)x & 63) || (x_ld & 15) || x_ld < Cw || ((uintptr_t)w_img & 15)) return PIVLFN_EINVAL;
    h.s2 = 0; h.cpp = 1;
    if (stride == 2) {
        if (KH != 3 || KW != 3 || (H & 1) || (W & 1) || (Cin % 32)) return PIVLFN_EUNSUPPORTED;
        h.s2 = 1; h.cpp = Cin / 32;
        h.KH = 2; h.KW = 2; h.H = H / 2; h.W = W / 2; h.x_shift = 0;
    } else {
        if (KH < 1 || KW < 1 || !(KH & 1) || !(KW & 1) || KH > 7 || KW > 7) return PIVLFN_EINVAL;
        h.KH = KH; h.KW = KW; h.H = H; h.W = W; h.x_shift = -(KW / 2);
    }
    h.bias = bias; h.y = reinterpret_cast<uint32_t*>(y); h.y_ld = y_ld;
    h.N = N; h.Cw = Cw; h.Cout = Cout; h.CoutP = CoutP; h.lrelu = lrelu; h.out_fmt = out_fmt;
    h.w_img = reinterpret_cast<const uint8_t*>(w_img); h.range_flag = range_flag; h.planar = 0; h.cout_st = Cout; h.quad = 0;
    {
        // experiment switch (default = the measured best): PIVLFN_P16_ISSUERS=1|2
        static int iss = -1;
        if (iss < 0) { const char* v = getenv("PIVLFN_P16_ISSUERS"); iss = v ? atoi(v) : 0; }
        h.one_issuer = iss == 1 ? 1 : 0;
    }
    if (out_fmt == OUT_P16) {
        if (((uintptr_t)y & 63) || (y_ld & 15) || y_ld < CoutP) return PIVLFN_EINVAL;
        h.quad = !(h.W & 3);
    } else if (out_fmt == OUT_F32) {
        if (((uintptr_t)y & 3) || y_ld < Cout) return PIVLFN_EINVAL;
        h.cout_st = (y_ld == ((Cout + 3) & ~3)) ? y_ld : Cout;
        h.quad = !(h.W & 3) && !((uintptr_t)y & 15) && !(y_ld & 3);
    } else if (out_fmt == OUT_PLANES) {
        if (((uintptr_t)y & 7) || plane_stride < 2LL * N * h.H * h.W) return PIVLFN_EINVAL;
        h.planar = plane_stride;
    } else {
        if (stride != 1 || ws || CoutP > 64) return PIVLFN_EUNSUPPORTED;
        if (tail->K != 3 && tail->K != 5 && tail->K != 7) return PIVLFN_EUNSUPPORTED;
        if (Cout != tail->K * tail->K || !tail->flow || !tail->wx || !tail->wy || !tail->bx || !tail->by || !tail->out) return PIVLFN_EINVAL;
        if (((uintptr_t)tail->flow & 7) || ((uintptr_t)tail->out & 7) || tail->flow == tail->out) return PIVLFN_EINVAL;
    }
    h.tl = tail ? *tail : TailArgs{};
    EncodeTiledFn enc = get_encode();
    if (!enc) return PIVLFN_EDRIVER;
    const int smem = configure(h, mode);
    if (smem <= 0) return PIVLFN_EUNSUPPORTED;
    const int halo_rows = HT_H * h.NT + h.KH - 1;
    CUtensorMap tmA;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cw, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)x_ld * 4, (cuuint64_t)W * x_ld * 4, (cuuint64_t)H * W * x_ld * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)((HT_W + h.KW - 1) * stride), (cuuint32_t)(halo_rows * stride), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        if (box[1] > 256 || box[2] > 256) return PIVLFN_EUNSUPPORTED;
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return PIVLFN_EINVAL;
    }
    const int nsm = pivlfn_num_sms();
    const int grid = h.total < nsm ? h.total : nsm;
    cudaStream_t st = (cudaStream_t)stream;
    static unsigned long long cfg6 = 0;
    cudaError_t e = pivlfn_optin_smem(conv_p16_kernel<6>, SMEM_BUDGET, cfg6);
    if (e != cudaSuccess) return (int)e;
    static int pdl = -1;                     // PIVLFN_P16_PDL=0: plain stream-ordered launches
    if (pdl < 0) { const char* v = getenv("PIVLFN_P16_PDL"); pdl = v ? atoi(v) : 1; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(P16_THREADS); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, conv_p16_kernel<6>, tmA, h);
    if (e != cudaSuccess) return (int)e;
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

}  // namespace

/* see include/pivlfn.h */
extern "C" int pivlfn_conv_p16(const void* x, int x_ld, int N, int H, int W, int Cin, const void* w_img, int mode,
                               const float* bias, void* y, int y_ld, int Cout, int KH, int KW, int stride, int lrelu,
                               int out_fmt, long long plane_stride, int* range_flag, void* stream) {
    return conv_p16_impl(x, x_ld, N, H, W, Cin, w_img, mode, bias, y, y_ld, Cout, KH, KW, stride, lrelu, out_fmt, plane_stride,
                         range_flag, nullptr, nullptr, stream);
}

/* see include/pivlfn.h */
extern "C" int pivlfn_conv_p16_warp(const void* x, int x_ld, int N, int H, int W, int Cin, const void* w_img, int mode,
                                    const float* bias, void* y, int y_ld, int Cout, int KH, int KW, int lrelu,
                                    const void* wsrc, int wsrc_ld, int wsrc_p16, const float* wflow, float wscale,
                                    int wc0, int wn, int* range_flag, void* stream) {
    WarpSrc ws{wsrc, wsrc_ld, wsrc_p16, wflow, wscale, wc0, wn};
    return conv_p16_impl(x, x_ld, N, H, W, Cin, w_img, mode, bias, y, y_ld, Cout, KH, KW, 1, lrelu, OUT_P16, 0, range_flag, &ws, nullptr, stream);
}

/* see include/pivlfn.h */
extern "C" int pivlfn_conv_p16_tail(const void* x, int x_ld, int N, int H, int W, int Cin, const void* w_img, const float* bias,
                                    int KH, int KW, int K, const float* flow_in, const float* wx, const float* bx, const float* wy,
                                    const float* by, float* flow_out, float* out_nchw, float final_scale, void* stream) {
    TailArgs t{reinterpret_cast<const float2*>(flow_in), wx, wy, bx, by, reinterpret_cast<float2*>(flow_out), out_nchw, final_scale, K};
    return conv_p16_impl(x, x_ld, N, H, W, Cin, w_img, 6, bias, nullptr, 0, K * K, KH, KW, 1, 0, OUT_TAIL, 0, nullptr, nullptr, &t, stream);
}
