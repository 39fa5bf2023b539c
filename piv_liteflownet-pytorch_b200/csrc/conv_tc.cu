// 3x3 stride-1 convolution (+ bias + LeakyReLU) as an implicit GEMM on the 5th-generation tensor cores:
//   D[128 pixels x Cout] += A[128 pixels x 32 ch] * B[32 ch x Cout]     per (filter tap, 32-channel chunk)
// tcgen05.mma kind::tf32 with the accumulator in TMEM, operands staged in shared memory by TMA
// (cp.async.bulk.tensor, 128B swizzle), mbarrier producer/consumer pipeline, warp-specialised:
//
//   warp 0        TMA producer: per stage one 4-D box of the NHWC activations at the tap's (dy,dx) offset
//                 (zero padding = TMA out-of-bounds fill, so no im2col buffer and no border code) and one box of
//                 the packed weights [Cout][tap][CinP]
//   warp 1        TMEM allocation + MMA issue (one elected lane)
//   warps 2..5    3xTF32 mode: split every landed A tile in shared memory into hi = rna_tf32(a) (in place) and
//                 lo = a - hi (second buffer);  all modes: epilogue TMEM -> registers -> bias + LeakyReLU -> NHWC store
//
// passes == 1: plain TF32 (D = A*Bhi).  passes == 3: error-compensated 3xTF32 (D = Ahi*Bhi + Alo*Bhi + Ahi*Blo),
// which reproduces fp32 convolution to ~1e-6 relative.  Replaces torch.nn.Conv2d(k=3,s=1,p=1)+LeakyReLU(0.1) of
// src/models.py:77-106 (NetC), :154-160 (conv_M), :197-204 (conv_S), :236-250 (conv_R).
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "p16.cuh"
#include "tc_ptx.cuh"

namespace {
using namespace tcptx;

constexpr int KC = 32;               // channels per stage: 32 fp32 = one 128-byte swizzle row
constexpr int A_BYTES = TILE_M * KC * 4;   // 16 KB
constexpr int NTHREADS = 192;
constexpr int EPI_WARP0 = 2;

// Sticky flag: an activation outside the fp16 range reached a PASSES == 4 convolution (host: pivlfn_f16_range_flag)
__device__ int g_f16_range_flag = 0;

struct ConvTcArgs {
    const float* bias;
    const float* res;        // optional residual view (added after the activation), Cout channels
    float* y;
    int y_ld, res_ld;
    int N, H, W, Cin;
    int Cout, CoutP;         // real output channels / MMA N (multiple of 16, zero-padded weights)
    int KW, ntaps, ox, oy;   // tap t reads the input at (x + t % KW + ox, y + t / KW + oy)
    int bw, bh, bn;          // spatial / batch extent of the 128-pixel tile (bw * bh * bn == 128)
    int tiles_x, tiles_y;
    int lrelu, stages, vec_store;
    int sx;                  // convolution stride (1 or 2): H, W above are the OUTPUT size, the tensor map walks the
                             // input with element stride sx so that a tap's box holds exactly the sampled pixels
};

constexpr int MAX_STAGES = 8;


// PASSES = 1 (TF32) or 3 (3xTF32).
template <int PASSES>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
               const __grid_constant__ CUtensorMap tmBlo, const ConvTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stage][A | Alo (3-pass) | Bhi | Blo (3-pass)], every tile 1024-byte aligned
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = a.CoutP * KC * 4;
    const int stage_bytes = (PASSES == 3 ? 2 : 1) * (A_BYTES + b_bytes);
    const int STAGES = a.stages;
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], ready_bar[MAX_STAGES], empty_bar[MAX_STAGES], accum_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float bias_s[128];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunk = (a.Cin + KC - 1) / KC;
    const int niter = a.ntaps * nchunk;
    const uint32_t ncols = tmem_cols_for(PASSES == 3 ? 2 * a.CoutP : a.CoutP);

    // tile coordinates
    const int tx = blockIdx.x % a.tiles_x;
    const int ty = (blockIdx.x / a.tiles_x) % a.tiles_y;
    const int tn = blockIdx.x / (a.tiles_x * a.tiles_y);
    const int x0 = tx * a.bw, y0 = ty * a.bh, n0 = tn * a.bn;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&ready_bar[s], 4);      // one arrival per split warp
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&accum_bar, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmBhi);
        if (PASSES == 3) tma_prefetch_desc(&tmBlo);
    }
    if (warp == 1) {
        // TMEM: fp32 accumulator columns (main, plus the low-order correction accumulator in 3xTF32 mode)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= EPI_WARP0) {
        for (int i = threadIdx.x - EPI_WARP0 * 32; i < a.CoutP; i += 128) bias_s[i] = (a.bias && i < a.Cout) ? a.bias[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < niter; ++it) {
                const int tap = it / nchunk, ch = it - tap * nchunk;
                const int dy = tap / a.KW + a.oy, dx = tap % a.KW + a.ox;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* st = smem + (size_t)stage * stage_bytes;
                uint8_t* sA = st;
                uint8_t* sBhi = st + (PASSES == 3 ? 2 : 1) * A_BYTES;
                mbar_expect_tx(&full_bar[stage], A_BYTES + (PASSES == 3 ? 2 : 1) * b_bytes);
                tma_load_4d(sA, &tmA, &full_bar[stage], ch * KC, x0 * a.sx + dx, y0 * a.sx + dy, n0);
                tma_load_3d(sBhi, &tmBhi, &full_bar[stage], ch * KC, tap, 0);
                if (PASSES == 3) tma_load_3d(sBhi + b_bytes, &tmBlo, &full_bar[stage], ch * KC, tap, 0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_tf32(a.CoutP);
            const uint32_t tmem_corr = tmem_base + (uint32_t)a.CoutP;      // 3xTF32: low-order terms accumulate separately
            int stage = 0;
            uint32_t phase = 0;
            uint32_t acc = 0, acc_corr = 0;
            for (int it = 0; it < niter; ++it) {
                const int tap = it / nchunk, ch = it - tap * nchunk;
                const int kleft = a.Cin - ch * KC;
                const int nk = kleft >= KC ? KC / 8 : (kleft + 7) / 8;      // K = 8 per tf32 MMA
                mbar_wait(PASSES == 3 ? &ready_bar[stage] : &full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sA = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t sBhi = sA + (PASSES == 3 ? 2 : 1) * A_BYTES;
                const uint64_t dA = make_smem_desc(sA), dBhi = make_smem_desc(sBhi);
                const uint64_t dAlo = make_smem_desc(sA + A_BYTES), dBlo = make_smem_desc(sBhi + b_bytes);
                for (int k = 0; k < nk; ++k) {
                    const uint64_t koff = (uint64_t)(k * 2);      // 32 bytes >> 4
                    umma_tf32(tmem_base, dA + koff, dBhi + koff, idesc, acc);
                    acc = 1;
                    if (PASSES == 3) {
                        umma_tf32(tmem_corr, dAlo + koff, dBhi + koff, idesc, acc_corr);
                        umma_tf32(tmem_corr, dA + koff, dBlo + koff, idesc, 1);
                        acc_corr = 1;
                    }
                }
                umma_commit(&empty_bar[stage]);                    // frees the smem slot when these MMAs retire
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(&accum_bar);                               // accumulators complete
        }
    } else {
        // ============================ split warps + epilogue ============================
        const int et = threadIdx.x - EPI_WARP0 * 32;               // 0..127
        if (PASSES == 3) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < niter; ++it) {
                mbar_wait(&full_bar[stage], phase);
                float4* pa = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes);
                float4* pl = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes + A_BYTES);
#pragma unroll
                for (int i = 0; i < A_BYTES / 16 / 128; ++i) {
                    const int idx = et + i * 128;
                    float4 v = pa[idx];
                    float4 h, l;
                    h.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xFFFFE000u);
                    h.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xFFFFE000u);
                    h.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xFFFFE000u);
                    h.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xFFFFE000u);
                    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                    pa[idx] = h;
                    pl[idx] = l;
                }
                fence_proxy_async();            // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        // ---- epilogue: thread = one accumulator row (pixel); TMEM lane quarter = warp % 4 ----------------------
        mbar_wait(&accum_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int px = row % a.bw, py = (row / a.bw) % a.bh, pn = row / (a.bw * a.bh);
        const int x = x0 + px, yy = y0 + py, n = n0 + pn;
        const bool live = x < a.W && yy < a.H && n < a.N;
        const size_t pix = ((size_t)n * a.H + yy) * a.W + x;
        float* dst = a.y + pix * a.y_ld;
        const float* rsd = a.res ? a.res + pix * a.res_ld : nullptr;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < a.CoutP; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(trow + (uint32_t)c0, v);
            if (PASSES == 3) {
                uint32_t u[16];
                tmem_ld16(trow + (uint32_t)(a.CoutP + c0), u);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            }
            if (live) {
                float o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float t = __uint_as_float(v[j]) + bias_s[c0 + j];
                    if (a.lrelu) t = lrelu_f(t);
                    o[j] = t;
                }
                if (a.vec_store && c0 + 16 <= a.Cout) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 w4 = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                        if (rsd) { w4.x += rsd[c0 + j]; w4.y += rsd[c0 + j + 1]; w4.z += rsd[c0 + j + 2]; w4.w += rsd[c0 + j + 3]; }
                        *reinterpret_cast<float4*>(dst + c0 + j) = w4;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < a.Cout) dst[c0 + j] = o[j] + (rsd ? rsd[c0 + j] : 0.f);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// =================================================================================================================
// Halo-resident variant (W >= 8, at least 3 taps).  The per-tap kernel above re-fetches a shifted 16 KB activation
// tile for every filter tap; here each 32-channel chunk of the input is fetched ONCE as a halo tile
//     [16*NT + KH - 1 rows][8 + KW - 1 pixels][32 ch]     (pixel pitch 128 B, dense rows, 128B swizzle)
// and every tap's A operand is a shifted window INTO that tile: the 8 rows of one core group are the 8 pixels of one
// output row (tile = 8 wide x 16 tall), consecutive groups are one halo row apart (SBO = pitch * 128 B).  The window
// start (ky * pitch + kx) * 128 B is not 1024-byte aligned; measured on B200 the swizzle XOR uses absolute
// shared-memory address bits (the same rule TMA used when it wrote the tile), so the descriptor's base-offset field
// stays 0 and any pitch works.  One CTA computes NT vertically stacked tiles (NT accumulator sets in TMEM) so that
// every streamed weight tile is used for NT * 128 pixels -- the weights re-streamed from L2 by every CTA were the
// bottleneck of the per-tap kernel.  3xTF32: buffers rotate raw(c+1) / hi(c) / lo(c) through three equal slots.
// =================================================================================================================
constexpr int HT_W = 8, HT_H = 16;         // output tile (UMMA M = 128 rows = 16 rows of 8 pixels)
constexpr int MAX_NT = 4;

__device__ __forceinline__ uint64_t make_smem_desc_halo(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;          // next 8-pixel group = next halo row
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

struct ConvHaloArgs {
    const float* bias;
    const float* res;
    float* y;
    int y_ld, res_ld;
    int N, H, W, Cin;
    int Cout, CoutP;
    int KH, KW;
    int tiles_x, tiles_y, total;   // work items = N * tiles_y * tiles_x groups of NT stacked 8x16 tiles
    int lrelu, vec_store;
    int NT;                  // vertically stacked 8x16 tiles per work item
    int nBuf, nB;            // activation buffer slots, weight ring depth
    int s2, cpp;             // s2 = 1: 3x3 stride-2 convolution restated as 2x2 block taps over the four pixel-parity phases of
                             // the input (space-to-depth done by TMA element strides): chunk c = parity (c / cpp) x 32-channel
                             // group (c % cpp); H, W are the OUTPUT size and Cin = 4 * input channels
    int tps;                 // filter taps per weight-ring stage (1, or 3 for small Cout: one barrier round trip and one
                             // tcgen05.commit per THREE taps -- at N <= 64 the per-tap waits / fences of the issuing thread, not
                             // the tensor pipe, set the pace: measured 320 cycles per (tap, 16 channels) against a pipe floor of
                             // 176, tools/ubench/umma_f16_rate.cu)
    const uint8_t* w_img;    // fp16 modes: the weights as the shared-memory image of the ring stages (pivlfn.model.stage_image:
                             // [chunk][tap][part][cout row][64 B, 64B-swizzled]) -- one stage = one contiguous bulk copy
    int slot_mode;           // 1 (fp16 modes with 3 slots): one raw slot + two pair slots, see slot_x / slot_l
    int corr;                // 3xTF32: accumulate the low-order terms in their own TMEM accumulator
    int nsets;               // TMEM accumulator sets (2 = epilogue of item i overlaps the MMAs of item i+1)
    int bo_mode;             // 0 (default): descriptor base_offset = 0 (see above); 1: base_offset = kx -- wrong on
                             // B200, kept as an experiment switch
    int cout_st;             // channels the epilogue may store: Cout, or Cout rounded up to 4 when the output rows are
                             // exactly that wide (dense buffer with its own padding) so that float4 stores can be used
    long long planar;        // 0: NHWC rows; else the output is stored as channel-PAIR planes, plane p = channels (2p, 2p+1)
                             // of all pixels, [pair][pixel][2] with this many floats per plane (flow-head tap planes)
    int x_shift;             // x coordinate of tap column 0 relative to the output pixel (normally -(KW/2); +1 for the stem)
    int split_trunc;         // 3xTF32 split: 1 (default) = leave a in place (measured: the tensor core reads only the top
                             // 19 bits of an fp32 operand, i.e. truncates) and write lo = a - trunc_tf32(a);
                             // 0 = hi = rna_tf32(a) written back, lo = a - hi
    int stage_off;           // vec_store == 5: byte offset (from the aligned dynamic shared memory base) of the epilogue staging
                             // area, 2 KB per epilogue warp; the output leaves through TMA tile stores (tmY)
    int stagger;             // > 0: CTA i delays its first load by stagger * i / gridDim.x cycles (one work-item period spread
                             // over the grid).  With a single TMEM accumulator set (Cout = 128) the epilogue cannot overlap the
                             // MMAs, and 148 CTAs in lock step all store their 128 KB tiles at the same moment: the burst runs at
                             // HBM write speed (measured 10.2k cycles per item) while the tensor cores idle.  De-phased CTAs
                             // store at the average rate instead.
    long long* dbg;          // optional: CTA 0 writes per-role wait-cycle totals (tools/profile_conv.py --trace)
    int out_p16;             // quad-store path only: the output rows are P16 groups (p16.cuh) instead of fp32
    int* range_flag;         // out_p16: raised when a result is not finite in fp16
};

#define DBG_T0() long long _t0 = a.dbg ? clock64() : 0
#define DBG_ADD(var) do { if (a.dbg) var += clock64() - _t0; } while (0)

// Stride-2 restatement (ConvHaloArgs::s2): block tap t = (ty, tx) of parity phase par = (py, px) carries a non-zero weight
// only if ty >= 1 - py and tx >= 1 - px (input row 2y-1 is the odd row of block y-1; the even row of block y-1 is never
// read).  Every role skips the other (tap, parity) pairs in the same way: 9 of 16 products remain, the original 9 taps.
__device__ __forceinline__ bool s2_tap_used(int t, int par) {
    return ((t >> 1) >= 1 - (par >> 1)) && ((t & 1) >= 1 - (par & 1));
}

// buffer slot of chunk c's activations (hi after the split) and of its low-order part
// nbuf == 4: two (x, lo) pairs, plain double buffering -- the split of chunk c+1 overlaps the MMAs of chunk c;
// nbuf == 3: rotation raw(c+1) / x(c) / lo(c) when shared memory has no room for a fourth slot (N = 128)
// nbuf == -3 (fp16 modes, where the MMAs never read the raw tile): one raw slot + two pair slots; the raw slot is free as
// soon as chunk c's split is done, the pair slots alternate like in the 4-slot scheme
__device__ __forceinline__ int slot_x(int c, int nbuf) {
    if (nbuf == -3) return 0;
    return nbuf == 4 ? 2 * (c & 1) : (nbuf == 3 ? ((c % 3) == 0 ? 0 : ((c % 3) == 1 ? 2 : 1)) : c % nbuf);
}
__device__ __forceinline__ int slot_l(int c, int nbuf) {
    if (nbuf == -3) return 1 + (c & 1);
    return nbuf == 4 ? 2 * (c & 1) + 1 : (nbuf == 3 ? ((c % 3) == 0 ? 1 : ((c % 3) == 1 ? 0 : 2)) : 1);
}

constexpr int HALO_THREADS = 640;      // warp 0 TMA, 1-2 MMA issue, 4-7 + 16-19 epilogue (TMEM lane quarter = warp % 4), 8-15 split
constexpr int SPLIT_THREADS = 256;

// Persistent: one CTA per SM loops over work items; all roles walk the same global (item, chunk, tap) sequence so
// the mbarrier phases simply keep counting across items.
template <int PASSES>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                    const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmB16,
                    const __grid_constant__ CUtensorMap tmBlo16, const __grid_constant__ CUtensorMap tmY,
                    const ConvHaloArgs a) {
    // PASSES: 1 = plain TF32; 3 = 3xTF32 (all three products in tf32); 2 = TF32 main product + the two low-order
    // products in bf16 (kind::f16, K = 16: half the MMA instructions of the tf32 corrections, same fp32 accuracy class
    // because the corrections are ~2^-11 of the result and bf16 keeps 8 bits of them)
    //         4 = all three products in fp16 (kind::f16, fp16 has the 11 significant bits of tf32 at twice the MMA rate):
    //             a = f16(a) + 2^-11 * f16((a - f16(a)) * 2^11), same for w; D = a_hi*w_hi + 2^-11 * (a_lo'*w_hi + a_hi*w_lo');
    //             the raw fp32 tile is only the source of the split.  Activations must lie in the fp16 range (|a| < 65504;
    //             g_f16_range_flag is raised otherwise and the host falls back to mode 2).
    constexpr bool SPLIT = PASSES >= 2;
    //         5 = mode 4 with ONE accumulator per tile (Cout = 128: [main | corr] for two stacked tiles fills all 512 TMEM
    //             columns, so the epilogue cannot overlap the next item's MMAs and its global stores -- a per-SM limit of
    //             ~18 B/clk -- leave the tensor cores idle a quarter of the time).  Weights are pre-scaled, W = 256 w, and come
    //             as three tiles [f16(W) | f16(W - f16(W)) | f16(f16(W) * 2^-11)]: a_hi*W_hi + a_hi*W_lo + a_lo'*W_hi2 all
    //             carry the same scale and add up in the same TMEM columns; the epilogue multiplies by 2^-8.
    constexpr bool F16 = PASSES == 4 || PASSES == 5;
    constexpr bool F16D = PASSES == 4;          // dual accumulators [main | corr]
    constexpr bool F16S = PASSES == 5;          // single accumulator, three weight tiles
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int pitch = HT_W + a.KW - 1;
    const int halo_rows = HT_H * a.NT + a.KH - 1;
    const int halo_bytes = halo_rows * pitch * 128;
    const int slot_bytes = (halo_bytes + 1023) & ~1023;
    const int b_bytes = a.CoutP * KC * 4;
    const int b_stage = F16S ? 3 * (b_bytes / 2) : F16 ? b_bytes : (SPLIT ? 2 : 1) * b_bytes;      // [hi | lo], [hi | bf16(w) | bf16(w_lo)] or [f16(w) | f16(w_lo')]
    uint8_t* smemB = smem + (size_t)a.nBuf * slot_bytes;
    const int nbuf_k = a.slot_mode ? -3 : a.nBuf;
    __shared__ __align__(8) uint64_t a_full[2], a_ready[2], chunk_done[2], b_full[MAX_STAGES], b_empty[MAX_STAGES],
        acc_full[2], acc_empty[2];
    __shared__ long long dbg_a_issue[2];   // trace only: clock of the A load of chunk parity 0 / 1
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[128];   // fits in the 1 KB the static part is padded to anyway

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunk = (a.Cin + KC - 1) / KC;
    const int ntaps = a.KH * a.KW;
    if (threadIdx.x >= 128 && threadIdx.x < 256) {
        const int i = threadIdx.x - 128;
        bias_s[i] = (a.bias && i < a.Cout) ? a.bias[i] : 0.f;
    }
    const int set_cols = ((SPLIT && a.corr) ? 2 : 1) * a.NT * a.CoutP;
    const uint32_t ncols = tmem_cols_for(a.nsets * set_cols);
    const int G = gridDim.x;

    if (threadIdx.x == 0) {
        const uint32_t n_iss = a.NT >= 2 ? 2u : 1u;           // MMA-issuing threads, each commits to the barriers
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1); mbar_init(&a_ready[i], SPLIT_THREADS / 32); mbar_init(&chunk_done[i], n_iss);
            mbar_init(&acc_full[i], n_iss); mbar_init(&acc_empty[i], 8);
        }
        for (int i = 0; i < a.nB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], n_iss); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        if (!F16) tma_prefetch_desc(&tmBhi);
        if (PASSES == 3 || F16S) tma_prefetch_desc(&tmBlo);
        if (PASSES == 2 || F16) tma_prefetch_desc(&tmB16);
        if (a.vec_store == 5) tma_prefetch_desc(&tmY);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0 || warp == 3) {
        // ================================ TMA producer(s) =============================
        // fp16 modes with two raw slots: the activation tiles and the weight stages are issued by two different threads
        // (warp 0 / warp 3).  A single producer sits inside its tensor-load instructions nearly all the time (the engine's
        // queue is short), so the tile of chunk g+1 would only be issued after the weight stages of chunk g -- too late for
        // the load -> split chain of a thin layer.  The two streams only meet at the barriers they already use.
        const bool two_prod = F16 && a.nBuf == 4 && !a.slot_mode;
        const bool do_A = warp == 0, do_W = two_prod ? warp == 3 : warp == 0;
        if ((do_A || do_W) && elect_one()) {
            int bs = 0;
            uint32_t bphase = 0;
            int gc = 0;                                   // global chunk counter
            long long p_b = 0, p_a = 0;
            const long long p_begin = a.dbg ? clock64() : 0;
            // the chunk sequence (item, chunk) flattened: chunk g+1 follows chunk g across item boundaries
            auto next_of = [&](int w, int c, int& w2, int& c2) { c2 = c + 1; w2 = w; if (c2 == nchunk) { c2 = 0; w2 = w + G; } };
            auto coords_A = [&](int w, int c, int& c0, int& c1, int& c2, int& c3) {
                const int tx = w % a.tiles_x, ty = (w / a.tiles_x) % a.tiles_y;
                c3 = w / (a.tiles_x * a.tiles_y);
                if (a.s2) {
                    // parity phase (py, px) of the input, sampled every second pixel; the tile starts one block left / above
                    const int par = c / a.cpp, cc = c - par * a.cpp;
                    c0 = cc * KC; c1 = 2 * (tx * HT_W - 1) + (par & 1); c2 = 2 * (ty * HT_H * a.NT - 1) + (par >> 1);
                } else { c0 = c * KC; c1 = tx * HT_W + a.x_shift; c2 = ty * HT_H * a.NT - a.KH / 2; }
            };
            auto load_A = [&](int g, int w, int c) {
                // slot_x(g) was last read by global chunk g-2's MMAs (as its hi or its lo slot)
                // (slot_mode 1: the single raw slot is free once chunk g-1 has been split)
                // fp16 modes: the MMAs never read the raw tile, so its slot is free as soon as its previous tenant has been
                // split (chunk g-2 with two raw slots) -- a whole chunk earlier than the MMAs retire
                if (a.slot_mode) {
                    if (g >= 1) { DBG_T0(); mbar_wait(&a_ready[(g - 1) & 1], (uint32_t)(((g - 1) >> 1) & 1)); DBG_ADD(p_a); }
                } else if (F16 && a.nBuf == 4) {
                    if (g >= 2) { DBG_T0(); mbar_wait(&a_ready[(g - 2) & 1], (uint32_t)(((g - 2) >> 1) & 1)); DBG_ADD(p_a); }
                } else if (g >= 2) { DBG_T0(); mbar_wait(&chunk_done[(g - 2) & 1], (uint32_t)(((g - 2) >> 1) & 1)); DBG_ADD(p_a); }
                mbar_expect_tx(&a_full[g & 1], halo_bytes);
                if (a.dbg) *(volatile long long*)&dbg_a_issue[g & 1] = clock64();
                int c0, c1, c2, c3;
                coords_A(w, c, c0, c1, c2, c3);
                tma_load_4d(smem + (size_t)slot_x(g, nbuf_k) * slot_bytes, &tmA, &a_full[g & 1], c0, c1, c2, c3);
            };
            bool pre = false;                             // is chunk gc already issued?
            if (a.stagger > 0) {
                const long long t0 = clock64(), d = (long long)a.stagger * blockIdx.x / G;
                while (clock64() - t0 < d) __nanosleep(256);
            }
            if (do_A && !do_W) {
                for (int w = blockIdx.x; w < a.total; w += G)
                    for (int c = 0; c < nchunk; ++c, ++gc) load_A(gc, w, c);
            } else
            for (int w = blockIdx.x; w < a.total; w += G) {
                for (int c = 0; c < nchunk; ++c, ++gc) {
                    if (!do_A) pre = true;                // (the other producer thread loads the tiles)
                    if (!pre) load_A(gc, w, c);
                    pre = false;
                    int w2, c2;
                    next_of(w, c, w2, c2);
                    const bool has_next = w2 < a.total;
                    // with >= 3 slots (or 2 slots in 1-pass mode) the next chunk's slot is free as soon as chunk gc-1
                    // retired; issue it after nB weight tiles (by then chunk gc-1 has certainly retired)
                    bool next_issued = !has_next || a.nBuf < 2 || !do_A;
                    if (!next_issued && gc == 0) { load_A(gc + 1, w2, c2); next_issued = true; pre = true; }
                    int issued = 0;
                    for (int t = 0; t < ntaps; ++t) {
                        if (a.s2 && !s2_tap_used(t, c / a.cpp)) continue;
                        const int sub = a.tps > 1 ? t % a.tps : 0;            // position inside the stage (tps divides ntaps)
                        if (sub == 0) {
                            ++issued;
                            { DBG_T0(); mbar_wait(&b_empty[bs], bphase ^ 1); DBG_ADD(p_b); }
                            mbar_expect_tx(&b_full[bs], a.tps * b_stage);
                        }
                        uint8_t* sB = smemB + ((size_t)bs * a.tps + sub) * b_stage;
                        if (F16) {
                            // the tps taps x [hi | lo (| hi2)] tiles of the stage are contiguous in the weight image
                            if (sub == 0) bulk_load(sB, a.w_img + ((size_t)c * ntaps + t) * b_stage, (uint32_t)(a.tps * b_stage), &b_full[bs]);
                        } else tma_load_3d(sB, &tmBhi, &b_full[bs], c * KC, t, 0);
                        if (PASSES == 3) tma_load_3d(sB + b_bytes, &tmBlo, &b_full[bs], c * KC, t, 0);
                        if (PASSES == 2) tma_load_4d(sB + b_bytes, &tmB16, &b_full[bs], c * KC, 0, 0, t);   // [bf16(w) | bf16(w_lo)]
                        if (sub == a.tps - 1) {
                            if (++bs == a.nB) { bs = 0; bphase ^= 1; }
                            if (!next_issued && issued >= a.nB) { load_A(gc + 1, w2, c2); next_issued = true; pre = true; }
                        }
                    }
                    if (!next_issued && a.nBuf >= 2) { load_A(gc + 1, w2, c2); pre = true; }
                }
            }
            if (a.dbg && blockIdx.x == 0) { if (do_W) a.dbg[5] = p_b; if (do_A) { a.dbg[6] = p_a; a.dbg[13] = clock64() - p_begin; } }
        }
    } else if (warp == 1 || warp == 2) {
        // ================================ MMA issuers =================================
        // Two issuing threads (warps 1 and 2): issuer j feeds the stacked tiles i = j, j+2, ... (own TMEM accumulators),
        // because ONE thread cannot issue tf32 MMAs (64 cycles each at N = 128) fast enough.
        if (elect_one()) {
            const int issuer = warp - 1;
            const int n_issuers = a.NT >= 2 ? 2 : 1;
            if (issuer < n_issuers) {
            // Issue-rate matters: one thread feeds the tensor core, and every scalar instruction between two
            // tcgen05.mma is on the critical path (measured ~100 cycles per MMA with naive 64-bit descriptor math,
            // vs a 64-cycle MMA).  All descriptors are therefore (constant high word, 32-bit low word) pairs and the
            // per-MMA work is one or two 32-bit adds.
            const uint32_t idesc = make_idesc_tf32(a.CoutP);
            const uint32_t hiA = (uint32_t)(make_smem_desc_halo(0, (uint32_t)(pitch * 128), 0) >> 32);
            const uint32_t hiB = (uint32_t)(make_smem_desc(0) >> 32);
            const uint32_t lbo_bits = 1u << 16;
            const uint32_t tile16 = (uint32_t)(HT_H * pitch * 8);       // one stacked tile further, in 16-byte units
            const uint32_t blo16 = (uint32_t)(b_bytes >> 4);
            const uint32_t corr_off = a.corr ? (uint32_t)(a.NT * a.CoutP) : 0u;
            const uint32_t pitch8 = (uint32_t)(pitch * 8);
            // bf16 operand descriptors (mode 2): 64-byte rows, SWIZZLE_64B (layout type 4), 8-row group stride = pitch * 64 B
            const uint32_t hiA16 = (uint32_t)((((uint64_t)((pitch * 64) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61)) >> 32);
            const uint32_t hiB16 = (uint32_t)((((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61)) >> 32);
            const uint32_t idesc16 = F16 ? make_idesc_f16(a.CoutP) : make_idesc_bf16(a.CoutP);
            const uint32_t idesc16w = make_idesc_f16(2 * a.CoutP);
            const uint32_t tile_cols = (uint32_t)(F16D ? 2 * a.CoutP : a.CoutP);    // mode 4: [main | corr] per stacked tile
            const uint32_t half16 = (((uint32_t)(halo_rows * pitch * 64) + 511u) & ~511u) >> 4;       // bf16(a_lo) tile behind bf16(a)
            int bs = 0;
            uint32_t bphase = 0;
            int gc = 0, wl = 0;
            long long w_acc = 0, w_a = 0, w_b = 0;
            const long long t_begin = a.dbg ? clock64() : 0;
            if (F16 && !a.s2 && a.NT <= 2) {
                // ---- lean issue loop of the fp16 modes ------------------------------------------------------------------
                // The issuing thread is a single in-order thread: the ~150 dependent scalar instructions the general loop
                // below spends per tap (tap bookkeeping, stage addresses, mode tests) cost ~600 cycles, more than the MMAs
                // of a tap take on the tensor pipe when Cout <= 64 (352 / 448 cycles for Cout = 32 / 64 over both tiles,
                // tools/ubench/umma_f16_rate.cu).  Here a tap is: one add for the A window, one for the weight tile, the MMAs.
                const uint32_t tileA = (uint32_t)issuer * (tile16 / 2);
                const uint32_t stage16 = (uint32_t)(b_stage >> 4);
                const uint32_t ring16 = stage16 * (uint32_t)a.tps;
                const uint32_t bbase16 = (smem_u32(smemB) >> 4) | lbo_bits;
                const uint32_t row_step = (uint32_t)((pitch - a.KW) * 4);      // A window: next filter row, in 16-byte units
                const uint32_t cP = (uint32_t)a.CoutP;
                uint32_t bcur = bbase16;
                for (int w = blockIdx.x; w < a.total; w += G, ++wl) {
                    const int as = wl % a.nsets;
                    const uint32_t use = (uint32_t)(wl / a.nsets);
                    { DBG_T0(); mbar_wait(&acc_empty[as], (use & 1) ^ 1); DBG_ADD(w_acc); }
                    tc_fence_after();
                    const uint32_t t_main = tmem_base + (uint32_t)(as * set_cols) + (uint32_t)issuer * tile_cols;
                    uint32_t acc = 0;
                    for (int c = 0; c < nchunk; ++c, ++gc) {
                        const bool two = a.Cin - c * KC > 16;                      // second 16-channel K step present
                        { DBG_T0(); mbar_wait(&a_ready[gc & 1], (uint32_t)((gc >> 1) & 1)); DBG_ADD(w_a); }
                        tc_fence_after();
                        uint32_t A = ((smem_u32(smem + (size_t)slot_l(gc, nbuf_k) * slot_bytes) >> 4) | lbo_bits) + tileA;
                        int kx = 0;
                        for (int t = 0; t < ntaps;) {
                            { DBG_T0(); mbar_wait(&b_full[bs], bphase); DBG_ADD(w_b); }
                            tc_fence_after();
                            uint32_t b = bcur;
                            for (int sub = 0; sub < a.tps; ++sub, ++t, b += stage16) {
                                if (F16S) {
                                    umma_bf16_lohi(t_main, A, hiA16, b, hiB16, idesc16, acc);
                                    umma_bf16_lohi(t_main, A, hiA16, b + blo16 / 2, hiB16, idesc16, 1);
                                    umma_bf16_lohi(t_main, A + half16, hiA16, b + blo16, hiB16, idesc16, 1);
                                    if (two) {
                                        umma_bf16_lohi(t_main, A + 2, hiA16, b + 2, hiB16, idesc16, 1);
                                        umma_bf16_lohi(t_main, A + 2, hiA16, b + blo16 / 2 + 2, hiB16, idesc16, 1);
                                        umma_bf16_lohi(t_main, A + half16 + 2, hiA16, b + blo16 + 2, hiB16, idesc16, 1);
                                    }
                                } else {
                                    umma_bf16_lohi(t_main, A, hiA16, b, hiB16, idesc16w, acc);
                                    umma_bf16_lohi(t_main + cP, A + half16, hiA16, b, hiB16, idesc16, 1);
                                    if (two) {
                                        umma_bf16_lohi(t_main, A + 2, hiA16, b + 2, hiB16, idesc16w, 1);
                                        umma_bf16_lohi(t_main + cP, A + half16 + 2, hiA16, b + 2, hiB16, idesc16, 1);
                                    }
                                }
                                acc = 1;
                                A += 4;                                             // next tap of the filter row
                                if (++kx == a.KW) { kx = 0; A += row_step; }
                            }
                            umma_commit(&b_empty[bs]);
                            if (++bs == a.nB) { bs = 0; bphase ^= 1; bcur = bbase16; } else bcur += ring16;
                        }
                        umma_commit(&chunk_done[gc & 1]);
                    }
                    umma_commit(&acc_full[as]);
                }
            } else
            for (int w = blockIdx.x; w < a.total; w += G, ++wl) {
                const int as = wl % a.nsets;
                const uint32_t use = (uint32_t)(wl / a.nsets);
                { DBG_T0(); mbar_wait(&acc_empty[as], (use & 1) ^ 1); DBG_ADD(w_acc); }   // epilogue drained this set
                tc_fence_after();
                const uint32_t tset = tmem_base + (uint32_t)(as * set_cols);
                uint32_t acc = 0;
                for (int c = 0; c < nchunk; ++c, ++gc) {
                    const uint32_t par = (uint32_t)((gc >> 1) & 1);
                    const int kleft = a.Cin - c * KC;
                    const int nk = kleft >= KC ? KC / 8 : (kleft + 7) / 8;
                    { DBG_T0(); mbar_wait(SPLIT ? &a_ready[gc & 1] : &a_full[gc & 1], par); DBG_ADD(w_a); }
                    tc_fence_after();
                    const uint32_t aHi16 = (smem_u32(smem + (size_t)slot_x(gc, nbuf_k) * slot_bytes) >> 4) | lbo_bits;
                    const uint32_t aLo16 = (smem_u32(smem + (size_t)slot_l(gc, nbuf_k) * slot_bytes) >> 4) | lbo_bits;
                    const uint32_t p16_base = aLo16;                 // mode 2: the "lo" slot holds the two bf16 tiles
                    uint32_t row16 = 0;                               // (ky * pitch) * 8
                    int kx = 0;
                    for (int t = 0; t < ntaps; ++t) {
                        const uint32_t woff16 = row16 + (uint32_t)(kx * 8);
                        if (++kx == a.KW) { kx = 0; row16 += pitch8; }
                        if (a.s2 && !s2_tap_used(t, c / a.cpp)) continue;
                        const int sub = a.tps > 1 ? t % a.tps : 0;
                        if (sub == 0) {
                            { DBG_T0(); mbar_wait(&b_full[bs], bphase); DBG_ADD(w_b); }
                            tc_fence_after();
                        }
                        const uint32_t bHi16 = (smem_u32(smemB + ((size_t)bs * a.tps + sub) * b_stage) >> 4) | lbo_bits;
                        uint32_t a_hi = aHi16 + woff16 + issuer * tile16, a_lo = aLo16 + woff16 + issuer * tile16;
                        uint32_t t_main = tset + (uint32_t)issuer * tile_cols;
                        for (int i = issuer; i < a.NT; i += n_issuers) {
                            if (F16) {
                                // [f16(a) | f16(a_lo')] in the pair slot; [f16(w) | f16(w_lo')] in the weight stage are ONE
                                // K-major tile of 2*CoutP rows, so a_hi * [w_hi | w_lo'] is a single MMA of N = 2*CoutP that
                                // fills [main | corr] (the A tile is read from shared memory once instead of twice --
                                // shared-memory bandwidth, not the tensor pipe, bounds these layers), then a_lo' * w_hi -> corr
                                const uint32_t pa16 = p16_base + woff16 / 2 + (uint32_t)i * (tile16 / 2);
#pragma unroll
                                for (int kk = 0; kk < KC / 16; ++kk) {
                                    if (2 * kk < nk) {
                                        const uint32_t first = acc | (uint32_t)(kk > 0);
                                        if (F16S) {
                                            umma_bf16_lohi(t_main, pa16 + 2 * kk, hiA16, bHi16 + 2 * kk, hiB16, idesc16, first);
                                            umma_bf16_lohi(t_main, pa16 + 2 * kk, hiA16, bHi16 + blo16 / 2 + 2 * kk, hiB16, idesc16, 1);
                                            umma_bf16_lohi(t_main, pa16 + half16 + 2 * kk, hiA16, bHi16 + blo16 + 2 * kk, hiB16, idesc16, 1);
                                        } else {
                                        umma_bf16_lohi(t_main, pa16 + 2 * kk, hiA16, bHi16 + 2 * kk, hiB16, idesc16w, first);
                                        umma_bf16_lohi(t_main + (uint32_t)a.CoutP, pa16 + half16 + 2 * kk, hiA16, bHi16 + 2 * kk, hiB16, idesc16, 1);
                                        }
                                    }
                                }
                            } else {
#pragma unroll
                            for (int k = 0; k < KC / 8; ++k) {
                                if (k < nk) {
                                    const uint32_t first = acc | (uint32_t)(k > 0);
                                    umma_tf32_lohi(t_main, a_hi + 2 * k, hiA, bHi16 + 2 * k, hiB, idesc, first);
                                    if (PASSES == 3) {
                                        umma_tf32_lohi(t_main + corr_off, a_lo + 2 * k, hiA, bHi16 + 2 * k, hiB, idesc, a.corr ? first : 1u);
                                        umma_tf32_lohi(t_main + corr_off, a_hi + 2 * k, hiA, bHi16 + blo16 + 2 * k, hiB, idesc, 1);
                                    }
                                }
                            }
                            }
                            if (PASSES == 2) {
                                // bf16 tiles: 64-byte rows (32 channels), 64B swizzle; [bf16(a) | bf16(a_lo)] in the pair slot,
                                // [bf16(w) | bf16(w_lo)] behind the fp32 weights of the stage
                                const uint32_t pa16 = p16_base + woff16 / 2 + (uint32_t)i * (tile16 / 2);
                                const uint32_t b16 = bHi16 + blo16;
#pragma unroll
                                for (int kk = 0; kk < KC / 16; ++kk) {
                                    if (2 * kk < nk) {
                                        const uint32_t firstc = a.corr ? (acc | (uint32_t)(kk > 0)) : 1u;
                                        // a_lo * w
                                        umma_bf16_lohi(t_main + corr_off, pa16 + half16 + 2 * kk, hiA16, b16 + 2 * kk, hiB16, idesc16, firstc);
                                        // a * w_lo
                                        umma_bf16_lohi(t_main + corr_off, pa16 + 2 * kk, hiA16, b16 + blo16 / 2 + 2 * kk, hiB16, idesc16, 1);
                                    }
                                }
                            }
                            a_hi += n_issuers * tile16; a_lo += n_issuers * tile16; t_main += (uint32_t)n_issuers * tile_cols;
                        }
                        acc = 1;
                        if (sub == a.tps - 1) {
                            umma_commit(&b_empty[bs]);
                            if (++bs == a.nB) { bs = 0; bphase ^= 1; }
                        }
                    }
                    umma_commit(&chunk_done[gc & 1]);
                }
                umma_commit(&acc_full[as]);
            }
            if (a.dbg && blockIdx.x == 0 && issuer == 0) {
                a.dbg[0] = clock64() - t_begin; a.dbg[1] = w_acc; a.dbg[2] = w_a; a.dbg[3] = w_b; a.dbg[4] = wl;
            }
            }
        }
    } else if (warp >= 8 && warp < 16) {
        // ================================ 3xTF32 split warps ============================
        if (SPLIT) {
            const int et = threadIdx.x - 8 * 32;
            const int nvec = halo_bytes / 16;
            int gc = 0;
            long long s_w1 = 0, s_w2 = 0, s_busy = 0, s_lat = 0;
            for (int w = blockIdx.x; w < a.total; w += G) {
                for (int c = 0; c < nchunk; ++c, ++gc) {
                    // the lo slot of chunk gc was last in use by chunk gc-1's MMAs (3 slots) or chunk gc-2's (4 slots)
                    {
                        DBG_T0();
                        const int dep = gc - ((a.nBuf == 4 || a.slot_mode) ? 2 : 1);
                        if (dep >= 0) mbar_wait(&chunk_done[dep & 1], (uint32_t)((dep >> 1) & 1));
                        DBG_ADD(s_w1);
                    }
                    { DBG_T0(); mbar_wait(&a_full[gc & 1], (uint32_t)((gc >> 1) & 1)); DBG_ADD(s_w2); }
                    if (a.dbg) s_lat += clock64() - *(volatile long long*)&dbg_a_issue[gc & 1];
                    const long long s_t0 = a.dbg ? clock64() : 0;
                    float4* pa = reinterpret_cast<float4*>(smem + (size_t)slot_x(gc, nbuf_k) * slot_bytes);
                    float4* pl = reinterpret_cast<float4*>(smem + (size_t)slot_l(gc, nbuf_k) * slot_bytes);
                    // all loads of a batch are issued before the first use (latency-bound otherwise: the MMAs wait
                    // for this between two chunks)
                    constexpr int SB = 6;
                    if (PASSES == 2 || F16) {
                        // bf16(a) and bf16(a - trunc_tf32(a)) (mode 4: f16(a) and f16((a - f16(a)) * 2^11)) into the pair slot, 64-byte rows with the 64B swizzle applied
                        // by hand (absolute shared-memory address bits [7:8] -> [4:5], the rule TMA / UMMA use)
                        // Slots are 1024-byte aligned and both halves of the pair slot start on a 512-byte boundary, so the
                        // two swizzles compose to something simple.  Pair q = 16-byte chunks 2q, 2q+1 of the fp32 tile
                        // (128B swizzle: logical chunk = p ^ (p >> 3 & 7)); with x = (q >> 2) & 1 chunk 2q + x is the even
                        // logical chunk 2L (channels 8L .. 8L+3, L = (q & 3) ^ (q >> 3 & 3)) and its neighbour the odd one.
                        // The pair row of pixel q >> 2 wants logical 16-byte chunk L at physical chunk L ^ (q >> 3 & 3)
                        // = q & 3 (64B swizzle): the stores are linear in q.  Bonus: the x term makes the 8 lanes of a
                        // quarter warp hit 8 different bank groups on the loads.
                        const uint32_t half_b = ((uint32_t)(halo_rows * pitch * 64) + 511u) & ~511u;
                        const uint32_t pa_s = smem_u32(pa), ph_s = smem_u32(pl), plo_s = ph_s + half_b;
                        // one thread = two adjacent 16-byte chunks of a row = 8 consecutive channels -> one 16-byte store per half
                        const int npair = nvec >> 1;
                        constexpr int SP2 = 4;
                        float amax = 0.f;
                        for (int base = et; base < npair; base += SP2 * SPLIT_THREADS) {
                            float4 v0[SP2], v1[SP2];
#pragma unroll
                            for (int j = 0; j < SP2; ++j) {
                                const int q = base + j * SPLIT_THREADS;
                                if (q < npair) { const uint32_t p = (uint32_t)(2 * q + ((q >> 2) & 1)); v0[j] = lds128(pa_s + p * 16u); v1[j] = lds128(pa_s + (p ^ 1u) * 16u); }
                            }
#pragma unroll
                            for (int j = 0; j < SP2; ++j) {
                                const int q = base + j * SPLIT_THREADS;
                                if (q < npair) {
                                    const float4 a0 = v0[j], a1 = v1[j];
                                    uint4 h, l;
                                    if (F16) {
                                        h.x = pack_f16(a0.x, a0.y); h.y = pack_f16(a0.z, a0.w);
                                        h.z = pack_f16(a1.x, a1.y); h.w = pack_f16(a1.z, a1.w);
                                        l.x = pack_f16_lo(a0.x, a0.y, h.x); l.y = pack_f16_lo(a0.z, a0.w, h.y);
                                        l.z = pack_f16_lo(a1.x, a1.y, h.z); l.w = pack_f16_lo(a1.z, a1.w, h.w);
                                        amax = fmaxf(fmaxf(fmaxf(amax, fabsf(a0.x)), fmaxf(fabsf(a0.y), fabsf(a0.z))),
                                                     fmaxf(fmaxf(fabsf(a0.w), fabsf(a1.x)), fmaxf(fabsf(a1.y), fmaxf(fabsf(a1.z), fabsf(a1.w)))));
                                    } else {
                                    h.x = pack_bf16(a0.x, a0.y); h.y = pack_bf16(a0.z, a0.w);
                                    h.z = pack_bf16(a1.x, a1.y); h.w = pack_bf16(a1.z, a1.w);
#define PIVLFN_LO(f) ((f) - __uint_as_float(__float_as_uint(f) & 0xFFFFE000u))
                                    l.x = pack_bf16(PIVLFN_LO(a0.x), PIVLFN_LO(a0.y)); l.y = pack_bf16(PIVLFN_LO(a0.z), PIVLFN_LO(a0.w));
                                    l.z = pack_bf16(PIVLFN_LO(a1.x), PIVLFN_LO(a1.y)); l.w = pack_bf16(PIVLFN_LO(a1.z), PIVLFN_LO(a1.w));
#undef PIVLFN_LO
                                    }
                                    sts128(ph_s + (uint32_t)q * 16u, h);
                                    sts128(plo_s + (uint32_t)q * 16u, l);
                                }
                            }
                        }
                        // an activation that rounds to inf in fp16 (|x| >= 65520) left the range of this mode
                        if (F16 && amax >= 65520.f) g_f16_range_flag = 1;
                    } else
                    for (int base = et; base < nvec; base += SB * SPLIT_THREADS) {
                        float4 v[SB];
#pragma unroll
                        for (int j = 0; j < SB; ++j)
                            if (base + j * SPLIT_THREADS < nvec) v[j] = pa[base + j * SPLIT_THREADS];
#pragma unroll
                        for (int j = 0; j < SB; ++j) {
                            const int idx = base + j * SPLIT_THREADS;
                            if (idx < nvec) {
                                float4 h, l;
                                if (a.split_trunc) {
                                    h.x = __uint_as_float(__float_as_uint(v[j].x) & 0xFFFFE000u);
                                    h.y = __uint_as_float(__float_as_uint(v[j].y) & 0xFFFFE000u);
                                    h.z = __uint_as_float(__float_as_uint(v[j].z) & 0xFFFFE000u);
                                    h.w = __uint_as_float(__float_as_uint(v[j].w) & 0xFFFFE000u);
                                } else {
                                    h.x = __uint_as_float((__float_as_uint(v[j].x) + 0x1000u) & 0xFFFFE000u);
                                    h.y = __uint_as_float((__float_as_uint(v[j].y) + 0x1000u) & 0xFFFFE000u);
                                    h.z = __uint_as_float((__float_as_uint(v[j].z) + 0x1000u) & 0xFFFFE000u);
                                    h.w = __uint_as_float((__float_as_uint(v[j].w) + 0x1000u) & 0xFFFFE000u);
                                    pa[idx] = h;
                                }
                                l.x = v[j].x - h.x; l.y = v[j].y - h.y; l.z = v[j].z - h.z; l.w = v[j].w - h.w;
                                pl[idx] = l;
                            }
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_ready[gc & 1]);
                    if (a.dbg) s_busy += clock64() - s_t0;
                }
            }
            if (a.dbg && blockIdx.x == 0 && threadIdx.x == 8 * 32) { a.dbg[9] = s_w1; a.dbg[10] = s_w2; a.dbg[11] = s_busy; a.dbg[14] = s_lat; }
        }
    } else if (warp >= 4) {
        // ================================ epilogue warps ================================
        // two warps per TMEM lane quarter (warps 4-7: first half of the output channels, 16-19: second half)
        const int egrp = warp >= 16 ? 1 : 0;
        const bool halves = (a.CoutP % 32) == 0;                        // both halves must be multiples of 16 columns
        const int chalf = halves ? a.CoutP / 2 : a.CoutP;
        const int cbeg = halves ? egrp * chalf : 0;
        const int cend = halves ? cbeg + chalf : (egrp == 0 ? a.CoutP : 0);
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const float osc = F16S ? (1.f / 256.f) : 1.f;        // mode 5: the weights carry a factor 2^8 (fmaf(x, 1, b) == x + b)
        int wl = 0;
        uint32_t p16_bad = 0;
        long long e_wait = 0, e_busy = 0, e_ld = 0;
        for (int w = blockIdx.x; w < a.total; w += G, ++wl) {
            const int as = wl % a.nsets;
            const uint32_t use = (uint32_t)(wl / a.nsets);
            const int tx = w % a.tiles_x, ty = (w / a.tiles_x) % a.tiles_y, n = w / (a.tiles_x * a.tiles_y);
            const int x = tx * HT_W + (row & (HT_W - 1));
            { DBG_T0(); mbar_wait(&acc_full[as], use & 1); DBG_ADD(e_wait); }
            const long long e_t0 = a.dbg ? clock64() : 0;
            tc_fence_after();
            const uint32_t trow = tmem_base + (uint32_t)(as * set_cols) + ((uint32_t)(q * 32) << 16);
            uint32_t v[16], u[16];
            bool pre = false;                 // the (main, corr) columns of this group are already on their way (see below)
            for (int i = 0; i < a.NT; ++i) {
                const int yy = (ty * a.NT + i) * HT_H + row / HT_W;
                const bool live = x < a.W && yy < a.H && a.vec_store != 3;      // vec_store 3: experiment, no stores
                const size_t pix = ((size_t)n * a.H + yy) * a.W + x;
                float* dst = a.y + pix * a.y_ld;
                const float* rsd = a.res ? a.res + pix * a.res_ld : nullptr;
                // two 16-column TMEM loads in flight per wait: (main, corr) of the same columns in 3xTF32+corr mode,
                // otherwise two consecutive column groups
                const bool dual = (SPLIT && a.corr);
                const int cstep = dual ? 16 : 32;
                for (int c0 = cbeg; c0 < cend; c0 += cstep) {
                    const bool second = dual || (c0 + 16 < cend);
                    const long long l_t0 = a.dbg ? clock64() : 0;
                    if (!pre) {
                        tmem_ld16_nowait(trow + (uint32_t)((F16D ? 2 * i : i) * a.CoutP + c0), v);
                        if (second)
                            tmem_ld16_nowait(trow + (uint32_t)(F16D ? (2 * i + 1) * a.CoutP + c0
                                                                   : (dual ? (a.NT + i) * a.CoutP + c0 : i * a.CoutP + c0 + 16)), u);
                    }
                    pre = false;
                    tmem_ld_wait();
                    if (a.dbg) e_ld += clock64() - l_t0;
                    if (dual) {
                        const float cs = F16D ? (1.f / 2048.f) : 1.f;    // mode 4 keeps the low-order products scaled by 2^11
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(u[j]), cs, __uint_as_float(v[j])));
                    }
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if (half == 1 && (dual || !second)) break;
                        const int cb = c0 + half * 16;
                        if (a.vec_store == 5 && cb + 16 <= a.cout_st) {
                            // TMA tile store.  Direct st.global from the epilogue threads was measured to be a per-SM
                            // limit of ~18 B/clk (7k of the 10k epilogue cycles of a 128 KB work item, whatever the
                            // access pattern), and with one TMEM accumulator set the tensor cores idle meanwhile.  Here
                            // each warp parks its 32 pixels x 16 channels in a 2 KB staging tile (64B swizzle: conflict-free
                            // float4 stores) and one lane hands it to the TMA engine; the accumulators are released as soon
                            // as the last tile is parked, the global writes drain in the background.
                            const int ew = (warp & 3) + 4 * egrp;
                            uint8_t* stg = smem + a.stage_off + ew * 2048;
                            float4 t4[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float e[4];
                                const float4 b4 = *reinterpret_cast<const float4*>(&bias_s[cb + 4 * j]);      // one LDS.128, not four LDS
                                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    float t = fmaf(__uint_as_float(half ? u[4 * j + k] : v[4 * j + k]), osc, bb[k]);
                                    e[k] = a.lrelu ? lrelu_f(t) : t;
                                }
                                t4[j] = make_float4(e[0], e[1], e[2], e[3]);
                            }
                            if (lane == 0) bulk_wait_read0();          // the previous store has finished reading the tile
                            __syncwarp();
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                *reinterpret_cast<float4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = t4[j];
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_4d(&tmY, stg, cb, tx * HT_W, (ty * a.NT + i) * HT_H + 4 * q, n);
                                bulk_commit();
                            }
                        } else if (a.vec_store == 4 && cb + 16 <= a.cout_st) {
                            // Quad-transposed stores.  A thread owns one pixel's 16 channels (64 B); written directly, every
                            // store instruction touches 32 different 128-byte lines with one sector each, and these stores
                            // were measured to cost 7k cycles per 128 KB work item (the LSU/L1 path they share with the
                            // split warps, not HBM).  After a 4x4 transpose of float4 chunks inside each lane quad (4
                            // consecutive pixels of a tile row), lane j holds chunk j of all four pixels, so one store
                            // instruction writes 64 contiguous bytes per quad: 8 lines per instruction instead of 32.
                            float4 t4[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float e[4];
                                const float4 b4 = *reinterpret_cast<const float4*>(&bias_s[cb + 4 * j]);      // one LDS.128, not four LDS
                                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    float t = fmaf(__uint_as_float(half ? u[4 * j + k] : v[4 * j + k]), osc, bb[k]);
                                    e[k] = a.lrelu ? lrelu_f(t) : t;
                                }
                                t4[j] = make_float4(e[0], e[1], e[2], e[3]);
                            }
                            if (a.out_p16) {
                                const float r[16] = {t4[0].x, t4[0].y, t4[0].z, t4[0].w, t4[1].x, t4[1].y, t4[1].z, t4[1].w,
                                                     t4[2].x, t4[2].y, t4[2].z, t4[2].w, t4[3].x, t4[3].y, t4[3].z, t4[3].w};
                                uint4 e0, e1, e2, e3;
                                p16::encode16(r, e0, e1, e2, e3);
                                p16_bad |= p16::nonfinite_bits(e0) | p16::nonfinite_bits(e1);
                                t4[0] = make_float4(__uint_as_float(e0.x), __uint_as_float(e0.y), __uint_as_float(e0.z), __uint_as_float(e0.w));
                                t4[1] = make_float4(__uint_as_float(e1.x), __uint_as_float(e1.y), __uint_as_float(e1.z), __uint_as_float(e1.w));
                                t4[2] = make_float4(__uint_as_float(e2.x), __uint_as_float(e2.y), __uint_as_float(e2.z), __uint_as_float(e2.w));
                                t4[3] = make_float4(__uint_as_float(e3.x), __uint_as_float(e3.y), __uint_as_float(e3.z), __uint_as_float(e3.w));
                            }
                            if (dual) {
                                // v / u are consumed: start the TMEM loads of the next 16-column group (next stacked tile
                                // after the last group) now, so that their latency hides behind the transpose and the stores
                                int ni = i, nc = c0 + 16;
                                if (nc >= cend) { nc = cbeg; ++ni; }
                                if (ni < a.NT) {
                                    tmem_ld16_nowait(trow + (uint32_t)((F16D ? 2 * ni : ni) * a.CoutP + nc), v);
                                    tmem_ld16_nowait(trow + (uint32_t)(F16D ? (2 * ni + 1) * a.CoutP + nc : (a.NT + ni) * a.CoutP + nc), u);
                                    pre = true;
                                }
                            }
#pragma unroll
                            for (int sft = 2; sft >= 1; sft >>= 1) {
                                const bool up = (lane & sft) != 0;
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    if (i & sft) continue;
                                    // lanes with the bit clear keep t4[i] and trade t4[i ^ sft]; the others the opposite
                                    float4 snd = up ? t4[i] : t4[i ^ sft];
                                    float4 rcv;
                                    rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, sft);
                                    rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, sft);
                                    rcv.z = __shfl_xor_sync(0xffffffffu, snd.z, sft);
                                    rcv.w = __shfl_xor_sync(0xffffffffu, snd.w, sft);
                                    if (up) t4[i] = rcv; else t4[i ^ sft] = rcv;
                                }
                            }
                            const int lq = lane & 3;
                            if (yy < a.H && x - lq < a.W) {          // W % 4 == 0: a quad is live or dead as a whole
                                float* qb = a.y + (pix - lq) * a.y_ld + cb + lq * 4;
#pragma unroll
                                for (int r = 0; r < 4; ++r) *reinterpret_cast<float4*>(qb + (size_t)r * a.y_ld) = t4[r];
                            }
                        } else if (live) {
                            float o[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                float t = fmaf(__uint_as_float(half ? u[j] : v[j]), osc, bias_s[cb + j]);   // (slow path: scalar LDS)
                                if (a.lrelu) t = lrelu_f(t);
                                o[j] = t;
                            }
                            if (a.planar) {
#pragma unroll
                                for (int j = 0; j < 16; j += 2)
                                    if (cb + j < a.Cout)
                                        *reinterpret_cast<float2*>(a.y + (size_t)((cb + j) >> 1) * a.planar + pix * 2) =
                                            make_float2(o[j], o[j + 1]);
                            } else if (a.vec_store == 2 && cb + 16 <= a.cout_st && !rsd) {
                                st_global_v8(dst + cb, o);
                                st_global_v8(dst + cb + 8, o + 8);
                            } else if (a.vec_store) {
#pragma unroll
                                for (int j = 0; j < 16; j += 4) {
                                    if (cb + j + 4 <= a.cout_st) {
                                        float4 w4 = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                                        if (rsd) { w4.x += rsd[cb + j]; w4.y += rsd[cb + j + 1]; w4.z += rsd[cb + j + 2]; w4.w += rsd[cb + j + 3]; }
                                        *reinterpret_cast<float4*>(dst + cb + j) = w4;
                                    } else {
#pragma unroll
                                        for (int jj = j; jj < j + 4; ++jj)
                                            if (cb + jj < a.Cout) dst[cb + jj] = o[jj] + (rsd ? rsd[cb + jj] : 0.f);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (cb + j < a.Cout) dst[cb + j] = o[j] + (rsd ? rsd[cb + j] : 0.f);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
            if (a.dbg) e_busy += clock64() - e_t0;
        }
        if (a.out_p16 && a.range_flag && p16::any_nonfinite(p16_bad)) *a.range_flag = 1;
        if (a.vec_store == 5 && lane == 0) bulk_wait0();      // all tile stores of this warp have completed
        if (a.dbg && blockIdx.x == 0 && warp == 4 && lane == 0) { a.dbg[7] = e_wait; a.dbg[8] = e_busy; a.dbg[12] = e_ld; }
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

int num_sms() { return pivlfn_num_sms(); }

int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

constexpr int SMEM_BUDGET = 200 * 1024;
constexpr int HALO_SMEM_BUDGET = 226 * 1024;   // 227 KB opt-in per CTA minus the static part (padded to 1 KB by the
                                               // 1024-byte alignment of the dynamic array, which also makes align-up slack unnecessary)

template <int PASSES>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo, ConvTcArgs& a, int grid,
           cudaStream_t st) {
    const int stage_bytes = (PASSES == 3 ? 2 : 1) * (A_BYTES + a.CoutP * KC * 4);
    int stages = (SMEM_BUDGET - 1024) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    const int niter = a.ntaps * ((a.Cin + KC - 1) / KC);
    if (stages > niter) stages = niter;
    if (stages < 2 && niter >= 2) return PIVLFN_EUNSUPPORTED;
    a.stages = stages;
    const int smem = stages * stage_bytes + 1024;
    auto kern = conv_tc_kernel<PASSES>;
    static unsigned long long configured = 0;
    {
        cudaError_t e = pivlfn_optin_smem(kern, SMEM_BUDGET, configured);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<grid, NTHREADS, smem, st>>>(tmA, tmBhi, tmBlo, a);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

// weights [CoutP][ntaps][CinP] -> box (32 channels, 1 tap, CoutP rows)
int encode_weights(EncodeTiledFn enc, CUtensorMap* tm, const float* w, int CinP, int ntaps, int CoutP) {
    cuuint64_t dims[3] = {(cuuint64_t)CinP, (cuuint64_t)ntaps, (cuuint64_t)CoutP};
    cuuint64_t strides[2] = {(cuuint64_t)CinP * 4, (cuuint64_t)ntaps * CinP * 4};
    cuuint32_t box[3] = {(cuuint32_t)KC, 1, (cuuint32_t)CoutP};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PIVLFN_OK : PIVLFN_EINVAL;
}

// bf16 weights [CoutP][ntaps][CinP] -> box (32 channels = 64 bytes, 1 tap, CoutP rows), 64B swizzle
int encode_weights_bf16(EncodeTiledFn enc, CUtensorMap* tm, const void* w, int CinP, int ntaps, int CoutP) {
    cuuint64_t dims[3] = {(cuuint64_t)CinP, (cuuint64_t)ntaps, (cuuint64_t)CoutP};
    cuuint64_t strides[2] = {(cuuint64_t)CinP * 2, (cuuint64_t)ntaps * CinP * 2};
    cuuint32_t box[3] = {(cuuint32_t)KC, 1, (cuuint32_t)CoutP};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PIVLFN_OK : PIVLFN_EINVAL;
}

// All 16-bit weight tiles of one ring stage in ONE TMA instruction (the producer thread is blocked ~250 cycles per tensor
// load whatever its size, and a thin layer's chunk has 18 of them otherwise): the pack [parts][CoutP][ntaps][CinP] seen as
// (channel, cout, part, tap) -> box (32 channels = 64 bytes, CoutP, parts, tps taps) lands in shared memory as
// [tap][part][cout][64 B], i.e. per tap the [hi | lo (| hi2)] tiles the MMAs expect.
int encode_weights16_stage(EncodeTiledFn enc, CUtensorMap* tm, const void* w, int CinP, int ntaps, int CoutP, int parts, int tps) {
    cuuint64_t dims[4] = {(cuuint64_t)CinP, (cuuint64_t)CoutP, (cuuint64_t)parts, (cuuint64_t)ntaps};
    cuuint64_t strides[3] = {(cuuint64_t)ntaps * CinP * 2, (cuuint64_t)CoutP * ntaps * CinP * 2, (cuuint64_t)CinP * 2};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)CoutP, (cuuint32_t)parts, (cuuint32_t)tps};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PIVLFN_OK : PIVLFN_EINVAL;
}

void choose_tile(ConvTcArgs& a) {
    a.bw = pow2_ceil(a.W < 16 ? a.W : 16);
    const int rem = TILE_M / a.bw;
    a.bh = pow2_ceil(a.H < rem ? a.H : rem);
    a.bn = TILE_M / (a.bw * a.bh);
    a.tiles_x = cdiv(a.W, a.bw);
    a.tiles_y = cdiv(a.H, a.bh);
}

// Environment switches of the halo kernel (experiments; defaults are the measured best).
struct HaloEnv { int use_halo, bo_mode, nt_limit, corr_mode, split_trunc, stagger, quad_store, tma_store, tps3; };
const HaloEnv& halo_env() {
    static HaloEnv e = {-1, 0, 0, 1, 0, 0, 1, 0, 1};
    if (e.use_halo < 0) {
        const char* v = getenv("PIVLFN_TC_HALO");
        e.use_halo = (v && v[0] == '0') ? 0 : 1;
#ifdef PIVLFN_DEBUG      // timing experiments (build with -DPIVLFN_DEBUG): not available in the product library
        v = getenv("PIVLFN_TC_BO");
        e.bo_mode = (v && v[0] == '1') ? 1 : 0;
        v = getenv("PIVLFN_TC_STAGGER");          // percent of one estimated work-item period; 0 = off
        e.stagger = v ? atoi(v) : 0;              // measured: no effect (the stores are not HBM-bound), off by default
#endif
        v = getenv("PIVLFN_TC_NT");
        e.nt_limit = v ? atoi(v) : 0;
        v = getenv("PIVLFN_TC_CORR");
        e.corr_mode = (v && v[0] == '0') ? 0 : 1;
        v = getenv("PIVLFN_TC_SPLIT_TRUNC");
        e.split_trunc = (v && v[0] == '0') ? 0 : 1;
        v = getenv("PIVLFN_TC_QUADSTORE");
        e.quad_store = (v && v[0] == '0') ? 0 : 1;
        v = getenv("PIVLFN_TC_TPS3");
        e.tps3 = (v && v[0] == '0') ? 0 : 1;
        v = getenv("PIVLFN_TC_TMASTORE");         // measured: no faster than direct stores (the limit is downstream of the SM,
        e.tma_store = (v && v[0] == '1') ? 1 : 0; // ~18 B/clk per SM either way) and it costs a weight-ring stage: off by default
    }
    return e;
}

// Choose NT / buffers / ring depth for the halo kernel; fills h and returns the dynamic shared memory size, or 0 when
// no configuration fits (caller falls back to the per-tap kernel).
constexpr int EPI_STAGE_BYTES = 8 * 2048;      // TMA-store staging: 2 KB per epilogue warp

int halo_configure(ConvHaloArgs& h, int passes, int* halo_rows_out) {
    const HaloEnv& env = halo_env();
    const int pitch = HT_W + h.KW - 1;
    const int b_stage = passes == 5 ? h.CoutP * KC * 6 : (passes == 4 ? 1 : (passes >= 2 ? 2 : 1)) * h.CoutP * KC * 4;
    // mode 4 scales its corrections: own accumulator; mode 5: everything in one accumulator
    const int corr = passes == 5 ? 0 : ((passes == 4 || (passes >= 2 && env.corr_mode)) ? 1 : 0);
    const int acc_mult = corr ? 2 : 1;
    // NT stacked tiles per work item: bounded by TMEM (512 columns, two accumulator sets wanted so that the epilogue
    // overlaps the next item's MMAs), by the image height and by shared memory
    int NT = env.nt_limit > 0 ? env.nt_limit : 2;
    if (NT > MAX_NT) NT = MAX_NT;
    while (NT > 1 && (acc_mult * NT * h.CoutP > 512 || HT_H * (NT - 1) >= h.H)) --NT;
    for (; NT >= 1; --NT) {
        if (NT == 3) continue;
        const int halo_rows = HT_H * NT + h.KH - 1;
        const int slot = (halo_rows * pitch * 128 + 1023) & ~1023;
        int nBuf = passes >= 2 ? 3 : 2;
        // a fourth slot (double-buffered (x, lo) pairs: no split bubble between chunks) when it still leaves a 3-deep
        // weight ring
        const int stage = h.vec_store == 5 ? EPI_STAGE_BYTES : 0;
        if (passes >= 2 && (HALO_SMEM_BUDGET - stage - 4 * slot) / b_stage >= (stage ? 2 : 3)) nBuf = 4;
        // three taps per weight stage for small Cout (see ConvHaloArgs::tps) when two such stages still fit
        int tps = 1;
        if (env.tps3 && passes >= 4 && !h.s2 && h.CoutP <= 64 && (h.KH * h.KW) % 3 == 0 &&
            (HALO_SMEM_BUDGET - stage - nBuf * slot) / (3 * b_stage) >= 2) tps = 3;
        int nB = (HALO_SMEM_BUDGET - stage - nBuf * slot) / (tps * b_stage);
        if (nB > MAX_STAGES) nB = MAX_STAGES;
        const int need = passes >= 2 ? 2 : 3;
        if (nB < need || halo_rows > 256) continue;
        h.corr = corr;
        h.nsets = (2 * acc_mult * NT * h.CoutP <= 512) ? 2 : 1;
        h.NT = NT; h.nBuf = nBuf; h.nB = nB; h.tps = tps;
        h.slot_mode = (passes >= 4 && nBuf == 3) ? 1 : 0;
        h.tiles_x = cdiv(h.W, HT_W); h.tiles_y = cdiv(h.H, HT_H * NT);
        const long long total = (long long)h.tiles_x * h.tiles_y * h.N;
        if (total > 0x7FFFFFFFLL) return 0;
        h.total = (int)total;
        h.bo_mode = env.bo_mode; h.split_trunc = env.split_trunc;
        {
            // estimated tensor-core cycles of one work item: M = 128 MMAs take N/2 cycles per 32 bytes of K
            const int nchunk = (h.Cin + KC - 1) / KC;
            const int per_chunk_tap = passes >= 4 ? 2 * 3 : (passes == 3 ? 4 * 3 : (passes == 2 ? 4 + 4 : 4));
            const long long item = (long long)h.KH * h.KW * nchunk * NT * per_chunk_tap * (h.CoutP / 2);
            const int n_cta = h.total < num_sms() ? h.total : num_sms();
            h.stagger = (h.nsets == 1 && h.total >= 4 * n_cta) ? (int)(item * env.stagger / 100) : 0;
        }
        *halo_rows_out = halo_rows;
        h.stage_off = (nBuf * slot + nB * tps * b_stage + 1023) & ~1023;
        return stage ? h.stage_off + stage : nBuf * slot + nB * tps * b_stage;
    }
    return 0;
}

int halo_launch(const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo, const CUtensorMap& tmB16,
                const CUtensorMap& tmBlo16, const ConvHaloArgs& h, int passes, int smem, cudaStream_t st,
                const CUtensorMap* tmYp = nullptr) {
    const CUtensorMap& tmY = tmYp ? *tmYp : tmA;
    int grid = h.total < num_sms() ? h.total : num_sms();
#ifdef PIVLFN_DEBUG
    { static int cap = -1; if (cap < 0) { const char* v = getenv("PIVLFN_TC_GRID"); cap = v ? atoi(v) : 0; } if (cap > 0 && cap < grid) grid = cap; }
#endif
    static unsigned long long cfg1 = 0, cfg2 = 0, cfg3 = 0, cfg4 = 0, cfg5 = 0;
    cudaError_t e = cudaSuccess;
    if (passes == 5) {
        e = pivlfn_optin_smem(conv_tc_halo_kernel<5>, HALO_SMEM_BUDGET, cfg5); if (e != cudaSuccess) return (int)e;
        conv_tc_halo_kernel<5><<<grid, HALO_THREADS, smem, st>>>(tmA, tmBhi, tmBlo, tmB16, tmBlo16, tmY, h);
    } else if (passes == 4) {
        e = pivlfn_optin_smem(conv_tc_halo_kernel<4>, HALO_SMEM_BUDGET, cfg4); if (e != cudaSuccess) return (int)e;
        conv_tc_halo_kernel<4><<<grid, HALO_THREADS, smem, st>>>(tmA, tmBhi, tmBlo, tmB16, tmBlo16, tmY, h);
    } else if (passes == 3) {
        e = pivlfn_optin_smem(conv_tc_halo_kernel<3>, HALO_SMEM_BUDGET, cfg3); if (e != cudaSuccess) return (int)e;
        conv_tc_halo_kernel<3><<<grid, HALO_THREADS, smem, st>>>(tmA, tmBhi, tmBlo, tmB16, tmBlo16, tmY, h);
    } else if (passes == 2) {
        e = pivlfn_optin_smem(conv_tc_halo_kernel<2>, HALO_SMEM_BUDGET, cfg2); if (e != cudaSuccess) return (int)e;
        conv_tc_halo_kernel<2><<<grid, HALO_THREADS, smem, st>>>(tmA, tmBhi, tmBlo, tmB16, tmBlo16, tmY, h);
    } else {
        e = pivlfn_optin_smem(conv_tc_halo_kernel<1>, HALO_SMEM_BUDGET, cfg1); if (e != cudaSuccess) return (int)e;
        conv_tc_halo_kernel<1><<<grid, HALO_THREADS, smem, st>>>(tmA, tmBhi, tmBlo, tmB16, tmBlo16, tmY, h);
    }
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

}  // namespace

/* Reads (and optionally clears) the sticky fp16-range flag of the passes == 4 convolutions.  Synchronises the device. */
extern "C" int pivlfn_f16_range_flag(int reset) {
    int v = 0;
    if (cudaMemcpyFromSymbol(&v, g_f16_range_flag, sizeof(int)) != cudaSuccess) return -1;
    if (reset && v) { const int z = 0; cudaMemcpyToSymbol(g_f16_range_flag, &z, sizeof(int)); }
    return v;
}

extern "C" int pivlfn_f16_range_flag_clear(void* stream) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_f16_range_flag) != cudaSuccess) return -1;
    return cudaMemsetAsync(p, 0, sizeof(int), (cudaStream_t)stream) == cudaSuccess ? PIVLFN_OK : -1;
}

static long long* g_conv_tc_dbg = nullptr;
/* Debug hook (not part of the public header): device buffer of >= 16 int64 that CTA 0 of the persistent conv kernel
 * fills with per-role wait-cycle totals; NULL switches the instrumentation off. */
extern "C" void pivlfn_debug_set_conv_trace(long long* dev_buf) { g_conv_tc_dbg = dev_buf; }

extern "C" int pivlfn_conv_tc(const float* x, int x_ld, int N, int H, int W, int Cin,
                              const float* w_hi, const float* w_lo, const void* w_c16, const float* bias,
                              float* y, int y_ld, int Cout, int KH, int KW, int stride, int lrelu,
                              const float* res, int res_ld, int passes, void* stream) {
    if (!x || !w_hi || !y || N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return PIVLFN_EINVAL;
    if (passes < 1 || passes > 5) return PIVLFN_EINVAL;
    if (passes >= 2 && !w_lo) return PIVLFN_EINVAL;
    if ((passes == 2 || passes >= 4) && (!w_c16 || ((uintptr_t)w_c16 & 15))) return PIVLFN_EINVAL;
    if (KH < 1 || KW < 1 || !(KH & 1) || !(KW & 1) || KH > 7 || KW > 7) return PIVLFN_EINVAL;
    if (stride != 1 && stride != 2) return PIVLFN_EINVAL;
    if (Cout > 128) return PIVLFN_EUNSUPPORTED;
    if (((uintptr_t)x & 15) || (x_ld & 3) || x_ld < Cin || ((uintptr_t)y & 3) || y_ld < Cout) return PIVLFN_EINVAL;
    if (res && res_ld < Cout) return PIVLFN_EINVAL;
    if (((uintptr_t)w_hi & 15) || (w_lo && ((uintptr_t)w_lo & 15))) return PIVLFN_EINVAL;
    EncodeTiledFn enc = get_encode();
    if (!enc) return PIVLFN_EDRIVER;

    const int CoutP = (Cout + 15) & ~15;
    const int CinP = (Cin + KC - 1) / KC * KC;
    // cout_st: channels the epilogue may store (Cout rounded up to 4 when the rows are exactly that wide and no residual
    // is added).  vec_store 2: 32-byte aligned rows -> 256-bit stores; 1: 16-byte aligned -> 128-bit stores; 0: scalar
    const int cout_st = (!res && y_ld == ((Cout + 3) & ~3)) ? y_ld : Cout;
    const int vec_store = (!((uintptr_t)y & 31) && !(y_ld & 7) && !(cout_st & 7)) ? 2
                        : ((!((uintptr_t)y & 15) && !(y_ld & 3) && cout_st >= 4) ? 1 : 0);
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap tmA, tmBhi, tmBlo;
    if (encode_weights(enc, &tmBhi, w_hi, CinP, KH * KW, CoutP)) return PIVLFN_EINVAL;
    if (passes >= 2) { if (encode_weights(enc, &tmBlo, w_lo, CinP, KH * KW, CoutP)) return PIVLFN_EINVAL; }
    else tmBlo = tmBhi;
    CUtensorMap tmB16 = tmBhi, tmBlo16 = tmBhi;
    CUtensorMap tmB3 = tmBlo;
    if (passes == 2 || passes >= 4) {
        // w_c16 = [bf16(w) | bf16(w - tf32(w))]; mode 4: [f16(w) | f16((w - f16(w)) * 2^11)]; mode 5 (W = 256 w):
        // [f16(W) | f16(W - f16(W)) | f16(f16(W) * 2^-11)]; each tile [CoutP][taps][CinP]
        const size_t half = (size_t)CoutP * KH * KW * CinP * 2;
        if (encode_weights_bf16(enc, &tmB16, w_c16, CinP, KH * KW, CoutP)) return PIVLFN_EINVAL;
        if (encode_weights_bf16(enc, &tmBlo16, (const char*)w_c16 + half, CinP, KH * KW, CoutP)) return PIVLFN_EINVAL;
        if (passes == 5 && encode_weights_bf16(enc, &tmB3, (const char*)w_c16 + 2 * half, CinP, KH * KW, CoutP)) return PIVLFN_EINVAL;
    }

    if (halo_env().use_halo && W >= HT_W && stride == 1) {
        // ---- halo-resident persistent path --------------------------------------------------------------------
        ConvHaloArgs h;
        h.bias = bias; h.res = res; h.res_ld = res_ld; h.y = y; h.y_ld = y_ld;
        h.N = N; h.H = H; h.W = W; h.Cin = Cin; h.Cout = Cout; h.CoutP = CoutP; h.KH = KH; h.KW = KW;
        h.lrelu = lrelu; h.vec_store = vec_store; h.cout_st = cout_st; h.dbg = g_conv_tc_dbg; h.x_shift = -(KW / 2);
        h.planar = 0;
        h.stage_off = 0; h.s2 = 0; h.cpp = 1; h.out_p16 = 0; h.range_flag = nullptr;
        if (vec_store >= 1 && !res && !(W & 3) && !(cout_st & 15) && halo_env().quad_store) h.vec_store = 4;
        // TMA tile stores: rows 16-byte aligned, at least one full 16-channel group, no residual to add
        const bool tma_ok = vec_store >= 1 && !res && cout_st >= 16 && halo_env().tma_store;
        if (tma_ok) h.vec_store = 5;
#ifdef PIVLFN_DEBUG
        if (getenv("PIVLFN_TC_NOSTORE")) h.vec_store = 3;      // timing experiment only: results are not written
#endif
        int halo_rows = 0;
        int smem = halo_configure(h, passes, &halo_rows);
        if (smem <= 0 && h.vec_store == 5) {                   // no room for the staging area: direct stores
            h.vec_store = (!(W & 3) && !(cout_st & 15) && halo_env().quad_store) ? 4 : vec_store;
            smem = halo_configure(h, passes, &halo_rows);
        }
        CUtensorMap tmY;
        if (smem > 0 && h.vec_store == 5) {
            // output [N][H][W][cout_st] with pixel pitch y_ld; box = 16 channels x 8 pixels x 4 rows (one epilogue warp)
            cuuint64_t dims[4] = {(cuuint64_t)cout_st, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
            cuuint64_t strides[3] = {(cuuint64_t)y_ld * 4, (cuuint64_t)W * y_ld * 4, (cuuint64_t)H * W * y_ld * 4};
            cuuint32_t box[4] = {16, (cuuint32_t)HT_W, 4, 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = enc(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, y, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return PIVLFN_EINVAL;
        }
        if (smem > 0) {
            cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
            cuuint64_t strides[3] = {(cuuint64_t)x_ld * 4, (cuuint64_t)W * x_ld * 4, (cuuint64_t)H * W * x_ld * 4};
            cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)(HT_W + KW - 1), (cuuint32_t)halo_rows, 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return PIVLFN_EINVAL;
            h.w_img = passes >= 4 ? (const uint8_t*)w_c16 : nullptr;
            if (passes == 2 && encode_weights16_stage(enc, &tmB16, w_c16, CinP, KH * KW, CoutP, 2, h.tps)) return PIVLFN_EINVAL;
            return halo_launch(tmA, tmBhi, passes == 5 ? tmB3 : tmBlo, tmB16, tmBlo16, h, passes, smem, st, h.vec_store == 5 ? &tmY : nullptr);
        }
    }
    if (passes == 2 || passes >= 4) passes = 3;          // the per-tap kernel (tiny levels) has no bf16-correction variant

    // ---- per-tap kernel: tiny levels, and stride-2 convolutions (the tap's box is fetched with TMA element stride 2) ----
    ConvTcArgs a;
    a.bias = bias; a.res = res; a.res_ld = res_ld; a.y = y; a.y_ld = y_ld;
    const int Ho = (H + 2 * (KH / 2) - KH) / stride + 1, Wo = (W + 2 * (KW / 2) - KW) / stride + 1;
    a.N = N; a.H = Ho; a.W = Wo; a.Cin = Cin; a.Cout = Cout; a.CoutP = CoutP; a.lrelu = lrelu;
    a.KW = KW; a.ntaps = KH * KW; a.ox = -(KW / 2); a.oy = -(KH / 2); a.sx = stride;
    a.vec_store = (vec_store && !(Cout & 3)) ? 1 : 0;
    choose_tile(a);
    const long long grid = (long long)a.tiles_x * a.tiles_y * cdiv(N, a.bn);
    if (grid > 0x7FFFFFFFLL) return PIVLFN_EINVAL;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)x_ld * 4, (cuuint64_t)W * x_ld * 4, (cuuint64_t)H * W * x_ld * 4};
        cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)(a.bw * stride), (cuuint32_t)(a.bh * stride), (cuuint32_t)a.bn};
        cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return PIVLFN_EINVAL;
    }
    return passes == 3 ? launch<3>(tmA, tmBhi, tmBlo, a, (int)grid, st) : launch<1>(tmA, tmBhi, tmBlo, a, (int)grid, st);
}

// NetC.conv1 (src/models.py:70-73): 7x7, 3 -> 32, stride 1, pad 3 on the zero-bordered NHWC4 image buffer
// img_pad [N, H, W+8, 4] (pixel x of the image at column x+4).  One GEMM-K row per filter row: the 8 pixels
// x-3..x+4 of input row y+ky-3 are 32 contiguous floats in memory, fetched through a tensor map whose pixel
// stride (16 B) is smaller than its box row (128 B), i.e. overlapping windows -- no im2col buffer.
// Weights: [32][7][32] with column index kx*4 + c (kx = 7 and c = 3 zero).
extern "C" int pivlfn_conv_stem_tc(const float* img_pad, int N, int H, int W,
                                   const float* w_hi, const float* w_lo, const void* w_c16, const float* bias,
                                   float* y, int y_ld, int lrelu, int passes, void* stream) {
    if (!img_pad || !w_hi || !y || N <= 0 || H <= 0 || W <= 0) return PIVLFN_EINVAL;
    if (passes < 1 || passes > 4) return PIVLFN_EINVAL;
    if (passes >= 2 && !w_lo) return PIVLFN_EINVAL;
    if ((passes == 2 || passes == 4) && !w_c16) return PIVLFN_EINVAL;
    if (((uintptr_t)img_pad & 15) || ((uintptr_t)y & 15) || (y_ld & 3) || y_ld < 32) return PIVLFN_EINVAL;
    EncodeTiledFn enc = get_encode();
    if (!enc) return PIVLFN_EDRIVER;
    if (halo_env().use_halo && W >= HT_W) {
        // one GEMM-K "chunk" of 32 floats = 8 pixels x 4 channels; 7 taps = the 7 filter rows; window column 0 is pixel
        // x - 3 = padded column x + 1
        ConvHaloArgs h;
        h.bias = bias; h.res = nullptr; h.res_ld = 0; h.y = y; h.y_ld = y_ld;
        h.N = N; h.H = H; h.W = W; h.Cin = 32; h.Cout = 32; h.CoutP = 32; h.KH = 7; h.KW = 1;
        h.lrelu = lrelu; h.vec_store = (!((uintptr_t)y & 31) && !(y_ld & 7)) ? 2 : 1; h.cout_st = 32; h.dbg = g_conv_tc_dbg; h.x_shift = 1; h.planar = 0;
        h.stage_off = 0; h.s2 = 0; h.cpp = 1; h.out_p16 = 0; h.range_flag = nullptr;
        int halo_rows = 0;
        const int smem = halo_configure(h, passes, &halo_rows);
        if (smem > 0) {
            CUtensorMap tA, tBhi, tBlo;
            const cuuint64_t Wp = (cuuint64_t)W + 8;
            cuuint64_t dims[4] = {32, (cuuint64_t)W + 1, (cuuint64_t)H, (cuuint64_t)N};
            cuuint64_t strides[3] = {16, Wp * 16, (cuuint64_t)H * Wp * 16};
            cuuint32_t box[4] = {32, (cuuint32_t)HT_W, (cuuint32_t)halo_rows, 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(img_pad), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return PIVLFN_EUNSUPPORTED;
            if (encode_weights(enc, &tBhi, w_hi, 32, 7, 32)) return PIVLFN_EINVAL;
            if (passes >= 2) { if (encode_weights(enc, &tBlo, w_lo, 32, 7, 32)) return PIVLFN_EINVAL; }
            else tBlo = tBhi;
            CUtensorMap tB16 = tBhi, tBlo16 = tBhi;
            h.w_img = passes == 4 ? (const uint8_t*)w_c16 : nullptr;
            if (passes == 2 && encode_weights16_stage(enc, &tB16, w_c16, 32, 7, 32, 2, h.tps)) return PIVLFN_EINVAL;
            return halo_launch(tA, tBhi, tBlo, tB16, tBlo16, h, passes, smem, (cudaStream_t)stream);
        }
    }
    ConvTcArgs a;
    a.bias = bias; a.res = nullptr; a.res_ld = 0; a.y = y; a.y_ld = y_ld;
    a.N = N; a.H = H; a.W = W; a.Cin = 32; a.Cout = 32; a.CoutP = 32; a.lrelu = lrelu;
    a.KW = 1; a.ntaps = 7; a.ox = 1; a.oy = -3; a.vec_store = 1; a.sx = 1;
    choose_tile(a);
    const long long grid = (long long)a.tiles_x * a.tiles_y * cdiv(N, a.bn);
    if (grid > 0x7FFFFFFFLL) return PIVLFN_EINVAL;
    CUtensorMap tmA, tmBhi, tmBlo;
    {
        const cuuint64_t Wp = (cuuint64_t)W + 8;
        cuuint64_t dims[4] = {32, (cuuint64_t)W + 1, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {16, Wp * 16, (cuuint64_t)H * Wp * 16};
        cuuint32_t box[4] = {32, (cuuint32_t)a.bw, (cuuint32_t)a.bh, (cuuint32_t)a.bn};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(img_pad), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return PIVLFN_EUNSUPPORTED;
    }
    if (passes == 2 || passes == 4) passes = 3;
    if (encode_weights(enc, &tmBhi, w_hi, 32, 7, 32)) return PIVLFN_EINVAL;
    if (passes == 3) { if (encode_weights(enc, &tmBlo, w_lo, 32, 7, 32)) return PIVLFN_EINVAL; }
    else tmBlo = tmBhi;
    cudaStream_t st = (cudaStream_t)stream;
    return passes == 3 ? launch<3>(tmA, tmBhi, tmBlo, a, (int)grid, st) : launch<1>(tmA, tmBhi, tmBlo, a, (int)grid, st);
}

// The same stem for the P16 pipeline: fp16-split arithmetic (passes == 4: the image tile is split in shared memory -- the
// image is the one activation that does not exist in P16 form), output written as P16 groups (p16.cuh) for the P16
// consumers NetC.conv2, NetC_ext and moduleFeat.  w_img: stage_image of the [2][32][7][32] fp16 pack.
extern "C" int pivlfn_conv_stem_p16(const float* img_pad, int N, int H, int W, const void* w_img, const float* bias,
                                    void* y, int y_ld, int lrelu, int* range_flag, void* stream) {
    if (!img_pad || !w_img || !y || N <= 0 || H <= 0 || W < HT_W || (W & 3)) return PIVLFN_EINVAL;
    if (((uintptr_t)img_pad & 15) || ((uintptr_t)w_img & 15) || ((uintptr_t)y & 63) || (y_ld & 15) || y_ld < 32) return PIVLFN_EINVAL;
    EncodeTiledFn enc = get_encode();
    if (!enc) return PIVLFN_EDRIVER;
    ConvHaloArgs h;
    h.bias = bias; h.res = nullptr; h.res_ld = 0; h.y = reinterpret_cast<float*>(y); h.y_ld = y_ld;
    h.N = N; h.H = H; h.W = W; h.Cin = 32; h.Cout = 32; h.CoutP = 32; h.KH = 7; h.KW = 1;
    h.lrelu = lrelu; h.vec_store = 4; h.cout_st = 32; h.dbg = g_conv_tc_dbg; h.x_shift = 1; h.planar = 0;
    h.stage_off = 0; h.s2 = 0; h.cpp = 1; h.out_p16 = 1; h.range_flag = range_flag;
    int halo_rows = 0;
    const int smem = halo_configure(h, 4, &halo_rows);
    if (smem <= 0) return PIVLFN_EUNSUPPORTED;
    CUtensorMap tA;
    const cuuint64_t Wp = (cuuint64_t)W + 8;
    cuuint64_t dims[4] = {32, (cuuint64_t)W + 1, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {16, Wp * 16, (cuuint64_t)H * Wp * 16};
    cuuint32_t box[4] = {32, (cuuint32_t)HT_W, (cuuint32_t)halo_rows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(img_pad), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return PIVLFN_EUNSUPPORTED;
    h.w_img = (const uint8_t*)w_img;
    return halo_launch(tA, tA, tA, tA, tA, h, 4, smem, (cudaStream_t)stream);      // fp16 modes read weights through w_img only
}

// Flow heads (last layer of conv_M / conv_S, src/models.py:161,205: KxK, 32 -> 2, K = 7 or 5) restated as
//   D[pixel, tap*2 + co] = sum_c x[pixel, c] * w[co, c, tap]          (a 1x1 convolution to 2*K*K channels: this call)
//   flow[p, co] = bias[co] + res[p, co] + sum_tap D[p + offset(tap), tap*2 + co]        (pivlfn_flow_head_sum)
// The direct KxK form needs K*K * Cin/8 tensor-core instructions per 128 pixels with only 2 of N = 16 columns useful
// (and tcgen05.mma has a ~39-cycle floor per instruction); this form needs Cin/8 instructions at N = 112.
// planes: [K*K][N*H*W][2] fp32, written as channel-pair planes so that the gather of the second step is coalesced.
extern "C" int pivlfn_conv1x1_pairs_tc(const float* x, int x_ld, int N, int H, int W, int Cin,
                                       const float* w_hi, const float* w_lo, const void* w_c16,
                                       float* planes, int npair, int passes, void* stream) {
    if (!x || !w_hi || !planes || N <= 0 || H <= 0 || W < HT_W || Cin <= 0 || npair <= 0 || 2 * npair > 128) return PIVLFN_EINVAL;
    if (passes < 1 || passes > 4) return PIVLFN_EINVAL;
    if (passes >= 2 && !w_lo) return PIVLFN_EINVAL;
    if ((passes == 2 || passes == 4) && !w_c16) return PIVLFN_EINVAL;
    if (((uintptr_t)x & 15) || (x_ld & 3) || x_ld < Cin || ((uintptr_t)planes & 7)) return PIVLFN_EINVAL;
    EncodeTiledFn enc = get_encode();
    if (!enc) return PIVLFN_EDRIVER;
    const int Cout = 2 * npair, CoutP = (Cout + 15) & ~15, CinP = (Cin + KC - 1) / KC * KC;
    CUtensorMap tmA, tmBhi, tmBlo;
    if (encode_weights(enc, &tmBhi, w_hi, CinP, 1, CoutP)) return PIVLFN_EINVAL;
    if (passes >= 2) { if (encode_weights(enc, &tmBlo, w_lo, CinP, 1, CoutP)) return PIVLFN_EINVAL; }
    else tmBlo = tmBhi;
    CUtensorMap tmB16 = tmBhi, tmBlo16 = tmBhi;
    if (passes == 2 || passes == 4) {
        const size_t half = (size_t)CoutP * CinP * 2;
        if (encode_weights_bf16(enc, &tmB16, w_c16, CinP, 1, CoutP)) return PIVLFN_EINVAL;
        if (encode_weights_bf16(enc, &tmBlo16, (const char*)w_c16 + half, CinP, 1, CoutP)) return PIVLFN_EINVAL;
    }
    ConvHaloArgs h;
    h.bias = nullptr; h.res = nullptr; h.res_ld = 0; h.y = planes; h.y_ld = 0;
    h.N = N; h.H = H; h.W = W; h.Cin = Cin; h.Cout = Cout; h.CoutP = CoutP; h.KH = 1; h.KW = 1;
    h.lrelu = 0; h.vec_store = 0; h.cout_st = Cout; h.dbg = g_conv_tc_dbg; h.x_shift = 0;
    h.planar = (long long)N * H * W * 2;
    h.stage_off = 0; h.s2 = 0; h.cpp = 1; h.out_p16 = 0; h.range_flag = nullptr;
    int halo_rows = 0;
    const int smem = halo_configure(h, passes, &halo_rows);
    if (smem <= 0) return PIVLFN_EUNSUPPORTED;
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)x_ld * 4, (cuuint64_t)W * x_ld * 4, (cuuint64_t)H * W * x_ld * 4};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)HT_W, (cuuint32_t)halo_rows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return PIVLFN_EINVAL;
    h.w_img = passes == 4 ? (const uint8_t*)w_c16 : nullptr;
    if (passes == 2 && encode_weights16_stage(enc, &tmB16, w_c16, CinP, 1, CoutP, 2, h.tps)) return PIVLFN_EINVAL;
    return halo_launch(tmA, tmBhi, tmBlo, tmB16, tmBlo16, h, passes, smem, (cudaStream_t)stream);
}

// 3x3 stride-2 convolutions of NetC (src/models.py:77-106) on the halo kernel in the fp16 modes.  out(y, x) reads input rows
// 2y-1, 2y, 2y+1: written as 2(y + by) + py these are (by, py) = (-1, 1), (0, 0), (0, 1), so the layer is a 2x2-tap stride-1
// convolution over the four parity phases of the input (space-to-depth), and a parity phase is exactly what a TMA box with
// element stride 2 fetches: no rearranged copy of the input exists anywhere.  The (by, py) = (-1, 0) combinations carry zero
// weights.  w16: the 16-bit pack of pivlfn_conv_tc for passes 4 / 5 of the restated weights [CoutP][4 taps][4 * Cin]
// (tap = (by + 1) * 2 + (bx + 1), channel = (py * 2 + px) * Cin + c).  H, W: INPUT size (even); Cin % 32 == 0.
extern "C" int pivlfn_conv_s2_tc(const float* x, int x_ld, int N, int H, int W, int Cin, const void* w16, const float* bias,
                                 float* y, int y_ld, int Cout, int lrelu, int passes, void* stream) {
    if (!x || !w16 || !y || N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return PIVLFN_EINVAL;
    if (passes != 4 && passes != 5) return PIVLFN_EINVAL;
    if ((H & 1) || (W & 1) || (Cin % KC) || Cout > 128) return PIVLFN_EUNSUPPORTED;
    if (((uintptr_t)x & 15) || (x_ld & 3) || x_ld < Cin || ((uintptr_t)y & 3) || y_ld < Cout || ((uintptr_t)w16 & 15)) return PIVLFN_EINVAL;
    const int Ho = H / 2, Wo = W / 2;
    if (Wo < HT_W) return PIVLFN_EUNSUPPORTED;
    EncodeTiledFn enc = get_encode();
    if (!enc) return PIVLFN_EDRIVER;
    const int CoutP = (Cout + 15) & ~15, CinR = 4 * Cin;
    const int cout_st = (y_ld == ((Cout + 3) & ~3)) ? y_ld : Cout;
    const int vec_store = (!((uintptr_t)y & 31) && !(y_ld & 7) && !(cout_st & 7)) ? 2
                        : ((!((uintptr_t)y & 15) && !(y_ld & 3) && cout_st >= 4) ? 1 : 0);
    ConvHaloArgs h;
    h.bias = bias; h.res = nullptr; h.res_ld = 0; h.y = y; h.y_ld = y_ld;
    h.N = N; h.H = Ho; h.W = Wo; h.Cin = CinR; h.Cout = Cout; h.CoutP = CoutP; h.KH = 2; h.KW = 2;
    h.lrelu = lrelu; h.vec_store = vec_store; h.cout_st = cout_st; h.dbg = g_conv_tc_dbg; h.x_shift = 0; h.planar = 0;
    h.stage_off = 0; h.s2 = 1; h.cpp = Cin / KC; h.out_p16 = 0; h.range_flag = nullptr;
    if (vec_store >= 1 && !(Wo & 3) && !(cout_st & 15) && halo_env().quad_store) h.vec_store = 4;
    int halo_rows = 0;
    const int smem = halo_configure(h, passes, &halo_rows);
    if (smem <= 0) return PIVLFN_EUNSUPPORTED;
    CUtensorMap tmA, tmB16, tmBlo16, tmB3;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)x_ld * 4, (cuuint64_t)W * x_ld * 4, (cuuint64_t)H * W * x_ld * 4};
        cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)(2 * (HT_W + 1)), (cuuint32_t)(2 * halo_rows), 1};
        cuuint32_t estr[4] = {1, 2, 2, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return PIVLFN_EINVAL;
    }
    const size_t tile = (size_t)CoutP * 4 * CinR * 2;
    if (encode_weights_bf16(enc, &tmB16, w16, CinR, 4, CoutP)) return PIVLFN_EINVAL;
    if (encode_weights_bf16(enc, &tmBlo16, (const char*)w16 + tile, CinR, 4, CoutP)) return PIVLFN_EINVAL;
    tmB3 = tmBlo16;
    if (passes == 5 && encode_weights_bf16(enc, &tmB3, (const char*)w16 + 2 * tile, CinR, 4, CoutP)) return PIVLFN_EINVAL;
    h.w_img = (const uint8_t*)w16;
    return halo_launch(tmA, tmB16, tmB3, tmB16, tmBlo16, h, passes, smem, (cudaStream_t)stream);
}
