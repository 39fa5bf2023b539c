// 49-displacement (+-3) cost volume, stride 1 or 2.
//
// Replaces src/correlation.py:285-344 (_FunctionCorrelation.forward) and its kernels
// kernel_Correlation_rearrange (:9-34) + kernel_Correlation_updateOutput (:36-104).  The reference
// first writes two zero-padded NHWC copies to HBM and then runs one 32-thread block per output pixel;
// here the padding/rearrange is folded into the shared-memory tile load, each thread keeps its
// displacement accumulators in registers, and nothing but the inputs and the 49-channel result
// touches HBM.
//
//   out[b,(dy+3)*7+(dx+3),y,x] = (1/C) * sum_c f1[b,c,y*s,x*s] * f2[b,c,(y+dy)*s,(x+dx)*s]
//
// Two entry points: NCHW (the public FunctionCorrelation operator) and NHWC (model-internal, with the
// backwarp of f2 and the LeakyReLU of src/models.py:171-184 fused in).
#include <stdlib.h>
#include "common.cuh"
#include "p16.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// NCHW (the public FunctionCorrelation operator): CTA = 32 x 8 output pixels, 224 threads = 7 warps.  Warp = displacement
// row dy; lane = (4-pixel segment of a row, row pair): a thread owns 2 rows x 4 adjacent pixels x 7 displacements dx =
// 56 accumulators.  Channels are staged 8 at a time as PLANES in shared memory ([channel][row][column], the NCHW order, so
// the rearrange kernels of the reference fold into coalesced row loads): per channel a thread reads its 4 f1 values and the
// 12 f2 values they share as four float4 per row (consecutive lanes = consecutive float4: conflict-free), 56 FMA for 8
// shared-memory loads.  (One pixel and 49 accumulators per thread needed one load per FMA and ran at 19 % of the HBM
// roofline.)
// ------------------------------------------------------------------------------------------------
constexpr int NC_TX = 32, NC_TY = 8, NC_CK = 8;
constexpr int NC_SW = 40, NC_SH = NC_TY + 6;            // f2 tile: 38 columns used, row pitch 40 floats
constexpr int NC_THREADS = 224;

__global__ void __launch_bounds__(NC_THREADS, 2)
corr_nchw_kernel(const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ out,
                 int C, int H, int W, int Ho, int Wo, int s) {
    __shared__ __align__(16) float s2[NC_CK][NC_SH][NC_SW];
    __shared__ __align__(16) float s1[NC_CK][NC_TY][NC_TX];
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * NC_TX, y0 = blockIdx.y * NC_TY;
    const int tid = threadIdx.x, lane = tid & 31, dy = tid >> 5;
    const int seg = lane & 7, ty = lane >> 3;            // rows ty and ty + 4
    const size_t plane = (size_t)H * W;
    const float* f1b = f1 + (size_t)b * C * plane;
    const float* f2b = f2 + (size_t)b * C * plane;

    float acc[2][4][7];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int d = 0; d < 7; ++d) acc[r][i][d] = 0.f;

    for (int c0 = 0; c0 < C; c0 += NC_CK) {
        __syncthreads();
        // f2 tile: 8 channels x 14 rows = 112 (channel, row) lines of 38 sampled pixels; a warp takes whole lines (lanes =
        // columns: coalesced, no per-element index arithmetic), four lines = up to 8 loads in flight per lane
        const int wp = tid >> 5;
        for (int p0 = wp; p0 < NC_CK * NC_SH; p0 += 4 * 7) {
            float va[4], vb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int p = p0 + 7 * e;
                va[e] = vb[e] = 0.f;
                if (p < NC_CK * NC_SH) {
                    const int cc = p / NC_SH, j = p - cc * NC_SH;
                    const int iy = (y0 + j - 3) * s;
                    if (c0 + cc < C && iy >= 0 && iy < H) {
                        const float* rowp = f2b + (size_t)(c0 + cc) * plane + (size_t)iy * W;
                        const int ixa = (x0 + lane - 3) * s, ixb = (x0 + lane + 29) * s;
                        if (ixa >= 0 && ixa < W) va[e] = __ldg(rowp + ixa);
                        if (lane < 6 && ixb >= 0 && ixb < W) vb[e] = __ldg(rowp + ixb);
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int p = p0 + 7 * e;
                if (p < NC_CK * NC_SH) {
                    const int cc = p / NC_SH, j = p - cc * NC_SH;
                    s2[cc][j][lane] = va[e];
                    if (lane < 6) s2[cc][j][lane + 32] = vb[e];
                }
            }
        }
        // f1 tile: 8 channels x 8 rows = 64 lines of 32 pixels
        for (int p0 = wp; p0 < NC_CK * NC_TY; p0 += 4 * 7) {
            float va[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int p = p0 + 7 * e;
                va[e] = 0.f;
                if (p < NC_CK * NC_TY) {
                    const int cc = p / NC_TY, j = p - cc * NC_TY;
                    const int oy = y0 + j, ox = x0 + lane;
                    if (c0 + cc < C && oy < Ho && ox < Wo)
                        va[e] = __ldg(f1b + (size_t)(c0 + cc) * plane + (size_t)(oy * s) * W + ox * s);
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int p = p0 + 7 * e;
                if (p < NC_CK * NC_TY) s1[p / NC_TY][p % NC_TY][lane] = va[e];
            }
        }
        __syncthreads();
#pragma unroll 2
        for (int cc = 0; cc < NC_CK; ++cc) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = ty + 4 * r;
                const float4 a4 = *reinterpret_cast<const float4*>(&s1[cc][row][seg * 4]);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                float v[12];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float4 t = *reinterpret_cast<const float4*>(&s2[cc][row + dy][seg * 4 + q * 4]);
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int d = 0; d < 7; ++d) acc[r][i][d] = fmaf(a[i], v[i + d], acc[r][i][d]);
            }
        }
    }
    const float inv = 1.f / (float)C;
    const size_t oplane = (size_t)Ho * Wo;
    const bool vec = !(Wo & 3) && !((uintptr_t)out & 15);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int oy = y0 + ty + 4 * r, ox = x0 + seg * 4;
        if (oy >= Ho || ox >= Wo) continue;
#pragma unroll
        for (int d = 0; d < 7; ++d) {
            float* o = out + ((size_t)b * 49 + dy * 7 + d) * oplane + (size_t)oy * Wo + ox;
            if (vec && ox + 3 < Wo) {
                *reinterpret_cast<float4*>(o) = make_float4(acc[r][0][d] * inv, acc[r][1][d] * inv, acc[r][2][d] * inv, acc[r][3][d] * inv);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (ox + i < Wo) o[i] = acc[r][i][d] * inv;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// NHWC: CTA = 16 x 8 output pixels, 224 threads = 7 warps.  Warp = displacement row dy, lane = (4-pixel segment of an
// output row, row): a thread owns 4 horizontally adjacent pixels x 7 displacements dx = 28 accumulators, so that per
// channel quad it reads its 4 f1 pixels and the 4+6 f2 pixels they share ONCE (14 float4 loads for 112 FMA; one pixel per
// thread needed 29 loads for the same work, and shared-memory reads, not HBM, bound this kernel).  Channels are staged 32
// at a time into shared memory (one full 128-byte line per gathered pixel): the f1 tile and the f2 tile (+-3 halo, sampled every s pixels).  The backwarp of f2
// (src/models.py:171) is folded into the tile load: the bilinear taps of every tile pixel are computed ONCE per CTA
// (they do not depend on the channel chunk), so the per-chunk gathers are independent loads with no flow -> address
// dependency and are issued twelve at a time per thread.  Two CTAs per SM overlap one CTA's load phase with the
// other's compute phase.
// Bank conflicts: pixel pitch 36 floats = 9 float4 (odd), so 8 CONSECUTIVE pixels are conflict-free, but lanes here are 4
// pixels apart.  Each lane therefore walks the 8 channel quads of a chunk in a rotated order, q = (qi + (seg >> 1)) & 7;
// with row pitches of 22 (f2) and 18 (f1, padded) pixels the eight float4 addresses of every quarter-warp then fall into
// eight different 16-byte bank groups.  The channel sum is order-independent per thread, so the rotation only permutes
// the fp32 summation order.
// ------------------------------------------------------------------------------------------------
constexpr int NH_TX = 16, NH_TY = 8, NH_CK = 32, NH_PITCH = 36, NH_NQ = NH_CK / 4;
constexpr int NH_SW = NH_TX + 6, NH_SH = NH_TY + 6;
constexpr int NH_S1W = NH_TX + 2;                        // padded f1 row pitch (pixels)
constexpr int NH_NPIX2 = NH_SW * NH_SH;                 // 308 tile pixels of f2
constexpr int NH_THREADS = 224;
constexpr int NH_S1 = NH_S1W * NH_TY * NH_PITCH, NH_S2 = NH_NPIX2 * NH_PITCH;
constexpr int NH_OLD = 52;                               // staged output row: 49 displacements + pad

__device__ __forceinline__ float4 ld_quad(const float* src, int c, int C) {
    if (c + 3 < C) return __ldg(reinterpret_cast<const float4*>(src));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    v.x = __ldg(src);
    if (c + 1 < C) v.y = __ldg(src + 1);
    if (c + 2 < C) v.z = __ldg(src + 2);
    return v;
}

// channels c .. c+3 of a P16 pixel row (p16.cuh): 8 bytes of hi + 8 bytes of lo' -> 4 floats (pad channels of the last group are
// stored as zeros, so no channel-count test is needed)
__device__ __forceinline__ float4 ld_quad_p16(const float* pixel_row, int c) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(pixel_row) + (c >> 4) * 64;
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(p + (c & 15) * 2));
    const uint32_t l = __ldg(reinterpret_cast<const uint32_t*>(p + 32 + (c & 15)));
    float v[4];
    p16::decode4(h.x, h.y, l, v);
    return make_float4(v[0], v[1], v[2], v[3]);
}

// F1P / F2P: the feature maps are P16 (f*_ld = pixel pitch in words either way); OUTP: the result is written as P16 groups
// (64 channels: 49 displacements + zero pad) for a tensor-core consumer instead of fp32 rows
template <bool F1P, bool F2P, bool OUTP>
__global__ void __launch_bounds__(NH_THREADS, 3)
corr_nhwc_kernel(const float* __restrict__ f1, int f1_ld, const float* __restrict__ f2, int f2_ld,
                 const float* __restrict__ flow, float fscale, float* __restrict__ out, int out_ld,
                 int C, int H, int W, int Ho, int Wo, int s, int lrelu, int* __restrict__ range_flag) {
    extern __shared__ __align__(16) float sbuf[];        // f1 tile | f2 tile (reused as the output staging tile [128][NH_OLD]) | taps
    float* const s1 = sbuf;
    float* const s2 = sbuf + NH_S1;
    float4* const tapw = reinterpret_cast<float4*>(sbuf + NH_S1 + NH_S2);      // bilinear weights of the tile pixels
    int2* const tapxy = reinterpret_cast<int2*>(tapw + NH_NPIX2);              // top-left tap (x0, y0)
    static_assert(NH_TX * NH_TY * NH_OLD <= NH_S1 + NH_S2 && NH_S1 % 4 == 0, "output staging must fit in the tiles");
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * NH_TX, y0 = blockIdx.y * NH_TY;
    const int tid = threadIdx.x;
    const size_t img = (size_t)n * H * W;
    const int nch = (C + NH_CK - 1) / NH_CK;

    // ---- bilinear taps of the 308 f2-tile pixels, once per CTA -----------------------------------------------
    for (int p = tid; p < NH_NPIX2; p += NH_THREADS) {
        const int i = p % NH_SW, j = p / NH_SW;
        const int iy = (y0 + j - 3) * s, ix = (x0 + i - 3) * s;
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
        int2 xy = make_int2(0, 0);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
            float fx = 0.f, fy = 0.f;
            if (flow != nullptr) {
                const float2 fl = __ldg(reinterpret_cast<const float2*>(flow) + img + (size_t)iy * W + ix);
                fx = fl.x * fscale; fy = fl.y * fscale;
            }
            const BilinearTaps t = make_taps((float)ix + fx, (float)iy + fy, H, W);
            wv = make_float4(t.w00, t.w01, t.w10, t.w11);
            xy = make_int2(t.x0, t.y0);
        }
        tapw[p] = wv;
        tapxy[p] = xy;
    }

    const int lane = tid & 31, dy = tid >> 5;            // warp = displacement row
    const int seg = lane & 3, ty = lane >> 2;            // 4-pixel segment of row ty
    const int rot = seg >> 1;
    float acc[4][7];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int d = 0; d < 7; ++d) acc[i][d] = 0.f;

    for (int ch = 0; ch < nch; ++ch) {
        const int c0 = ch * NH_CK;
        __syncthreads();                                  // previous chunk consumed (and tap table visible)
        // ---- f1 tile: 128 pixels x 4 quads ------------------------------------------------------------------
        for (int item = tid; item < NH_TX * NH_TY * NH_NQ; item += NH_THREADS) {
            const int q = item % NH_NQ, p = item / NH_NQ;
            const int lx = p % NH_TX, ly = p / NH_TX;
            const int px = x0 + lx, py = y0 + ly;
            const int c = c0 + q * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (px < Wo && py < Ho && c < C) {
                const float* row = f1 + (img + (size_t)(py * s) * W + px * s) * f1_ld;
                v = F1P ? ld_quad_p16(row, c) : ld_quad(row + c, c, C);
            }
            *reinterpret_cast<float4*>(&s1[(ly * NH_S1W + lx) * NH_PITCH + q * 4]) = v;
        }
        // ---- f2 tile (+halo): 1232 items, three at a time (12 independent gathers in flight per thread) --------
        constexpr int UNR = 2;
        for (int base = tid; base < NH_NPIX2 * NH_NQ; base += UNR * NH_THREADS) {
            float4 u[UNR][4];
            float4 wv[UNR];
            bool ok[UNR];
#pragma unroll
            for (int e = 0; e < UNR; ++e) {
                const int item = base + e * NH_THREADS;
                ok[e] = item < NH_NPIX2 * NH_NQ;
                const int q = item % NH_NQ, p = ok[e] ? (item / NH_NQ) : 0;
                const int c = c0 + q * 4;
                wv[e] = tapw[p];
                const int2 xy = tapxy[p];
                const float wgt[4] = {wv[e].x, wv[e].y, wv[e].z, wv[e].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    u[e][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok[e] && c < C && wgt[k] != 0.f) {
                        const float* row = f2 + (img + (size_t)(xy.y + (k >> 1)) * W + (xy.x + (k & 1))) * f2_ld;
                        u[e][k] = F2P ? ld_quad_p16(row, c) : ld_quad(row + c, c, C);
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < UNR; ++e) {
                if (ok[e]) {
                    const int item = base + e * NH_THREADS;
                    const int q = item % NH_NQ, p = item / NH_NQ;
                    const float wgt[4] = {wv[e].x, wv[e].y, wv[e].z, wv[e].w};
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        v.x = fmaf(wgt[k], u[e][k].x, v.x); v.y = fmaf(wgt[k], u[e][k].y, v.y);
                        v.z = fmaf(wgt[k], u[e][k].z, v.z); v.w = fmaf(wgt[k], u[e][k].w, v.w);
                    }
                    *reinterpret_cast<float4*>(&s2[p * NH_PITCH + q * 4]) = v;
                }
            }
        }
        __syncthreads();
        // ---- 4 pixels x 7 dx per thread, channel quads in rotated order ----------------------------------------
        const float* a0 = &s1[(ty * NH_S1W + seg * 4) * NH_PITCH];
        const float* b0 = &s2[((ty + dy) * NH_SW + seg * 4) * NH_PITCH];
#pragma unroll 2
        for (int qi = 0; qi < NH_NQ; ++qi) {
            const int q = ((qi + rot) & (NH_NQ - 1)) * 4;
            float4 a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * NH_PITCH + q);
            // stream the 10 f2 pixels: pixel j serves (i, d = j - i) for every output pixel i it is in range of, so only two
            // of them are live at a time (register pressure decides how many CTAs share an SM during the gather phases)
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                const float4 v = *reinterpret_cast<const float4*>(b0 + j * NH_PITCH + q);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int d = j - i;
                    if (d >= 0 && d < 7) {
                        float t = acc[i][d];
                        t = fmaf(a[i].x, v.x, t);
                        t = fmaf(a[i].y, v.y, t);
                        t = fmaf(a[i].z, v.z, t);
                        t = fmaf(a[i].w, v.w, t);
                        acc[i][d] = t;
                    }
                }
            }
        }
    }
    // ---- stage the 128 x 49 results in shared memory, then store whole pixel rows ---------------------------------
    __syncthreads();                                      // everyone is done reading the tiles
    const float inv = 1.f / (float)C;
    float* so = sbuf;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int d = 0; d < 7; ++d) {
            const float v = acc[i][d] * inv;
            so[(ty * NH_TX + seg * 4 + i) * NH_OLD + dy * 7 + d] = lrelu ? lrelu_f(v) : v;
        }
    __syncthreads();
    if (OUTP) {
        // 8 units of 8 channels per pixel (channels >= 49 are zero), each encoded as a 16-byte hi and a 16-byte lo' vector
        uint32_t bad = 0;
        for (int item = tid; item < NH_TX * NH_TY * 8; item += NH_THREADS) {
            const int p = item >> 3, un = item & 7;
            const int ox = x0 + p % NH_TX, oy = y0 + p / NH_TX;
            if (ox < Wo && oy < Ho) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = (un * 8 + j < 49) ? so[p * NH_OLD + un * 8 + j] : 0.f;
                uint4 h;
                uint2 l, g;
                p16::encode8(v, h, l, g);
                bad |= p16::nonfinite_bits(h);
                uint8_t* o = reinterpret_cast<uint8_t*>(out + ((size_t)n * Ho * Wo + (size_t)oy * Wo + ox) * out_ld);
                *reinterpret_cast<uint4*>(o + p16::unit_off_bytes(un)) = h;
                *reinterpret_cast<uint2*>(o + p16::unit_lo8_bytes(un)) = l;
                *reinterpret_cast<uint2*>(o + p16::unit_hi8_bytes(un)) = g;
            }
        }
        if (range_flag && p16::any_nonfinite(bad)) *range_flag = 1;
        return;
    }
    // rows of exactly 52 floats are a dedicated buffer with its own padding: whole float4 rows (pad channels written as 0)
    const bool vec = out_ld == NH_OLD && !((uintptr_t)out & 15);
    if (vec) {
        for (int item = tid; item < NH_TX * NH_TY * (NH_OLD / 4); item += NH_THREADS) {
            const int p = item / (NH_OLD / 4), q = item % (NH_OLD / 4);
            const int ox = x0 + p % NH_TX, oy = y0 + p / NH_TX;
            if (ox < Wo && oy < Ho) {
                float4 v = *reinterpret_cast<const float4*>(&so[p * NH_OLD + q * 4]);
                if (q == NH_OLD / 4 - 1) { v.y = 0.f; v.z = 0.f; v.w = 0.f; }
                *reinterpret_cast<float4*>(out + ((size_t)n * Ho * Wo + (size_t)oy * Wo + ox) * out_ld + q * 4) = v;
            }
        }
    } else {
        for (int item = tid; item < NH_TX * NH_TY * 49; item += NH_THREADS) {
            const int p = item / 49, k = item % 49;
            const int ox = x0 + p % NH_TX, oy = y0 + p / NH_TX;
            if (ox < Wo && oy < Ho) out[((size_t)n * Ho * Wo + (size_t)oy * Wo + ox) * out_ld + k] = so[p * NH_OLD + k];
        }
    }
}

}  // namespace

extern "C" int pivlfn_corr_nchw(const float* first, const float* second, float* out,
                                int B, int C, int H, int W, int stride, void* stream) {
    if (!first || !second || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0) return PIVLFN_EINVAL;
    if (stride != 1 && stride != 2) return PIVLFN_EINVAL;
    if (B > 65535) return PIVLFN_EINVAL;
    const int Ho = cdiv(H, stride), Wo = cdiv(W, stride);
    dim3 grid(cdiv(Wo, NC_TX), cdiv(Ho, NC_TY), B);
    corr_nchw_kernel<<<grid, NC_THREADS, 0, (cudaStream_t)stream>>>(first, second, out, C, H, W, Ho, Wo, stride);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

namespace {

// ------------------------------------------------------------------------------------------------
// Gradients of the public NCHW operator (src/correlation.py:106-234 kernels, :348-405 host side): training only.
//   gradFirst [b,c,Y,X]  = (1/C) sum_{p,o} gradOut[b, (p+3)*7+(o+3), Y/s, X/s] * second[b,c, Y + s p, X + s o]
//   gradSecond[b,c,Y,X]  = (1/C) sum_{p,o} gradOut[b, (p+3)*7+(o+3), Y/s - p, X/s - o] * first[b,c, Y - s p, X - s o]
// at positions with Y % s == 0 and X % s == 0 (the only ones the forward samples; all others get zero), terms whose
// gradOut / feature index falls outside contribute nothing.  One thread per gradient element, X fastest: gradOut and the
// feature rows are read coalesced; the (p, o) sum runs in the reference's order (p outer, o inner).
// ------------------------------------------------------------------------------------------------
template <bool SECOND>
__global__ void __launch_bounds__(256)
corr_grad_nchw_kernel(const float* __restrict__ other, const float* __restrict__ gout, float* __restrict__ grad,
                      int C, int H, int W, int Ho, int Wo, int s, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int X = (int)(i % W);
        long long t = i / W;
        const int Y = (int)(t % H);
        t /= H;
        const int c = (int)(t % C);
        const long long b = t / C;
        float sum = 0.f;
        if (Y % s == 0 && X % s == 0) {
            const int y0 = Y / s, x0 = X / s;
            const float* ob = other + (b * C + c) * (long long)H * W;
            const float* gb = gout + b * 49 * (long long)Ho * Wo;
            for (int p = -3; p <= 3; ++p)
                for (int o = -3; o <= 3; ++o) {
                    const int op = (p + 3) * 7 + (o + 3);
                    if (!SECOND) {
                        // out[y0, x0] saw second at (Y + s p, X + s o)
                        const int yy = Y + s * p, xx = X + s * o;
                        if (y0 < Ho && x0 < Wo && yy >= 0 && yy < H && xx >= 0 && xx < W)
                            sum += __ldg(gb + ((long long)op * Ho + y0) * Wo + x0) * __ldg(ob + (long long)yy * W + xx);
                    } else {
                        // out[y0 - p, x0 - o] saw second at (Y, X) next to first at (Y - s p, X - s o)
                        const int y = y0 - p, x = x0 - o;
                        if (y >= 0 && y < Ho && x >= 0 && x < Wo)
                            sum += __ldg(gb + ((long long)op * Ho + y) * Wo + x) * __ldg(ob + (long long)(y * s) * W + x * s);
                    }
                }
        }
        grad[i] = sum / (float)C;
    }
}

}  // namespace

/* see include/pivlfn.h */
extern "C" int pivlfn_corr_backward_nchw(const float* first, const float* second, const float* grad_out,
                                         float* grad_first, float* grad_second, int B, int C, int H, int W, int stride, void* stream) {
    if (!first || !second || !grad_out || B <= 0 || C <= 0 || H <= 0 || W <= 0) return PIVLFN_EINVAL;
    if (stride != 1 && stride != 2) return PIVLFN_EINVAL;
    const int Ho = cdiv(H, stride), Wo = cdiv(W, stride);
    const long long total = (long long)B * C * H * W;
    long long g = (total + 255) / 256;
    const long long cap = (long long)pivlfn_num_sms() * 32;
    const int grid = (int)(g < cap ? g : cap);
    cudaStream_t st = (cudaStream_t)stream;
    if (grad_first) {
        corr_grad_nchw_kernel<false><<<grid, 256, 0, st>>>(second, grad_out, grad_first, C, H, W, Ho, Wo, stride, total);
        PIVLFN_LAUNCHED();
    }
    if (grad_second) {
        corr_grad_nchw_kernel<true><<<grid, 256, 0, st>>>(first, grad_out, grad_second, C, H, W, Ho, Wo, stride, total);
        PIVLFN_LAUNCHED();
    }
    return pivlfn_last_error();
}

namespace {
template <bool F1P, bool F2P, bool OUTP>
int launch_corr_nhwc(const float* f1, int f1_ld, const float* f2, int f2_ld, const float* flow, float flow_scale, float* out,
                     int out_ld, int N, int H, int W, int C, int stride, int lrelu, int* range_flag, cudaStream_t st) {
    const int Ho = cdiv(H, stride), Wo = cdiv(W, stride);
    dim3 grid(cdiv(Wo, NH_TX), cdiv(Ho, NH_TY), N);
    constexpr int smem = (NH_S1 + NH_S2) * 4 + NH_NPIX2 * (16 + 8);
    static unsigned long long configured = 0;
    {
        cudaError_t e = pivlfn_optin_smem(corr_nhwc_kernel<F1P, F2P, OUTP>, smem, configured);
        if (e != cudaSuccess) return (int)e;
    }
    corr_nhwc_kernel<F1P, F2P, OUTP><<<grid, NH_THREADS, smem, st>>>(f1, f1_ld, f2, f2_ld, flow, flow_scale, out, out_ld,
                                                                      C, H, W, Ho, Wo, stride, lrelu, range_flag);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}
}  // namespace

extern "C" int pivlfn_corr_nhwc(const float* f1, int f1_ld, const float* f2, int f2_ld,
                                const float* flow, float flow_scale, float* out, int out_ld,
                                int N, int H, int W, int C, int stride, int lrelu, void* stream) {
    if (!f1 || !f2 || !out || N <= 0 || C <= 0 || H <= 0 || W <= 0) return PIVLFN_EINVAL;
    if (stride != 1 && stride != 2) return PIVLFN_EINVAL;
    if (f1_ld < C || f2_ld < C || out_ld < 49 || N > 65535) return PIVLFN_EINVAL;
    // float4 tile loads need 16-byte aligned pixel rows
    if (((uintptr_t)f1 & 15) || ((uintptr_t)f2 & 15) || (f1_ld & 3) || (f2_ld & 3)) return PIVLFN_EINVAL;
    if (flow && ((uintptr_t)flow & 7)) return PIVLFN_EINVAL;
    return launch_corr_nhwc<false, false, false>(f1, f1_ld, f2, f2_ld, flow, flow_scale, out, out_ld, N, H, W, C, stride, lrelu,
                                                 nullptr, (cudaStream_t)stream);
}

/* see include/pivlfn.h */
extern "C" int pivlfn_corr_p16(const void* f1, int f1_ld, int f1_p16, const void* f2, int f2_ld, int f2_p16,
                               const float* flow, float flow_scale, void* out, int out_ld, int out_p16,
                               int N, int H, int W, int C, int stride, int lrelu, int* range_flag, void* stream) {
    if (!f1 || !f2 || !out || N <= 0 || C <= 0 || H <= 0 || W <= 0) return PIVLFN_EINVAL;
    if (stride != 1 && stride != 2) return PIVLFN_EINVAL;
    if (f1_ld < C || f2_ld < C || N > 65535) return PIVLFN_EINVAL;
    if (f1_p16 ? (((uintptr_t)f1 & 63) || (f1_ld & 15) || f1_ld < ((C + 15) & ~15)) : (((uintptr_t)f1 & 15) || (f1_ld & 3))) return PIVLFN_EINVAL;
    if (f2_p16 ? (((uintptr_t)f2 & 63) || (f2_ld & 15) || f2_ld < ((C + 15) & ~15)) : (((uintptr_t)f2 & 15) || (f2_ld & 3))) return PIVLFN_EINVAL;
    if (out_p16 ? (((uintptr_t)out & 63) || (out_ld & 15) || out_ld < 64) : (out_ld < 49)) return PIVLFN_EINVAL;
    if (flow && ((uintptr_t)flow & 7)) return PIVLFN_EINVAL;
    const float* a = reinterpret_cast<const float*>(f1);
    const float* b = reinterpret_cast<const float*>(f2);
    float* o = reinterpret_cast<float*>(out);
    cudaStream_t st = (cudaStream_t)stream;
#define PIVLFN_CORR_CASE(A, B, O) \
    if ((f1_p16 != 0) == A && (f2_p16 != 0) == B && (out_p16 != 0) == O) \
        return launch_corr_nhwc<A, B, O>(a, f1_ld, b, f2_ld, flow, flow_scale, o, out_ld, N, H, W, C, stride, lrelu, range_flag, st);
    PIVLFN_CORR_CASE(true, false, false)
    PIVLFN_CORR_CASE(true, false, true)
    PIVLFN_CORR_CASE(true, true, false)
    PIVLFN_CORR_CASE(true, true, true)
    PIVLFN_CORR_CASE(false, false, true)
    PIVLFN_CORR_CASE(false, true, false)
    PIVLFN_CORR_CASE(false, true, true)
    PIVLFN_CORR_CASE(false, false, false)
#undef PIVLFN_CORR_CASE
    return PIVLFN_EINVAL;
}
