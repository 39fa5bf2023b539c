// Memory-bound glue kernels of the P16 pipeline (p16.cuh): the producers that hand MMA-ready (hi | lo8 | hi8) rows to the
// tensor-core convolutions of conv_p16.cu, and the converters between P16 and fp32 NHWC.
#include "common.cuh"
#include "p16.cuh"

namespace {

inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = (long long)pivlfn_num_sms() * 32;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

__device__ __forceinline__ uint4 ldg_u4(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg_u4(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ uint2 ldg_u2(const void* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ void stg_u2(void* p, const uint2& v) { *reinterpret_cast<uint2*>(p) = v; }

// ---- fp32 NHWC <-> P16 -----------------------------------------------------------------------------------------------
// one thread = one pixel x 8 channels
__global__ void p16_encode_kernel(const float* __restrict__ x, int x_ld, int C, uint8_t* __restrict__ y, int y_ld,
                                  long long npix, int* __restrict__ flag) {
    const int U = ((C + 15) >> 4) * 2;
    const long long total = npix * U;
    uint32_t bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / U;
        const int u = (int)(i - p * U);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (8 * u + j < C) ? __ldg(x + p * x_ld + 8 * u + j) : 0.f;
        uint4 h;
        uint2 l, g;
        p16::encode8(v, h, l, g);
        bad |= p16::nonfinite_bits(h);
        uint8_t* o = y + p * (long long)y_ld * 4;
        stg_u4(o + p16::unit_off_bytes(u), h);
        stg_u2(o + p16::unit_lo8_bytes(u), l);
        stg_u2(o + p16::unit_hi8_bytes(u), g);
    }
    if (flag && p16::any_nonfinite(bad)) *flag = 1;
}

__global__ void p16_decode_kernel(const uint8_t* __restrict__ x, int x_ld, int C, float* __restrict__ y, int y_ld, long long npix) {
    const int U = (C + 7) >> 3;
    const long long total = npix * U;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / U;
        const int u = (int)(i - p * U);
        const uint8_t* s = x + p * (long long)x_ld * 4;
        float v[8];
        p16::decode8(ldg_u4(s + p16::unit_off_bytes(u)), ldg_u2(s + p16::unit_lo8_bytes(u)), v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (8 * u + j < C) y[p * y_ld + 8 * u + j] = v[j];
    }
}

// ---- backwarp (src/models.py:20-35) -> P16 slice of the Subpixel concat buffer ----------------------------------------
// 8x8 pixel patch per block, 4 lanes per pixel walking the 8-channel units (see warp_nhwc_kernel in misc.cu); the input is
// fp32 NHWC (NetC_ext output for the second image: consumed only by the cost volume and this kernel) or P16 (NetC features
// of the coarse levels).  The output unit is encoded and written as one 16-byte and two 8-byte vectors.
template <bool IN_P16>
__global__ void __launch_bounds__(256, 4)
warp_p16_kernel(const uint8_t* __restrict__ in, int in_ld, const float2* __restrict__ flow, float scale,
                uint8_t* __restrict__ out, int out_ld, int N, int H, int W, int C, int* __restrict__ flag) {
    const int U = C >> 3;                                   // C % 16 == 0
    const int tiles_x = (W + 7) >> 3, tiles_y = (H + 7) >> 3;
    const long long ntiles = (long long)N * tiles_y * tiles_x;
    const int pp = threadIdx.x >> 2, lq = threadIdx.x & 3;
    uint32_t bad = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tx = (int)(tile % tiles_x);
        const long long t2 = tile / tiles_x;
        const int ty = (int)(t2 % tiles_y);
        const long long n = t2 / tiles_y;
        const int x = tx * 8 + (pp & 7), y = ty * 8 + (pp >> 3);
        if (x >= W || y >= H) continue;
        const long long img = n * H * W;
        const long long p = img + (long long)y * W + x;
        const float2 fl = __ldg(flow + p);
        const BilinearTaps tp = make_taps((float)x + fl.x * scale, (float)y + fl.y * scale, H, W);
        const float wgt[4] = {tp.w00, tp.w01, tp.w10, tp.w11};
        const uint8_t* src[4];
        bool on[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            on[k] = wgt[k] != 0.f;          // taps outside the frame are never dereferenced
            src[k] = in + (img + (long long)(tp.y0 + (k >> 1)) * W + (tp.x0 + (k & 1))) * (long long)in_ld * 4;
        }
        uint8_t* o = out + p * (long long)out_ld * 4;
        for (int u = lq; u < U; u += 4) {
            uint4 a[4], b[4];                                  // IN_P16: b.x, b.y = the 8 lo8 bytes
            const int off = IN_P16 ? p16::unit_off_bytes(u) : u * 32;
            const int off2 = IN_P16 ? p16::unit_lo8_bytes(u) : off + 16;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                a[k] = on[k] ? ldg_u4(src[k] + off) : make_uint4(0, 0, 0, 0);
                if (IN_P16) {
                    const uint2 l = on[k] ? ldg_u2(src[k] + off2) : make_uint2(0, 0);
                    b[k] = make_uint4(l.x, l.y, 0, 0);
                } else {
                    b[k] = on[k] ? ldg_u4(src[k] + off2) : make_uint4(0, 0, 0, 0);
                }
            }
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float t[8];
                if (IN_P16) p16::decode8(a[k], make_uint2(b[k].x, b[k].y), t);
                else {
                    t[0] = __uint_as_float(a[k].x); t[1] = __uint_as_float(a[k].y); t[2] = __uint_as_float(a[k].z); t[3] = __uint_as_float(a[k].w);
                    t[4] = __uint_as_float(b[k].x); t[5] = __uint_as_float(b[k].y); t[6] = __uint_as_float(b[k].z); t[7] = __uint_as_float(b[k].w);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaf(wgt[k], t[j], v[j]);
            }
            uint4 h;
            uint2 l, g;
            p16::encode8(v, h, l, g);
            bad |= p16::nonfinite_bits(h);
            stg_u4(o + p16::unit_off_bytes(u), h);
            stg_u2(o + p16::unit_lo8_bytes(u), l);
            stg_u2(o + p16::unit_hi8_bytes(u), g);
        }
    }
    if (flag && p16::any_nonfinite(bad)) *flag = 1;
}

// ---- depthwise ConvTranspose 4x4 s2 (upCorr_M, src/models.py:151-152): fp32 NHWC in -> P16 out ---------------------------
// one thread = one INPUT pixel position x 4 channels -> the 2x2 output block (see deconv4x4s2_dw_block_kernel in misc.cu): the
// 3x3 input neighbourhood is read once (9 float4), the 16 taps of the 4 channels come from shared memory, and every output
// is written as 8 bytes of hi + 4 bytes of lo8 + 4 bytes of hi8 -- four consecutive lanes complete the 64-byte group.  (8 channels per
// thread kept 18 float4 live, 98 registers, 23 % occupancy: 2.2 TB/s; this shape runs at 2.5x the occupancy.)
constexpr int DC_MAXC = 64;
__global__ void __launch_bounds__(256)
deconv4x4s2_dw_p16_kernel(const float* __restrict__ in, int in_ld, int in_c4, const float* __restrict__ w,
                          uint8_t* __restrict__ out, int out_ld, int N, int H, int W, int C, int* __restrict__ flag) {
    __shared__ __align__(16) float w_s[16 * DC_MAXC];
    for (int i = threadIdx.x; i < 16 * DC_MAXC; i += blockDim.x) {
        const int tap = i / DC_MAXC, c = i - tap * DC_MAXC;
        w_s[i] = c < C ? __ldg(w + c * 16 + tap) : 0.f;
    }
    __syncthreads();
    const int Q = ((C + 15) >> 4) * 4;                      // channel quads of the P16 output (whole 16-channel groups)
    const int Wo = 2 * W;
    const long long total = (long long)N * H * W * Q;
    uint32_t bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i % Q);
        const long long p = i / Q;
        const int ix = (int)(p % W);
        const long long t = p / W;
        const int iy = (int)(t % H);
        const long long n = t / H;
        const int c = q * 4;
        const float* base = in + (n * H * W) * in_ld + c;
        const bool has = c + 4 <= in_c4;                    // this quad exists in the input rows (else: zero pad of the last group)
        float4 v[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const int yy = iy + a - 1, xx = ix + b - 1;
                v[a][b] = (has && yy >= 0 && yy < H && xx >= 0 && xx < W)
                              ? __ldg(reinterpret_cast<const float4*>(base + ((long long)yy * W + xx) * in_ld))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const int ky = dy ? (a ? 0 : 2) : (a ? 1 : 3), kx = dx ? (b ? 0 : 2) : (b ? 1 : 3);
                        const float4 x4 = v[dy + a][dx + b];
                        const float4 ww = *reinterpret_cast<const float4*>(&w_s[(ky * 4 + kx) * DC_MAXC + c]);
                        acc.x = fmaf(x4.x, ww.x, acc.x); acc.y = fmaf(x4.y, ww.y, acc.y);
                        acc.z = fmaf(x4.z, ww.z, acc.z); acc.w = fmaf(x4.w, ww.w, acc.w);
                    }
                // pad channels: zero weights -> exact zeros whatever the input pads hold (unless they are non-finite)
                if (c + 1 >= C) acc.y = 0.f;
                if (c + 2 >= C) acc.z = 0.f;
                if (c + 3 >= C) acc.w = 0.f;
                if (c >= C) acc.x = 0.f;
                const uint32_t h0 = p16::pack_hi(acc.x, acc.y), h1 = p16::pack_hi(acc.z, acc.w);
                bad |= p16::nonfinite_bits(h0) | p16::nonfinite_bits(h1);
                const long long op = (n * 2 * H + 2 * iy + dy) * Wo + 2 * ix + dx;
                uint8_t* o = out + op * (long long)out_ld * 4 + (c >> 4) * 64;
                *reinterpret_cast<uint2*>(o + (c & 15) * 2) = make_uint2(h0, h1);
                *reinterpret_cast<uint32_t*>(o + 32 + (c & 15)) = p16::pack_lo4(acc.x, acc.y, acc.z, acc.w, h0, h1);
                *reinterpret_cast<uint32_t*>(o + 48 + (c & 15)) = p16::pack_e5m2x4(acc.x, acc.y, acc.z, acc.w);
            }
    }
    if (flag && p16::any_nonfinite(bad)) *flag = 1;
}

// ---- regularisation inputs (src/models.py:275-277) -> the (err, rm_u, rm_v) group of the conv_R concat buffer ---------
constexpr int MEAN_PARTS = 32;          // == pivlfn_flow_mean_parts()
__global__ void reg_input_p16_kernel(const float4* __restrict__ img1, const float4* __restrict__ img2,
                                     const float2* __restrict__ flow, float scale, const float* __restrict__ partial,
                                     uint8_t* __restrict__ out, int out_ld, int N, int H, int W, int* __restrict__ flag) {
    const long long HW = (long long)H * W, total = (long long)N * HW;
    const int lane = threadIdx.x & 31;
    const float inv = 1.f / (float)HW;
    uint32_t bad = 0;
    // warp-uniform loop (whole warps enter; lanes past the end idle): the flow mean of the warp's image is the sum of its
    // MEAN_PARTS == 32 partials, one per lane, reduced by shuffles (64 loads per pixel otherwise: the kernel was issue-bound)
    for (long long pb = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); pb < total; pb += (long long)gridDim.x * blockDim.x) {
        const long long p = pb + lane;
        const bool act = p < total;
        const long long n = (act ? p : total - 1) / HW;
        const long long n0 = __shfl_sync(0xffffffffu, n, 0);
        const float2 part = __ldg(reinterpret_cast<const float2*>(partial) + n0 * MEAN_PARTS + lane);
        float mu = part.x, mv = part.y;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            mu += __shfl_xor_sync(0xffffffffu, mu, o);
            mv += __shfl_xor_sync(0xffffffffu, mv, o);
        }
        if (!act) continue;
        if (n != n0) {                                  // the warp straddles two images
            mu = 0.f; mv = 0.f;
            for (int i = 0; i < MEAN_PARTS; ++i) {
                mu += __ldg(partial + (n * MEAN_PARTS + i) * 2);
                mv += __ldg(partial + (n * MEAN_PARTS + i) * 2 + 1);
            }
        }
        mu *= inv; mv *= inv;
        const int x = (int)(p % W), y = (int)((p / W) % H);
        const float2 fl = __ldg(flow + p);
        const BilinearTaps tp = make_taps((float)x + fl.x * scale, (float)y + fl.y * scale, H, W);
        const float wgt[4] = {tp.w00, tp.w01, tp.w10, tp.w11};
        float wr = 0.f, wg = 0.f, wb = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (wgt[k] != 0.f) {
                float4 u = __ldg(img2 + (n * H + (tp.y0 + (k >> 1))) * W + (tp.x0 + (k & 1)));
                wr = fmaf(wgt[k], u.x, wr); wg = fmaf(wgt[k], u.y, wg); wb = fmaf(wgt[k], u.z, wb);
            }
        }
        const float4 a = __ldg(img1 + p);
        const float dr = a.x - wr, dg = a.y - wg, db = a.z - wb;
        const float e = sqrtf(dr * dr + dg * dg + db * db), ru = fl.x - mu, rv = fl.y - mv;
        const uint32_t h0 = p16::pack_hi(e, ru), h1 = p16::pack_hi(rv, 0.f);
        bad |= p16::nonfinite_bits(h0) | p16::nonfinite_bits(h1);
        uint8_t* o = out + p * (long long)out_ld * 4;
        // the whole 64-byte group (channels 3..15 are zeros): two full 32-byte sectors, no partial-sector writes
        stg_u4(o, make_uint4(h0, h1, 0u, 0u));
        stg_u4(o + 16, make_uint4(0u, 0u, 0u, 0u));
        stg_u4(o + 32, make_uint4(p16::pack_lo4(e, ru, rv, 0.f, h0, h1), 0u, 0u, 0u));
        stg_u4(o + 48, make_uint4(p16::pack_e5m2x4(e, ru, rv, 0.f), 0u, 0u, 0u));
    }
    if (flag && p16::any_nonfinite(bad)) *flag = 1;
}

// ---- second half of the tensor-core flow head -------------------------------------------------------------------------------
// The KxK 32 -> 2 flow head (src/models.py:161,205) runs as a 1xK convolution to 2K channels (row ky*2 + co holds
// sum_{kx,c} x[p + kx - P, c] w[co, c, ky, kx]; conv_p16.cu, OUT_PLANES: plane ky = [pixel][2]); this kernel adds the K row
// planes at their vertical offsets (zero outside the frame), the bias and the residual flow, and writes the dense fp32 flow
// plus (optionally) its P16 group in the Subpixel concat buffer (the torch.cat of src/models.py:216).
// HORIZ: the transposed split -- a Kx1 convolution to 2K column channels (kx*2 + co), summed over kx at horizontal offsets (the
// Kx1 halo tile is 8 pixels wide instead of 8 + K - 1: conv_p16 runs it 1.6x faster than the 1xK form).
template <int K, bool HORIZ>
__global__ void __launch_bounds__(256)
head_rows_sum_kernel(const float2* __restrict__ planes, long long plane_pix, const float* __restrict__ bias,
                     const float2* __restrict__ res, float2* __restrict__ out, uint8_t* __restrict__ out_p16, int p16_ld,
                     int N, int H, int W, int* __restrict__ flag) {
    constexpr int P = K / 2;
    const long long HW = (long long)H * W, total = (long long)N * HW;
    const float b0 = bias ? __ldg(bias) : 0.f, b1 = bias ? __ldg(bias + 1) : 0.f;
    uint32_t bad = 0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const int y = (int)((p / W) % H), x = (int)(p % W);
        float su = 0.f, sv = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int q = (HORIZ ? x : y) + k - P;
            if (q >= 0 && q < (HORIZ ? W : H)) {
                const float2 d = __ldg(planes + (long long)k * plane_pix + p + (long long)(k - P) * (HORIZ ? 1 : W));
                su += d.x; sv += d.y;
            }
        }
        su += b0; sv += b1;
        if (res) { const float2 r = __ldg(res + p); su += r.x; sv += r.y; }
        out[p] = make_float2(su, sv);
        if (out_p16) {
            const uint32_t h0 = p16::pack_hi(su, sv);
            bad |= p16::nonfinite_bits(h0);
            uint8_t* o = out_p16 + p * (long long)p16_ld * 4;
            stg_u4(o, make_uint4(h0, 0u, 0u, 0u));        // the whole 64-byte group: full sectors
            stg_u4(o + 16, make_uint4(0u, 0u, 0u, 0u));
            stg_u4(o + 32, make_uint4(p16::pack_lo4(su, sv, 0.f, 0.f, h0, 0u), 0u, 0u, 0u));
            stg_u4(o + 48, make_uint4(p16::pack_e5m2x4(su, sv, 0.f, 0.f), 0u, 0u, 0u));
        }
    }
    if (flag && p16::any_nonfinite(bad)) *flag = 1;
}

}  // namespace

extern "C" int pivlfn_p16_encode(const float* x, int x_ld, int C, void* y, int y_ld, long long npix, int* range_flag, void* stream) {
    if (!x || !y || C <= 0 || npix <= 0 || x_ld < C) return PIVLFN_EINVAL;
    if (((uintptr_t)y & 63) || (y_ld & 15) || y_ld < ((C + 15) & ~15)) return PIVLFN_EINVAL;
    const long long total = npix * (((C + 15) >> 4) * 2);
    p16_encode_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, x_ld, C, reinterpret_cast<uint8_t*>(y), y_ld, npix, range_flag);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

extern "C" int pivlfn_p16_decode(const void* x, int x_ld, int C, float* y, int y_ld, long long npix, void* stream) {
    if (!x || !y || C <= 0 || npix <= 0 || y_ld < C) return PIVLFN_EINVAL;
    if (((uintptr_t)x & 63) || (x_ld & 15) || x_ld < ((C + 15) & ~15)) return PIVLFN_EINVAL;
    const long long total = npix * ((C + 7) >> 3);
    p16_decode_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(x), x_ld, C, y, y_ld, npix);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

extern "C" int pivlfn_warp_p16(const void* in, int in_ld, int in_p16, const float* flow, float scale, void* out, int out_ld,
                               int N, int H, int W, int C, int* range_flag, void* stream) {
    if (!in || !flow || !out || N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 15) || in_ld < C || out_ld < C) return PIVLFN_EINVAL;
    if (((uintptr_t)out & 63) || (out_ld & 15) || ((uintptr_t)flow & 7)) return PIVLFN_EINVAL;
    if (in_p16 ? (((uintptr_t)in & 63) || (in_ld & 15)) : (((uintptr_t)in & 15) || (in_ld & 3))) return PIVLFN_EINVAL;
    const long long ntiles = (long long)N * ((H + 7) / 8) * ((W + 7) / 8);
    const long long cap = (long long)pivlfn_num_sms() * 64;
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(in);
    uint8_t* dst = reinterpret_cast<uint8_t*>(out);
    const float2* fl = reinterpret_cast<const float2*>(flow);
    if (in_p16) warp_p16_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(src, in_ld, fl, scale, dst, out_ld, N, H, W, C, range_flag);
    else warp_p16_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(src, in_ld, fl, scale, dst, out_ld, N, H, W, C, range_flag);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

extern "C" int pivlfn_deconv4x4s2_dw_p16(const float* in, int in_ld, const float* w, void* out, int out_ld,
                                         int N, int H, int W, int C, int* range_flag, void* stream) {
    if (!in || !w || !out || N <= 0 || H <= 0 || W <= 0 || C <= 0 || C > DC_MAXC) return PIVLFN_EINVAL;
    const int c4 = (C + 3) & ~3;
    if (((uintptr_t)in & 15) || (in_ld & 3) || in_ld < c4) return PIVLFN_EINVAL;
    if (((uintptr_t)out & 63) || (out_ld & 15) || out_ld < ((C + 15) & ~15)) return PIVLFN_EINVAL;
    const long long total = (long long)N * H * W * (((C + 15) >> 4) * 4);
    deconv4x4s2_dw_p16_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        in, in_ld, in_ld & ~3, w, reinterpret_cast<uint8_t*>(out), out_ld, N, H, W, C, range_flag);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

extern "C" int pivlfn_reg_input_p16(const float* img1, const float* img2, const float* flow, float scale,
                                    const float* partial, void* out, int out_ld, int N, int H, int W, int* range_flag, void* stream) {
    if (!img1 || !img2 || !flow || !partial || !out || N <= 0 || H <= 0 || W <= 0 || out_ld < 16) return PIVLFN_EINVAL;
    if (((uintptr_t)img1 & 15) || ((uintptr_t)img2 & 15) || ((uintptr_t)flow & 7) || ((uintptr_t)out & 63) || (out_ld & 15)) return PIVLFN_EINVAL;
    if (pivlfn_flow_mean_parts() != MEAN_PARTS) return PIVLFN_EINVAL;
    const long long total = (long long)N * H * W;
    reg_input_p16_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(img1), reinterpret_cast<const float4*>(img2), reinterpret_cast<const float2*>(flow), scale,
        partial, reinterpret_cast<uint8_t*>(out), out_ld, N, H, W, range_flag);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

namespace {
int head_sum_impl(bool horiz, const float* planes, int K, const float* bias, const float* res, float* out,
                  void* out_p16, int p16_ld, int N, int H, int W, int* range_flag, void* stream) {
    if (!planes || !out || N <= 0 || H <= 0 || W <= 0) return PIVLFN_EINVAL;
    if (((uintptr_t)planes & 7) || ((uintptr_t)out & 7) || (res && ((uintptr_t)res & 7))) return PIVLFN_EINVAL;
    if (out_p16 && (((uintptr_t)out_p16 & 63) || (p16_ld & 15) || p16_ld < 16)) return PIVLFN_EINVAL;
    const long long total = (long long)N * H * W;
    const float2* pl = reinterpret_cast<const float2*>(planes);
    const float2* rs = reinterpret_cast<const float2*>(res);
    float2* o = reinterpret_cast<float2*>(out);
    uint8_t* o16 = reinterpret_cast<uint8_t*>(out_p16);
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(total, 256);
#define PIVLFN_HEAD_SUM(KK)                                                                                                   \
    if (horiz) head_rows_sum_kernel<KK, true><<<g, 256, 0, st>>>(pl, total, bias, rs, o, o16, p16_ld, N, H, W, range_flag);  \
    else head_rows_sum_kernel<KK, false><<<g, 256, 0, st>>>(pl, total, bias, rs, o, o16, p16_ld, N, H, W, range_flag)
    switch (K) {
        case 3: PIVLFN_HEAD_SUM(3); break;
        case 5: PIVLFN_HEAD_SUM(5); break;
        case 7: PIVLFN_HEAD_SUM(7); break;
        default: return PIVLFN_EINVAL;
    }
#undef PIVLFN_HEAD_SUM
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}
}  // namespace

/* see include/pivlfn.h */
extern "C" int pivlfn_head_rows_sum(const float* planes, int K, const float* bias, const float* res, float* out,
                                    void* out_p16, int p16_ld, int N, int H, int W, int* range_flag, void* stream) {
    return head_sum_impl(false, planes, K, bias, res, out, out_p16, p16_ld, N, H, W, range_flag, stream);
}

/* see include/pivlfn.h */
extern "C" int pivlfn_head_cols_sum(const float* planes, int K, const float* bias, const float* res, float* out,
                                    void* out_p16, int p16_ld, int N, int H, int W, int* range_flag, void* stream) {
    return head_sum_impl(true, planes, K, bias, res, out, out_p16, p16_ld, N, H, W, range_flag, stream);
}
