// Memory-bound glue kernels of the LiteFlowNet forward pass (all NHWC fp32):
// input prep, image pyramid, depthwise 2x up-convolution, standalone backwarp, flow mean,
// regularisation inputs (mean removal + brightness error with the backwarp fused in) and the
// regularisation tail (negative-square softmax + unfold + weighted sum in ONE kernel).
#include "common.cuh"

long long g_pivlfn_launches = 0;

extern "C" int pivlfn_abi_version(void) { return 1; }
extern "C" long long pivlfn_launch_count(void) { return g_pivlfn_launches; }
extern "C" int pivlfn_device_is_sm100(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10;
}

namespace {

constexpr int MEAN_PARTS = 32;

// ---- src/models.py:321-323 + NCHW -> NHWC4 ------------------------------------------------------
__global__ void prep_images_kernel(float* __restrict__ img1, float* __restrict__ img2, float4* __restrict__ out,
                                   float4* __restrict__ out_pad, int W,
                                   int B, int HW, float m10, float m11, float m12, float m20, float m21, float m22) {
    const long long total = 2LL * B * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int which = i >= (long long)B * HW;
        const long long r = which ? i - (long long)B * HW : i;
        const long long b = r / HW, p = r % HW;
        float* src = (which ? img2 : img1) + b * 3 * HW + p;
        float v0 = src[0] - (which ? m20 : m10);
        float v1 = src[HW] - (which ? m21 : m11);
        float v2 = src[2LL * HW] - (which ? m22 : m12);
        src[0] = v0; src[HW] = v1; src[2LL * HW] = v2;   // the reference mutates its inputs in place
        out[i] = make_float4(v0, v1, v2, 0.f);
        if (out_pad) {
            const long long row = i / W;          // (image, y) row index; rows of the padded copy are W+8 pixels
            out_pad[row * (W + 8) + 4 + (i - row * W)] = make_float4(v0, v1, v2, 0.f);
        }
    }
}

// ---- src/models.py:336-343 -------------------------------------------------------------------------
__global__ void avgpool2_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W, int C) {
    const int Ho = H / 2, Wo = W / 2;
    const long long total = (long long)N * Ho * Wo * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long p = i / C;
        int ox = (int)(p % Wo);
        long long t = p / Wo;
        int oy = (int)(t % Ho);
        long long n = t / Ho;
        const float* s = in + ((n * H + 2 * oy) * W + 2 * ox) * C + c;
        // same association as upsample_bilinear2d: 0.5*(0.5*a + 0.5*b) + 0.5*(0.5*c + 0.5*d)
        float top = 0.5f * s[0] + 0.5f * s[C];
        float bot = 0.5f * s[(long long)W * C] + 0.5f * s[(long long)W * C + C];
        out[i] = 0.5f * top + 0.5f * bot;
    }
}

// ---- ConvTranspose2d(C,C,4,s2,p1,groups=C,bias=False): src/models.py:144-145,151-152 ---------------
// out[oy,ox,c] = sum over the (at most) 2x2 input pixels with ky = oy+1-2*iy, kx = ox+1-2*ix in [0,3].
// One thread = one output pixel x V consecutive channels (V = 4: float4 traffic; V = 1: any layout).
// Accumulation order: increasing (iy, ix), like a gather formulation of conv_transpose.
template <int V>
__global__ void deconv4x4s2_dw_kernel(const float* __restrict__ in, int in_ld, const float* __restrict__ w,
                                      float* __restrict__ out, int out_ld, int N, int H, int W, int C) {
    const int Ho = 2 * H, Wo = 2 * W;
    const int G = (C + V - 1) / V;
    const long long total = (long long)N * Ho * Wo * G;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % G);
        const long long p = i / G;
        const int ox = (int)(p % Wo);
        const long long t = p / Wo;
        const int oy = (int)(t % Ho);
        const long long n = t / Ho;
        // oy even (2k): (iy,ky) in {(k-1,3),(k,1)};  oy odd (2k+1): {(k,2),(k+1,0)}   (listed in increasing iy)
        const int iy0 = (oy & 1) ? (oy >> 1) : (oy >> 1) - 1, ky0 = (oy & 1) ? 2 : 3;
        const int iy1 = iy0 + 1, ky1 = (oy & 1) ? 0 : 1;
        const int ix0 = (ox & 1) ? (ox >> 1) : (ox >> 1) - 1, kx0 = (ox & 1) ? 2 : 3;
        const int ix1 = ix0 + 1, kx1 = (ox & 1) ? 0 : 1;
        const bool vy0 = iy0 >= 0, vy1 = iy1 < H, vx0 = ix0 >= 0, vx1 = ix1 < W;
        const int c = g * V;
        const float* base = in + n * H * W * in_ld + c;
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0.f;
        const int iys[2] = {iy0, iy1}, kys[2] = {ky0, ky1}, ixs[2] = {ix0, ix1}, kxs[2] = {kx0, kx1};
        const bool vys[2] = {vy0, vy1}, vxs[2] = {vx0, vx1};
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                if (vys[a] && vxs[b]) {
                    const float* src = base + ((long long)iys[a] * W + ixs[b]) * in_ld;
                    float v[V];
                    if (V == 4) {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(src));
                        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                    } else {
                        v[0] = __ldg(src);
                    }
#pragma unroll
                    for (int j = 0; j < V; ++j)
                        if (c + j < C) acc[j] = fmaf(v[j], __ldg(w + (c + j) * 16 + kys[a] * 4 + kxs[b]), acc[j]);
                }
            }
        float* o = out + p * out_ld + c;
        if (V == 4) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else o[0] = acc[0];
    }
}

// Vector path (upCorr_M, 49 channels padded to 52): one thread = one INPUT pixel position x 4 channels -> the 2x2 output
// block (2iy..2iy+1, 2ix..2ix+1), which reads the 3x3 input neighbourhood once (9 float4 loads for 4 float4 stores
// instead of 16) with the 16 taps of its 4 channels taken from shared memory ([tap][channel], conflict-free float4
// reads) instead of 16 scalar global loads per output.  Same accumulation order as the kernel above.
constexpr int DECONV_MAXC = 64;
__global__ void __launch_bounds__(256)
deconv4x4s2_dw_block_kernel(const float* __restrict__ in, int in_ld, const float* __restrict__ w,
                            float* __restrict__ out, int out_ld, int N, int H, int W, int C) {
    __shared__ __align__(16) float w_s[16 * DECONV_MAXC];
    const int CP = (C + 3) & ~3;
    for (int i = threadIdx.x; i < 16 * CP; i += blockDim.x) {
        const int tap = i / CP, c = i - tap * CP;
        w_s[tap * DECONV_MAXC + c] = c < C ? __ldg(w + c * 16 + tap) : 0.f;
    }
    __syncthreads();
    const int G = CP / 4;
    const int Wo = 2 * W;
    const long long total = (long long)N * H * W * G;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % G);
        const long long p = i / G;
        const int ix = (int)(p % W);
        const long long t = p / W;
        const int iy = (int)(t % H);
        const long long n = t / H;
        const int c = g * 4;
        const float* base = in + (n * H * W) * in_ld + c;
        float4 v[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const int yy = iy + a - 1, xx = ix + b - 1;
                v[a][b] = (yy >= 0 && yy < H && xx >= 0 && xx < W)
                              ? __ldg(reinterpret_cast<const float4*>(base + ((long long)yy * W + xx) * in_ld))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        // output row 2iy   : input rows (iy-1, ky 3), (iy, ky 1);   row 2iy+1: (iy, ky 2), (iy+1, ky 0); same along x
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
                        const float4 q = v[dy + a][dx + b];
                        const float4 ww = *reinterpret_cast<const float4*>(&w_s[(ky * 4 + kx) * DECONV_MAXC + c]);
                        acc.x = fmaf(q.x, ww.x, acc.x); acc.y = fmaf(q.y, ww.y, acc.y);
                        acc.z = fmaf(q.z, ww.z, acc.z); acc.w = fmaf(q.w, ww.w, acc.w);
                    }
                if (c + 1 >= C) acc.y = 0.f;        // pad channels are written as exact zeros whatever the input pads hold
                if (c + 2 >= C) acc.z = 0.f;
                if (c + 3 >= C) acc.w = 0.f;
                const long long op = (n * 2 * H + 2 * iy + dy) * Wo + 2 * ix + dx;
                *reinterpret_cast<float4*>(out + op * out_ld + c) = acc;
            }
    }
}

// ---- channel-slice copy (the torch.cat of src/models.py:216,280) ------------------------------------------
template <bool VEC>
__global__ void copy_nhwc_kernel(const float* __restrict__ in, int in_ld, float* __restrict__ out, int out_ld,
                                 long long npix, int C) {
    const int Q = VEC ? C / 4 : C;
    const long long total = npix * Q;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / Q;
        const int q = (int)(i % Q);
        if (VEC)
            *reinterpret_cast<float4*>(out + p * out_ld + q * 4) =
                __ldg(reinterpret_cast<const float4*>(in + p * in_ld + q * 4));
        else
            out[p * out_ld + q] = __ldg(in + p * in_ld + q);
    }
}

// ---- backwarp, standalone (src/models.py:20-35) ---------------------------------------------------------
// One block = one 8x8 pixel patch (a 2-D patch keeps the bilinear taps of neighbouring pixels in L1), 256 threads =
// 64 pixels x 4 lanes.  The four lanes of a pixel compute its taps once and then walk the channel quads q = lane, lane+4,
// ...: every load instruction fetches 64 contiguous bytes per pixel and tap, and the per-quad instruction count is ~45
// (the previous one-quad-per-thread version spent 229 instructions per quad on index arithmetic and taps and was
// issue-bound at 64 % issue utilisation with DRAM at 40 %).
__global__ void __launch_bounds__(256, 4)
warp_nhwc_kernel(const float* __restrict__ in, int in_ld, const float2* __restrict__ flow, float scale,
                 float* __restrict__ out, int out_ld, int N, int H, int W, int C) {
    const int Q = (C + 3) >> 2;
    const int tiles_x = (W + 7) >> 3, tiles_y = (H + 7) >> 3;
    const long long ntiles = (long long)N * tiles_y * tiles_x;
    const int pp = threadIdx.x >> 2, lq = threadIdx.x & 3;
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
        // taps with zero weight (outside the frame) are never dereferenced (their value must not matter, even if non-finite)
        const float* src[4];
        bool on[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            on[k] = wgt[k] != 0.f;
            src[k] = in + (img + (long long)(tp.y0 + (k >> 1)) * W + (tp.x0 + (k & 1))) * in_ld;
        }
        float* o = out + p * out_ld;
        if ((C & 3) == 0) {
            for (int q0 = lq; q0 < Q; q0 += 8) {
                // two quads in flight: 8 independent 16-byte loads per thread (more would cost the third resident block)
                float4 u[2][4];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int q = q0 + 4 * j;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        u[j][k] = (q < Q && on[k]) ? __ldg(reinterpret_cast<const float4*>(src[k]) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int q = q0 + 4 * j;
                    if (q < Q) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            v.x = fmaf(wgt[k], u[j][k].x, v.x); v.y = fmaf(wgt[k], u[j][k].y, v.y);
                            v.z = fmaf(wgt[k], u[j][k].z, v.z); v.w = fmaf(wgt[k], u[j][k].w, v.w);
                        }
                        *(reinterpret_cast<float4*>(o) + q) = v;
                    }
                }
            }
        } else {
            for (int c = lq; c < C; c += 4) {
                float v = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) if (on[k]) v = fmaf(wgt[k], __ldg(src[k] + c), v);
                o[c] = v;
            }
        }
    }
}

// ---- flow mean partials (src/models.py:275) --------------------------------------------------------------
__global__ void __launch_bounds__(256)
flow_mean_kernel(const float2* __restrict__ flow, float* __restrict__ partial, int HW) {
    const int n = blockIdx.y, part = blockIdx.x;
    const long long per = ((long long)HW + MEAN_PARTS - 1) / MEAN_PARTS;
    const long long beg = part * per, end = min((long long)HW, beg + per);
    float su = 0.f, sv = 0.f;
    for (long long i = beg + threadIdx.x; i < end; i += 256) {
        float2 f = __ldg(flow + (long long)n * HW + i);
        su += f.x; sv += f.y;
    }
    __shared__ float ru[8], rv[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        su += __shfl_xor_sync(0xffffffffu, su, o);
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
    }
    if ((threadIdx.x & 31) == 0) { ru[threadIdx.x >> 5] = su; rv[threadIdx.x >> 5] = sv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; ++i) { a += ru[i]; b += rv[i]; }
        partial[((long long)n * MEAN_PARTS + part) * 2 + 0] = a;
        partial[((long long)n * MEAN_PARTS + part) * 2 + 1] = b;
    }
}

// ---- regularisation inputs (src/models.py:275-277) ---------------------------------------------------------
__global__ void reg_input_kernel(const float4* __restrict__ img1, const float4* __restrict__ img2,
                                 const float2* __restrict__ flow, float scale, const float* __restrict__ partial,
                                 float* __restrict__ out, int out_ld, int N, int H, int W) {
    const long long HW = (long long)H * W, total = (long long)N * HW;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        const long long n = p / HW;
        const int x = (int)(p % W), y = (int)((p / W) % H);
        float mu = 0.f, mv = 0.f;
#pragma unroll 8
        for (int i = 0; i < MEAN_PARTS; ++i) {
            mu += __ldg(partial + (n * MEAN_PARTS + i) * 2);
            mv += __ldg(partial + (n * MEAN_PARTS + i) * 2 + 1);
        }
        const float inv = 1.f / (float)HW;
        mu *= inv; mv *= inv;
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
        float* o = out + p * out_ld;
        o[0] = sqrtf(dr * dr + dg * dg + db * db);
        o[1] = fl.x - mu;
        o[2] = fl.y - mv;
    }
}

// ---- regularisation tail (src/models.py:281-302), one kernel ---------------------------------------------------
// d_k = exp(-x_k^2 - max_j(-x_j^2));  u_out = (sum_k wx_k d_k u_N(k) + bx) / sum_k d_k, same for v.
//
// Bulk variant (aligned rows): persistent blocks of 128 threads walk tiles of 128 consecutive pixels.  A tile's K*K
// distance rows are ONE contiguous block of 128 * dist_ld floats, fetched with a single cp.async.bulk (1-D TMA copy,
// no registers) into one of two shared-memory buffers while the previous tile is being computed.  Every thread then
// reads its own row as float4 (pitch dist_ld = 52 / 28 / 12 floats: the 8 lanes of a quarter-warp hit disjoint banks).
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// exp(-x^2 - max_j(-x_j^2)) = 2^((min_j x_j^2 - x^2) * log2 e): FADD + FMUL + one MUFU.EX2 (ex2.approx: relative error 2^-22,
// far inside the flow tolerance; the argument is <= 0, the result in (0, 1])
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int K>
__device__ __forceinline__ void reg_tail_pixel(const float* __restrict__ d, const float2* __restrict__ flow, long long n,
                                               int x, int y, int H, int W, const float* swx, const float* swy, float bxv,
                                               float byv, float& u, float& v) {
    constexpr int KK = K * K, P = K / 2;
    constexpr float LOG2E = 1.4426950408889634f;
    float s[KK];
    float mn = INFINITY;
#pragma unroll
    for (int k = 0; k < KK; ++k) { s[k] = d[k] * d[k]; mn = fminf(mn, s[k]); }
    float sum = 0.f, au = 0.f, av = 0.f;
    // ONE code path for interior and border pixels: K row pointers with constant column offsets, and a K+K-bit validity mask
    // that predicates the neighbour loads (a divergent border branch made whole warps run both paths and the other warps of
    // the block wait for them at the tile barrier: 29 % of the lanes were idle)
    unsigned cm = 0, rm = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        cm |= (unsigned)(x + j - P >= 0 && x + j - P < W) << j;
        rm |= (unsigned)(y + j - P >= 0 && y + j - P < H) << j;
    }
    const float2* c0 = flow + (n * H + (y - P)) * W + (x - P);
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
        const float2* row = c0 + (long long)ky * W;
        const unsigned m = ((rm >> ky) & 1u) ? cm : 0u;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            const int k = ky * K + kx;
            const float e = ex2_approx((mn - s[k]) * LOG2E);       // (mn - s) first: no overflow of mn * log2 e
            sum += e;
            float2 f = make_float2(0.f, 0.f);
            if ((m >> kx) & 1u) f = __ldg(row + kx);
            au = fmaf(swx[k], e * f.x, au);
            av = fmaf(swy[k], e * f.y, av);
        }
    }
    const float r = 1.f / sum;
    u = (au + bxv) * r;
    v = (av + byv) * r;
}

template <int K>
__global__ void __launch_bounds__(128, 4)
reg_tail_bulk_kernel(const float* __restrict__ dist, int dist_ld, const float2* __restrict__ flow,
                     const float* __restrict__ wx, const float* __restrict__ bx,
                     const float* __restrict__ wy, const float* __restrict__ by,
                     float2* __restrict__ flow_out, float* __restrict__ out_nchw, float final_scale,
                     int N, int H, int W) {
    constexpr int KK = K * K;
    extern __shared__ __align__(128) float sbuf[];       // [2][128 * dist_ld]
    __shared__ float swx[KK], swy[KK];
    __shared__ __align__(8) unsigned long long bar[2];
    for (int i = threadIdx.x; i < KK; i += 128) { swx[i] = wx[i]; swy[i] = wy[i]; }
    const float bxv = __ldg(bx), byv = __ldg(by);
    const long long HW = (long long)H * W, total = (long long)N * HW;
    const long long ntiles = (total + 127) / 128;
    const int tile_floats = 128 * dist_ld;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](long long tile, int buf) {
        const long long p0 = tile * 128;
        const unsigned bytes = (unsigned)(min(128LL, total - p0) * dist_ld * 4);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(&bar[buf])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_addr_u32(sbuf + buf * tile_floats)), "l"(dist + p0 * dist_ld), "r"(bytes),
                       "r"(smem_addr_u32(&bar[buf])) : "memory");
    };
    if (threadIdx.x == 0 && (long long)blockIdx.x < ntiles) issue(blockIdx.x, 0);
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const long long next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < ntiles) issue(next, buf ^ 1);      // buf^1 was released by the barrier below
        const unsigned parity = (unsigned)((it >> 1) & 1);
        unsigned done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_addr_u32(&bar[buf])), "r"(parity) : "memory");
        } while (!done);
        const long long p = tile * 128 + threadIdx.x;
        if (p < total) {
            const long long n = p / HW;
            const int x = (int)(p % W), y = (int)((p / W) % H);
            // own row -> registers through float4 shared loads
            const float4* row = reinterpret_cast<const float4*>(sbuf + buf * tile_floats + threadIdx.x * dist_ld);
            float d[(KK + 3) & ~3];
#pragma unroll
            for (int j = 0; j < ((KK + 3) >> 2); ++j) {
                const float4 q = row[j];
                d[4 * j] = q.x; d[4 * j + 1] = q.y; d[4 * j + 2] = q.z; d[4 * j + 3] = q.w;
            }
            float u, v;
            reg_tail_pixel<K>(d, flow, n, x, y, H, W, swx, swy, bxv, byv, u, v);
            flow_out[p] = make_float2(u, v);
            if (out_nchw) {
                const long long q = p - n * HW;
                out_nchw[(n * 2 + 0) * HW + q] = u * final_scale;
                out_nchw[(n * 2 + 1) * HW + q] = v * final_scale;
            }
        }
        __syncthreads();                                  // everybody is done with buf before it is refilled
    }
}

// Generic variant (any alignment / pitch): one thread per pixel reading its row directly.
template <int K>
__global__ void __launch_bounds__(128)
reg_tail_kernel(const float* __restrict__ dist, int dist_ld, const float2* __restrict__ flow,
                const float* __restrict__ wx, const float* __restrict__ bx,
                const float* __restrict__ wy, const float* __restrict__ by,
                float2* __restrict__ flow_out, float* __restrict__ out_nchw, float final_scale,
                int N, int H, int W) {
    constexpr int KK = K * K;
    __shared__ float swx[KK], swy[KK];
    for (int i = threadIdx.x; i < KK; i += blockDim.x) { swx[i] = wx[i]; swy[i] = wy[i]; }
    __syncthreads();
    const float bxv = __ldg(bx), byv = __ldg(by);
    const long long HW = (long long)H * W, total = (long long)N * HW;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        const long long n = p / HW;
        const int x = (int)(p % W), y = (int)((p / W) % H);
        float d[KK];
#pragma unroll
        for (int k = 0; k < KK; ++k) d[k] = __ldg(dist + p * dist_ld + k);
        float u, v;
        reg_tail_pixel<K>(d, flow, n, x, y, H, W, swx, swy, bxv, byv, u, v);
        flow_out[p] = make_float2(u, v);
        if (out_nchw) {
            const long long q = p - n * HW;
            out_nchw[(n * 2 + 0) * HW + q] = u * final_scale;
            out_nchw[(n * 2 + 1) * HW + q] = v * final_scale;
        }
    }
}

template <int K>
int launch_reg_tail(const float* dist, int dist_ld, const float2* fi, const float* wx, const float* bx, const float* wy,
                    const float* by, float2* fo, float* out_nchw, float final_scale, int N, int H, int W, cudaStream_t st) {
    const long long total = (long long)N * H * W;
    const bool bulk = !((uintptr_t)dist & 15) && !(dist_ld & 3) && dist_ld >= ((K * K + 3) & ~3) && dist_ld <= 64;
    if (bulk) {
        const int smem = 2 * 128 * dist_ld * 4;
        static unsigned long long configured = 0;
        {
            cudaError_t e = pivlfn_optin_smem(reg_tail_bulk_kernel<K>, 2 * 128 * 64 * 4, configured);
            if (e != cudaSuccess) return (int)e;
        }
        const long long ntiles = (total + 127) / 128;
        const long long cap = 4LL * pivlfn_num_sms();
        const int grid = (int)(ntiles < cap ? ntiles : cap);
        reg_tail_bulk_kernel<K><<<grid, 128, smem, st>>>(dist, dist_ld, fi, wx, bx, wy, by, fo, out_nchw, final_scale, N, H, W);
    } else {
        long long g = (total + 127) / 128;
        if (g > 148LL * 32) g = 148LL * 32;
        reg_tail_kernel<K><<<(int)g, 128, 0, st>>>(dist, dist_ld, fi, wx, bx, wy, by, fo, out_nchw, final_scale, N, H, W);
    }
    return 0;
}

// ---- F.interpolate(mode='bilinear', align_corners=False) on NCHW (inference.py:46-49,57-61) ------------------
// src = (dst + 0.5) * (in / out) - 0.5 clamped at 0, second tap clamped at in-1; channels of even index are
// multiplied by mul_even and odd ones by mul_odd (the u *= W/W', v *= H/H' of inference.py:60-61).
__global__ void resize_bilinear_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int NC, int H, int W,
                                            int Ho, int Wo, float mul_even, float mul_odd) {
    const float ry = (float)H / (float)Ho, rx = (float)W / (float)Wo;
    const long long total = (long long)NC * Ho * Wo;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wo);
        const long long t = i / Wo;
        const int oy = (int)(t % Ho);
        const long long c = t / Ho;
        const float sy = fmaxf(((float)oy + 0.5f) * ry - 0.5f, 0.f);
        const float sx = fmaxf(((float)ox + 0.5f) * rx - 0.5f, 0.f);
        const int y0 = min((int)sy, H - 1), x0 = min((int)sx, W - 1);
        const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
        const float ly = sy - (float)y0, lx = sx - (float)x0;
        const float* p = in + c * (long long)H * W;
        const float v = (1.f - ly) * ((1.f - lx) * __ldg(p + (long long)y0 * W + x0) + lx * __ldg(p + (long long)y0 * W + x1)) +
                        ly * ((1.f - lx) * __ldg(p + (long long)y1 * W + x0) + lx * __ldg(p + (long long)y1 * W + x1));
        out[i] = v * ((c & 1) ? mul_odd : mul_even);
    }
}

// ---- second half of a restated flow head (see pivlfn_conv1x1_pairs_tc): gather-sum of the K*K tap planes --------------
template <int K>
__global__ void __launch_bounds__(256)
flow_head_sum_kernel(const float2* __restrict__ planes, long long plane_pix, const float* __restrict__ bias,
                     const float* __restrict__ res, int res_ld, float* __restrict__ out, int out_ld, int N, int H, int W) {
    constexpr int P = K / 2;
    const long long HW = (long long)H * W, total = (long long)N * HW;
    const float b0 = bias ? __ldg(bias) : 0.f, b1 = bias ? __ldg(bias + 1) : 0.f;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(p % W), y = (int)((p / W) % H);
        float su = 0.f, sv = 0.f;
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            const int yy = y + ky - P;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const int xx = x + kx - P;
                if (xx >= 0 && xx < W) {
                    const float2 d = __ldg(planes + (long long)(ky * K + kx) * plane_pix + p + (long long)(ky - P) * W + (kx - P));
                    su += d.x; sv += d.y;
                }
            }
        }
        su += b0; sv += b1;
        if (res) { su += __ldg(res + p * res_ld); sv += __ldg(res + p * res_ld + 1); }
        out[p * out_ld] = su;
        out[p * out_ld + 1] = sv;
    }
}

inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148LL * 32;       // a few waves of the 148 SMs, grid-stride beyond
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

extern "C" int pivlfn_prep_images(float* img1, float* img2, float* out_nhwc4, float* out_pad, int B, int H, int W,
                                  const float* mean6, void* stream) {
    if (!img1 || !img2 || !out_nhwc4 || !mean6 || B <= 0 || H <= 0 || W <= 0) return PIVLFN_EINVAL;
    if (((uintptr_t)out_nhwc4 & 15) || ((uintptr_t)out_pad & 15)) return PIVLFN_EINVAL;
    const long long total = 2LL * B * H * W;
    prep_images_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        img1, img2, reinterpret_cast<float4*>(out_nhwc4), reinterpret_cast<float4*>(out_pad), W, B, H * W,
        mean6[0], mean6[1], mean6[2], mean6[3], mean6[4], mean6[5]);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

extern "C" int pivlfn_avgpool2(const float* in, float* out, int N, int H, int W, int C, void* stream) {
    if (!in || !out || N <= 0 || H <= 0 || W <= 0 || C <= 0 || (H & 1) || (W & 1)) return PIVLFN_EINVAL;
    const long long total = (long long)N * (H / 2) * (W / 2) * C;
    avgpool2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, N, H, W, C);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

extern "C" int pivlfn_deconv4x4s2_dw(const float* in, int in_ld, const float* w, float* out, int out_ld,
                                     int N, int H, int W, int C, void* stream) {
    if (!in || !w || !out || N <= 0 || H <= 0 || W <= 0 || C <= 0 || in_ld < C || out_ld < C) return PIVLFN_EINVAL;
    // float4 path: both views 16-byte aligned with room for the channel count rounded up to 4 (pad channels get 0)
    const bool vec = !((uintptr_t)in & 15) && !((uintptr_t)out & 15) && !(in_ld & 3) && !(out_ld & 3) &&
                     in_ld >= ((C + 3) & ~3) && out_ld >= ((C + 3) & ~3);
    if (vec && C <= DECONV_MAXC) {
        const long long total = (long long)N * H * W * ((C + 3) / 4);
        deconv4x4s2_dw_block_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, in_ld, w, out, out_ld, N, H, W, C);
    } else if (vec) {
        const long long total = (long long)N * 4 * H * W * ((C + 3) / 4);
        deconv4x4s2_dw_kernel<4><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, in_ld, w, out, out_ld, N, H, W, C);
    } else {
        const long long total = (long long)N * 4 * H * W * C;
        deconv4x4s2_dw_kernel<1><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, in_ld, w, out, out_ld, N, H, W, C);
    }
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

extern "C" int pivlfn_warp_nhwc(const float* in, int in_ld, const float* flow, float scale,
                                float* out, int out_ld, int N, int H, int W, int C, void* stream) {
    if (!in || !flow || !out || N <= 0 || H <= 0 || W <= 0 || C <= 0 || in_ld < C || out_ld < C) return PIVLFN_EINVAL;
    if (((uintptr_t)in & 15) || ((uintptr_t)out & 15) || (in_ld & 3) || (out_ld & 3) || ((uintptr_t)flow & 7))
        return PIVLFN_EINVAL;
    const long long ntiles = (long long)N * ((H + 7) / 8) * ((W + 7) / 8);
    warp_nhwc_kernel<<<(int)(ntiles < 148LL * 64 ? ntiles : 148LL * 64), 256, 0, (cudaStream_t)stream>>>(
        in, in_ld, reinterpret_cast<const float2*>(flow), scale, out, out_ld, N, H, W, C);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

extern "C" int pivlfn_copy_nhwc(const float* in, int in_ld, float* out, int out_ld, long long npix, int C,
                                void* stream) {
    if (!in || !out || npix <= 0 || C <= 0 || in_ld < C || out_ld < C) return PIVLFN_EINVAL;
    const bool vec = (C % 4 == 0) && !((uintptr_t)in & 15) && !((uintptr_t)out & 15) && !(in_ld & 3) && !(out_ld & 3);
    if (vec)
        copy_nhwc_kernel<true><<<grid_for(npix * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(in, in_ld, out, out_ld, npix, C);
    else
        copy_nhwc_kernel<false><<<grid_for(npix * C, 256), 256, 0, (cudaStream_t)stream>>>(in, in_ld, out, out_ld, npix, C);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

extern "C" int pivlfn_flow_mean_parts(void) { return MEAN_PARTS; }

extern "C" int pivlfn_flow_mean(const float* flow, float* partial, int N, int H, int W, void* stream) {
    if (!flow || !partial || N <= 0 || H <= 0 || W <= 0 || N > 65535 || ((uintptr_t)flow & 7)) return PIVLFN_EINVAL;
    flow_mean_kernel<<<dim3(MEAN_PARTS, N), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(flow), partial, H * W);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

extern "C" int pivlfn_reg_input(const float* img1, const float* img2, const float* flow, float scale,
                                const float* partial, float* out, int out_ld, int N, int H, int W, void* stream) {
    if (!img1 || !img2 || !flow || !partial || !out || N <= 0 || H <= 0 || W <= 0 || out_ld < 3) return PIVLFN_EINVAL;
    if (((uintptr_t)img1 & 15) || ((uintptr_t)img2 & 15) || ((uintptr_t)flow & 7)) return PIVLFN_EINVAL;
    const long long total = (long long)N * H * W;
    reg_input_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(img1), reinterpret_cast<const float4*>(img2),
        reinterpret_cast<const float2*>(flow), scale, partial, out, out_ld, N, H, W);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

extern "C" int pivlfn_reg_tail(const float* dist, int dist_ld, const float* flow_in,
                               const float* wx, const float* bx, const float* wy, const float* by,
                               float* flow_out, float* out_nchw, float final_scale,
                               int K, int N, int H, int W, void* stream) {
    if (!dist || !flow_in || !wx || !bx || !wy || !by || !flow_out || N <= 0 || H <= 0 || W <= 0) return PIVLFN_EINVAL;
    if (dist_ld < K * K || ((uintptr_t)flow_in & 7) || ((uintptr_t)flow_out & 7)) return PIVLFN_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const float2* fi = reinterpret_cast<const float2*>(flow_in);
    float2* fo = reinterpret_cast<float2*>(flow_out);
    int rc = 0;
    switch (K) {
        case 3: rc = launch_reg_tail<3>(dist, dist_ld, fi, wx, bx, wy, by, fo, out_nchw, final_scale, N, H, W, st); break;
        case 5: rc = launch_reg_tail<5>(dist, dist_ld, fi, wx, bx, wy, by, fo, out_nchw, final_scale, N, H, W, st); break;
        case 7: rc = launch_reg_tail<7>(dist, dist_ld, fi, wx, bx, wy, by, fo, out_nchw, final_scale, N, H, W, st); break;
        default: return PIVLFN_EINVAL;
    }
    if (rc) return rc;
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

extern "C" int pivlfn_resize_bilinear_nchw(const float* in, float* out, int NC, int H, int W, int Ho, int Wo,
                                           float mul_even, float mul_odd, void* stream) {
    if (!in || !out || NC <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return PIVLFN_EINVAL;
    const long long total = (long long)NC * Ho * Wo;
    resize_bilinear_nchw_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, NC, H, W, Ho, Wo,
                                                                                         mul_even, mul_odd);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

extern "C" int pivlfn_flow_head_sum(const float* planes, int K, const float* bias, const float* res, int res_ld,
                                    float* out, int out_ld, int N, int H, int W, void* stream) {
    if (!planes || !out || N <= 0 || H <= 0 || W <= 0 || out_ld < 2 || (res && res_ld < 2)) return PIVLFN_EINVAL;
    if ((uintptr_t)planes & 7) return PIVLFN_EINVAL;
    const long long total = (long long)N * H * W;
    const float2* pl = reinterpret_cast<const float2*>(planes);
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(total, 256);
    switch (K) {
        case 3: flow_head_sum_kernel<3><<<g, 256, 0, st>>>(pl, total, bias, res, res_ld, out, out_ld, N, H, W); break;
        case 5: flow_head_sum_kernel<5><<<g, 256, 0, st>>>(pl, total, bias, res, res_ld, out, out_ld, N, H, W); break;
        case 7: flow_head_sum_kernel<7><<<g, 256, 0, st>>>(pl, total, bias, res, res_ld, out, out_ld, N, H, W); break;
        default: return PIVLFN_EINVAL;
    }
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}
