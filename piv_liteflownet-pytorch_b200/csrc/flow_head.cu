// Flow heads: the last layer of conv_M / conv_S (src/models.py:161,205 and LiteFlowNet2 :499,547), a KxK convolution
// (K = 7, 5 or 3) from 32 channels to the 2 flow components, no activation, plus the residual flow ("+ xflow",
// src/models.py:186,216).
//
// With N = 2 output channels this layer is a poor tensor-core shape (2 useful columns of a 16-column MMA, one tiny
// weight tile per tap), and restating it as a 1x1 convolution to 2*K*K tap planes + a gather costs 784 bytes of HBM
// traffic per pixel.  It is 3136 FMA per pixel at K = 7 -- small enough for the CUDA cores in exact fp32, if the inner
// loop is FMA-bound rather than shared-memory-bound:
//
//   CTA  = 32 x 8 output pixels, 128 threads; the (32+K-1) x (8+K-1) input halo tile of 16 channels at a time and all
//          weights live in shared memory (55 KB at K = 7 -> four CTAs per SM cover each other's load phases)
//   warp = (channel group, row quad); lane = output column.  A thread owns 4 vertically adjacent outputs x 2 flow
//          components for 8 of the 16 channels of a phase: per (kx, channel quad) it loads the 4+K-1 input pixels of its
//          column once (float4 = 4 channels; consecutive lanes = consecutive pixels, pitch 20 floats: conflict-free)
//          and the K x 8 weights as warp-uniform broadcasts, then issues 4*K*8 FMAs -> ~9 FMA per shared-memory load.
//   The two channel halves are added through shared memory in a fixed order (deterministic).
#include "common.cuh"

namespace {

constexpr int FH_TX = 32, FH_TY = 8, FH_C = 32, FH_CP = 16, FH_PITCH = 20, FH_THREADS = 128;

template <int K>
__global__ void __launch_bounds__(FH_THREADS, 4)
flow_head_kernel(const float* __restrict__ x, int x_ld, const float* __restrict__ w, const float* __restrict__ bias,
                 const float* __restrict__ res, int res_ld, float* __restrict__ out, int out_ld,
                 float* __restrict__ out2, int out2_ld, int H, int W, int tiles_x, int tiles_y) {
    constexpr int R = K / 2, SW = FH_TX + K - 1, SH = FH_TY + K - 1, NPIX = SW * SH, NV = 4 + K - 1;
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                              // [NPIX][FH_PITCH]
    float* ws = smem + NPIX * FH_PITCH;            // [K*K][32 c][2 co]
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int bx = tile % tiles_x, by = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
    const int x0 = bx * FH_TX, y0 = by * FH_TY;
    const size_t img = (size_t)n * H * W;

    for (int i = tid; i < K * K * FH_C * 2 / 4; i += FH_THREADS)
        reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
    const int lane = tid & 31, wp = tid >> 5;
    const int yq = wp & 1, half = wp >> 1;
    float acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;

    // two phases of 16 input channels: the tile of one phase is 42.6 KB, so four CTAs fit on an SM and cover each other's
    // load phases; inside a phase warp pair `half` takes channel quads 2*half, 2*half + 1
#pragma unroll 1
    for (int ph = 0; ph < FH_C / FH_CP; ++ph) {
        if (ph) __syncthreads();                       // everyone is done reading the previous phase's tile
        // halo tile: 4 lanes per pixel (4 float4 = 16 channels), four independent loads in flight per thread
        constexpr int NITEM = NPIX * 4;
        for (int base = tid; base < NITEM; base += 4 * FH_THREADS) {
            float4 v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int item = base + e * FH_THREADS;
                v[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (item < NITEM) {
                    const int p = item >> 2, q = item & 3;
                    const int gy = y0 + p / SW - R, gx = x0 + p % SW - R;
                    if (gy >= 0 && gy < H && gx >= 0 && gx < W)
                        v[e] = __ldg(reinterpret_cast<const float4*>(x + (img + (size_t)gy * W + gx) * x_ld) + ph * 4 + q);
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int item = base + e * FH_THREADS;
                if (item < NITEM) *reinterpret_cast<float4*>(&xs[(item >> 2) * FH_PITCH + (item & 3) * 4]) = v[e];
            }
        }
        __syncthreads();

#pragma unroll 1
        for (int kx = 0; kx < K; ++kx) {
#pragma unroll 1
            for (int qi = 0; qi < 2; ++qi) {
                const int ql = half * 2 + qi;          // channel quad inside the phase
                const int cq = ph * 4 + ql;            // channel quad of the layer
                float4 xv[NV];
#pragma unroll
                for (int r = 0; r < NV; ++r)
                    xv[r] = *reinterpret_cast<const float4*>(&xs[((yq * 4 + r) * SW + lane + kx) * FH_PITCH + ql * 4]);
#pragma unroll
                for (int ky = 0; ky < K; ++ky) {
                    const float4 w0 = *reinterpret_cast<const float4*>(&ws[(ky * K + kx) * 64 + cq * 8]);       // c0: (u,v), c1: (u,v)
                    const float4 w1 = *reinterpret_cast<const float4*>(&ws[(ky * K + kx) * 64 + cq * 8 + 4]);   // c2, c3
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 v = xv[i + ky];
                        acc[i][0] = fmaf(v.x, w0.x, acc[i][0]); acc[i][1] = fmaf(v.x, w0.y, acc[i][1]);
                        acc[i][0] = fmaf(v.y, w0.z, acc[i][0]); acc[i][1] = fmaf(v.y, w0.w, acc[i][1]);
                        acc[i][0] = fmaf(v.z, w1.x, acc[i][0]); acc[i][1] = fmaf(v.z, w1.y, acc[i][1]);
                        acc[i][0] = fmaf(v.w, w1.z, acc[i][0]); acc[i][1] = fmaf(v.w, w1.w, acc[i][1]);
                    }
                }
            }
        }
    }
    // ---- add the two channel halves (half 1 -> shared memory -> half 0), bias, residual, store ------------------
    __syncthreads();                                 // everyone is done reading xs
    float2* part = reinterpret_cast<float2*>(xs);    // [256 pixels]
    if (half == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) part[(yq * 4 + i) * FH_TX + lane] = make_float2(acc[i][0], acc[i][1]);
    }
    __syncthreads();
    if (half == 0) {
        const float b0 = bias ? __ldg(bias) : 0.f, b1 = bias ? __ldg(bias + 1) : 0.f;
        const int gx = x0 + lane;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gy = y0 + yq * 4 + i;
            if (gx < W && gy < H) {
                const float2 o = part[(yq * 4 + i) * FH_TX + lane];
                float u = (acc[i][0] + o.x) + b0, v = (acc[i][1] + o.y) + b1;
                const size_t pix = img + (size_t)gy * W + gx;
                if (res) { u += __ldg(res + pix * res_ld); v += __ldg(res + pix * res_ld + 1); }
                out[pix * out_ld] = u;
                out[pix * out_ld + 1] = v;
                if (out2) { out2[pix * out2_ld] = u; out2[pix * out2_ld + 1] = v; }
            }
        }
    }
}

template <int K>
int launch_flow_head(const float* x, int x_ld, const float* w, const float* bias, const float* res, int res_ld,
                     float* out, int out_ld, float* out2, int out2_ld, int N, int H, int W, cudaStream_t st) {
    constexpr int SW = FH_TX + K - 1, SH = FH_TY + K - 1;
    constexpr int smem = (SW * SH * FH_PITCH + K * K * FH_C * 2) * 4;
    static unsigned long long configured = 0;
    {
        cudaError_t e = pivlfn_optin_smem(flow_head_kernel<K>, smem, configured);
        if (e != cudaSuccess) return (int)e;
    }
    const int tiles_x = cdiv(W, FH_TX), tiles_y = cdiv(H, FH_TY);
    const long long grid = (long long)tiles_x * tiles_y * N;
    if (grid > 0x7FFFFFFFLL) return PIVLFN_EINVAL;
    flow_head_kernel<K><<<(int)grid, FH_THREADS, smem, st>>>(x, x_ld, w, bias, res, res_ld, out, out_ld, out2, out2_ld, H, W, tiles_x, tiles_y);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

}  // namespace

extern "C" int pivlfn_flow_head(const float* x, int x_ld, int N, int H, int W, int Cin, const float* w, const float* bias,
                                const float* res, int res_ld, float* out, int out_ld, float* out2, int out2_ld, int K,
                                void* stream) {
    if (!x || !w || !out || N <= 0 || H <= 0 || W <= 0) return PIVLFN_EINVAL;
    if (Cin != FH_C) return PIVLFN_EUNSUPPORTED;
    if (((uintptr_t)x & 15) || (x_ld & 3) || x_ld < Cin || ((uintptr_t)w & 15) || out_ld < 2 || (res && res_ld < 2) || (out2 && out2_ld < 2)) return PIVLFN_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    switch (K) {
        case 3: return launch_flow_head<3>(x, x_ld, w, bias, res, res_ld, out, out_ld, out2, out2_ld, N, H, W, st);
        case 5: return launch_flow_head<5>(x, x_ld, w, bias, res, res_ld, out, out_ld, out2, out2_ld, N, H, W, st);
        case 7: return launch_flow_head<7>(x, x_ld, w, bias, res, res_ld, out, out_ld, out2, out2_ld, N, H, W, st);
        default: return PIVLFN_EINVAL;
    }
}
