// 49-displacement cost volume with the backwarp of f2 fused in, SOFTWARE-PIPELINED (the fine levels of the P16 pipeline:
// f1 = P16 slice of the Subpixel concat buffer, f2 = fp32 NHWC or P16, out = fp32 rows).
//
//   out[b, y, x, (dy+3)*7 + (dx+3)] = lrelu( (1/C) * sum_c f1[b, y*s, x*s, c] * warp(f2)[b, (y+dy)*s, (x+dx)*s, c] )
//   warp(f2)[p] = bilinear sample of f2 at p + scale * flow[p]  (zero outside)            src/models.py:169-184,
//                                                                                           src/correlation.py:36-104
//
// corr_nhwc_kernel (corr.cu) alternates "all threads gather a 32-channel chunk" and "all threads multiply" behind
// __syncthreads: ncu showed it waiting on the gathers (4.4 warps per issue stalled on the long scoreboard, issue slots 46 %
// busy).  Same tile (16 x 8 outputs, 7 warps = displacement rows, 4 pixels x 7 dx per thread) and the same conflict-free
// shared-memory reads here, but the channel loop runs in 8-channel chunks and is software-pipelined through REGISTERS: the 14
// gathers of chunk k+1 (3 f2 items x 4 taps + 2 f1 items) are issued BEFORE the FMAs of chunk k and consumed after them, so
// every thread always has its next chunk in flight while it multiplies; two CTAs per SM (128 registers per thread).
// (A warp-specialised producer / consumer variant with a 3-stage ring was measured first: 1.15 ms at level 1 against 0.80 ms
// for corr_nhwc_kernel -- 10 producer warps cannot keep enough gathers in flight.)
#include "common.cuh"
#include "p16.cuh"

namespace {

constexpr int TX = 16, TY = 8, CK = 8, NQ = CK / 4, PITCH = 12;
constexpr int SW = TX + 6, SH = TY + 6, NPIX2 = SW * SH;          // 22 x 14 = 308 sample points of f2
constexpr int S1W = TX + 2;                                         // padded f1 row pitch (pixels)
constexpr int NTHREADS = 224;
constexpr int S1_FLOATS = S1W * TY * PITCH, S2_FLOATS = NPIX2 * PITCH;
constexpr int OLD = 52;                                             // staged output row: 49 displacements + pad
constexpr int N2 = (NPIX2 * NQ + NTHREADS - 1) / NTHREADS;        // 3 f2 items per thread and chunk
constexpr int N1 = (TX * TY * NQ + NTHREADS - 1) / NTHREADS;      // 2 f1 items
constexpr int SMEM_FLOATS = (S1_FLOATS + S2_FLOATS) > TX * TY * OLD ? (S1_FLOATS + S2_FLOATS) : TX * TY * OLD;
constexpr int SMEM_BYTES = SMEM_FLOATS * 4 + NPIX2 * (16 + 8);

__device__ __forceinline__ float4 ld_quad_p16(const float* pixel_row, int c) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(pixel_row) + (c >> 4) * 64 + (c & 15) * 2;
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(p));
    const uint2 l = __ldg(reinterpret_cast<const uint2*>(p + 32));
    float4 v;
    p16::decode2(h.x, l.x, v.x, v.y);
    p16::decode2(h.y, l.y, v.z, v.w);
    return v;
}

template <bool F2P>
__global__ void __launch_bounds__(NTHREADS, 2)
corr_sp_kernel(const float* __restrict__ f1, int f1_ld, const float* __restrict__ f2, int f2_ld,
               const float* __restrict__ flow, float fscale, float* __restrict__ out, int out_ld,
               int C, int H, int W, int Ho, int Wo, int s, int lrelu) {
    extern __shared__ __align__(16) float sbuf[];
    float* const s1 = sbuf;
    float* const s2 = sbuf + S1_FLOATS;
    float4* const tapw = reinterpret_cast<float4*>(sbuf + SMEM_FLOATS);
    int2* const tapxy = reinterpret_cast<int2*>(tapw + NPIX2);
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int tid = threadIdx.x;
    const size_t img = (size_t)n * H * W;
    const int nch = C / CK;

    // ---- bilinear taps of the 308 f2 sample points, once per CTA ----
    for (int p = tid; p < NPIX2; p += NTHREADS) {
        const int i = p % SW, j = p / SW;
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
    __syncthreads();

    // The tap tables stay in shared memory and are re-read per chunk (two LDS per item): keeping the 12 tap pointers and weights
    // of a thread's items in registers next to the 56 staged values and the 28 accumulators spilled at 128 registers.
    float4 u2[N2][4], u1[N1];
    auto issue = [&](int ch) {                     // gathers of chunk ch -> registers (nothing is consumed here)
        const int c0 = ch * CK;
#pragma unroll
        for (int e = 0; e < N2; ++e) {
            const int item = tid + e * NTHREADS;
            const bool ok = item < NPIX2 * NQ;
            const int q = item & (NQ - 1), p = ok ? item >> 1 : 0;
            const float4 wv = tapw[p];
            const int2 xy = tapxy[p];
            const float wgt[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u2[e][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok && wgt[k] != 0.f) {              // taps outside the frame are never dereferenced
                    const float* row = f2 + (img + (size_t)(xy.y + (k >> 1)) * W + (xy.x + (k & 1))) * f2_ld;
                    u2[e][k] = F2P ? ld_quad_p16(row, c0 + q * 4) : __ldg(reinterpret_cast<const float4*>(row + c0 + q * 4));
                }
            }
        }
#pragma unroll
        for (int e = 0; e < N1; ++e) {
            const int item = tid + e * NTHREADS;
            const int q = item & (NQ - 1), p = item >> 1;
            const int px = x0 + p % TX, py = y0 + p / TX;
            u1[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (item < TX * TY * NQ && px < Wo && py < Ho)
                u1[e] = ld_quad_p16(f1 + (img + (size_t)(py * s) * W + px * s) * f1_ld, c0 + q * 4);
        }
    };
    auto commit = [&]() {                          // blend and store the staged chunk into the shared-memory tiles
#pragma unroll
        for (int e = 0; e < N2; ++e) {
            const int item = tid + e * NTHREADS;
            if (item < NPIX2 * NQ) {
                const int q = item & (NQ - 1), p = item >> 1;
                const float4 wv = tapw[p];
                const float wgt[4] = {wv.x, wv.y, wv.z, wv.w};
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    v.x = fmaf(wgt[k], u2[e][k].x, v.x); v.y = fmaf(wgt[k], u2[e][k].y, v.y);
                    v.z = fmaf(wgt[k], u2[e][k].z, v.z); v.w = fmaf(wgt[k], u2[e][k].w, v.w);
                }
                *reinterpret_cast<float4*>(&s2[p * PITCH + q * 4]) = v;
            }
        }
#pragma unroll
        for (int e = 0; e < N1; ++e) {
            const int item = tid + e * NTHREADS;
            if (item < TX * TY * NQ) {
                const int q = item & (NQ - 1), p = item >> 1;
                *reinterpret_cast<float4*>(&s1[((p / TX) * S1W + p % TX) * PITCH + q * 4]) = u1[e];
            }
        }
    };

    const int lane = tid & 31, dy = tid >> 5;            // warp = displacement row
    const int seg = lane & 3, ty = lane >> 2;            // 4-pixel segment of row ty
    const int rot = seg >> 1;                            // per-lane rotation of the quad order: conflict-free float4 reads
    float acc[4][7];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int d = 0; d < 7; ++d) acc[i][d] = 0.f;

    issue(0);
    commit();
    __syncthreads();
    const float* a0 = &s1[(ty * S1W + seg * 4) * PITCH];
    const float* b0 = &s2[((ty + dy) * SW + seg * 4) * PITCH];
    for (int ch = 0; ch < nch; ++ch) {
        if (ch + 1 < nch) issue(ch + 1);                 // in flight during the FMAs below
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
            const int q = ((qi + rot) & (NQ - 1)) * 4;
            float4 a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * PITCH + q);
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                const float4 v = *reinterpret_cast<const float4*>(b0 + j * PITCH + q);
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
        __syncthreads();                                 // everyone is done reading chunk ch
        if (ch + 1 < nch) {
            commit();
            __syncthreads();
        }
    }
    // ---- stage the 128 x 49 results in shared memory, then store whole pixel rows ----
    const float inv = 1.f / (float)C;
    float* so = sbuf;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int d = 0; d < 7; ++d) {
            const float v = acc[i][d] * inv;
            so[(ty * TX + seg * 4 + i) * OLD + dy * 7 + d] = lrelu ? lrelu_f(v) : v;
        }
    __syncthreads();
    const bool vec = out_ld == OLD && !((uintptr_t)out & 15);
    if (vec) {
        for (int item = tid; item < TX * TY * (OLD / 4); item += NTHREADS) {
            const int p = item / (OLD / 4), q = item % (OLD / 4);
            const int ox = x0 + p % TX, oy = y0 + p / TX;
            if (ox < Wo && oy < Ho) {
                float4 v = *reinterpret_cast<const float4*>(&so[p * OLD + q * 4]);
                if (q == OLD / 4 - 1) { v.y = 0.f; v.z = 0.f; v.w = 0.f; }
                *reinterpret_cast<float4*>(out + ((size_t)n * Ho * Wo + (size_t)oy * Wo + ox) * out_ld + q * 4) = v;
            }
        }
    } else {
        for (int item = tid; item < TX * TY * 49; item += NTHREADS) {
            const int p = item / 49, k = item % 49;
            const int ox = x0 + p % TX, oy = y0 + p / TX;
            if (ox < Wo && oy < Ho) out[((size_t)n * Ho * Wo + (size_t)oy * Wo + ox) * out_ld + k] = so[p * OLD + k];
        }
    }
}

}  // namespace

// fp32 output rows; f1 P16; f2 fp32 NHWC or P16; C % 8 == 0.  Returns PIVLFN_EUNSUPPORTED for other shapes (the caller falls back to
// corr_nhwc_kernel).
int pivlfn_corr_sp_launch(const void* f1, int f1_ld, const void* f2, int f2_ld, int f2_p16, const float* flow, float flow_scale,
                          float* out, int out_ld, int N, int H, int W, int C, int stride, int lrelu, cudaStream_t st) {
    if (C % CK || N > 65535) return PIVLFN_EUNSUPPORTED;
    const int Ho = cdiv(H, stride), Wo = cdiv(W, stride);
    dim3 grid(cdiv(Wo, TX), cdiv(Ho, TY), N);
    static unsigned long long cfg0 = 0, cfg1 = 0;
    const float* a = reinterpret_cast<const float*>(f1);
    const float* b = reinterpret_cast<const float*>(f2);
    if (f2_p16) {
        cudaError_t e = pivlfn_optin_smem(corr_sp_kernel<true>, SMEM_BYTES, cfg1);
        if (e != cudaSuccess) return (int)e;
        corr_sp_kernel<true><<<grid, NTHREADS, SMEM_BYTES, st>>>(a, f1_ld, b, f2_ld, flow, flow_scale, out, out_ld, C, H, W, Ho, Wo, stride, lrelu);
    } else {
        cudaError_t e = pivlfn_optin_smem(corr_sp_kernel<false>, SMEM_BYTES, cfg0);
        if (e != cudaSuccess) return (int)e;
        corr_sp_kernel<false><<<grid, NTHREADS, SMEM_BYTES, st>>>(a, f1_ld, b, f2_ld, flow, flow_scale, out, out_ld, C, H, W, Ho, Wo, stride, lrelu);
    }
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}
