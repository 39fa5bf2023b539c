// 49-displacement cost volume with the backwarp of f2 fused in, as a PERSISTENT, WARP-SPECIALISED kernel
// (the fine levels of the P16 pipeline: f1 = P16 slice of the Subpixel concat buffer, f2 = fp32 NHWC or P16, out = fp32 rows).
//
//   out[b, y, x, (dy+3)*7 + (dx+3)] = lrelu( (1/C) * sum_c f1[b, y*s, x*s, c] * warp(f2)[b, (y+dy)*s, (x+dx)*s, c] )
//   warp(f2)[p] = bilinear sample of f2 at p + scale * flow[p]  (zero outside)            src/models.py:169-184,
//                                                                                           src/correlation.py:36-104
//
// corr_nhwc_kernel (corr.cu) alternates "all threads gather" and "all threads multiply" phases behind __syncthreads and was
// bound by the un-overlapped gathers (ncu: 4.4 warps per issue stalled on the long scoreboard, issue slots 46 % busy) at
// 2.4x halo redundancy.  Here one CTA per SM walks 32 x 8 output tiles (halo redundancy 2.08x) and
//   * 10 PRODUCER warps compute the bilinear taps of the tile's 38 x 14 sample points once, then per 16-channel chunk gather
//     the four taps (12 independent 16-byte loads in flight per thread), blend, and store the warped f2 tile and the f1 tile
//     into a 3-stage shared-memory ring (mbarrier full / empty per stage, continuing across tiles);
//   * 14 CONSUMER warps (displacement row dy x tile half) own 4 adjacent pixels x 7 dx = 28 accumulators per thread for the
//     whole channel loop and read the ring: 14 shared-memory float4 loads per 112 FMA, conflict-free (pixel pitch 20 floats +
//     a per-lane rotation of the channel-quad order);
//   so the gathers of chunk k+1..k+2 are in flight while chunk k is multiplied.  Results leave from registers.
#include "common.cuh"
#include "p16.cuh"

namespace {

constexpr int TX = 32, TY = 8, CK = 16, NQ = CK / 4, PITCH = 20;
constexpr int SW = TX + 6, SH = TY + 6, NPIX2 = SW * SH;          // 38 x 14 = 532 sample points of f2
constexpr int NPIX1 = TX * TY;
constexpr int S1_FLOATS = NPIX1 * PITCH, S2_FLOATS = NPIX2 * PITCH;
constexpr int STAGE_FLOATS = S1_FLOATS + S2_FLOATS;               // 15760 floats = 63040 B
constexpr int NST = 3;
constexpr int NCONS_WARPS = 14, NPROD_WARPS = 10;
constexpr int NCONS = NCONS_WARPS * 32, NPROD = NPROD_WARPS * 32, NTHREADS = NCONS + NPROD;
constexpr int SMEM_BYTES = NST * STAGE_FLOATS * 4 + NPIX2 * (16 + 8);

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(b)), "r"(n)); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(saddr(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(saddr(b)), "r"(parity) : "memory");
    } while (!done);
}

__device__ __forceinline__ float4 ld_f1_quad_p16(const float* pixel_row, int c) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(pixel_row) + (c >> 4) * 64 + (c & 15) * 2;
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(p));
    const uint2 l = __ldg(reinterpret_cast<const uint2*>(p + 32));
    float4 v;
    p16::decode2(h.x, l.x, v.x, v.y);
    p16::decode2(h.y, l.y, v.z, v.w);
    return v;
}

template <bool F2P>
__global__ void __launch_bounds__(NTHREADS, 1)
corr_ps_kernel(const float* __restrict__ f1, int f1_ld, const float* __restrict__ f2, int f2_ld,
               const float* __restrict__ flow, float fscale, float* __restrict__ out, int out_ld,
               int N, int C, int H, int W, int Ho, int Wo, int s, int lrelu, int tiles_x, int tiles_y, int total) {
    extern __shared__ __align__(16) float sbuf[];
    float4* const tapw = reinterpret_cast<float4*>(sbuf + NST * STAGE_FLOATS);
    int2* const tapxy = reinterpret_cast<int2*>(tapw + NPIX2);
    __shared__ __align__(8) uint64_t full[NST], empty[NST];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nch = C / CK;
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { mb_init(&full[i], NPROD_WARPS); mb_init(&empty[i], NCONS_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= NCONS_WARPS) {
        // ======================================= producers =======================================
        const int pt = tid - NCONS;
        int st = 0;
        uint32_t use = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
            const int x0 = tx * TX, y0 = ty * TY;
            const size_t img = (size_t)n * H * W;
            // every producer is done gathering the previous tile (the tap table is single-buffered)
            asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
            for (int p = pt; p < NPIX2; p += NPROD) {
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
            asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
            for (int ch = 0; ch < nch; ++ch) {
                const int c0 = ch * CK;
                mb_wait(&empty[st], (use & 1) ^ 1);
                float* const s1 = sbuf + st * STAGE_FLOATS;
                float* const s2 = s1 + S1_FLOATS;
                // ---- warped f2 tile: 532 sample points x 4 channel quads, three items (12 gathers) in flight per thread ----
                constexpr int UNR = 3;
                for (int base = pt; base < NPIX2 * NQ; base += UNR * NPROD) {
                    float4 u[UNR][4];
                    float4 wv[UNR];
#pragma unroll
                    for (int e = 0; e < UNR; ++e) {
                        const int item = base + e * NPROD;
                        const bool ok = item < NPIX2 * NQ;
                        const int q = item & (NQ - 1), p = ok ? (item >> 2) : 0;
                        const int c = c0 + q * 4;
                        wv[e] = ok ? tapw[p] : make_float4(0.f, 0.f, 0.f, 0.f);
                        const int2 xy = tapxy[p];
                        const float wgt[4] = {wv[e].x, wv[e].y, wv[e].z, wv[e].w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            u[e][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (wgt[k] != 0.f) {          // taps outside the frame are never dereferenced
                                const float* row = f2 + (img + (size_t)(xy.y + (k >> 1)) * W + (xy.x + (k & 1))) * f2_ld;
                                u[e][k] = F2P ? ld_f1_quad_p16(row, c) : __ldg(reinterpret_cast<const float4*>(row + c));
                            }
                        }
                    }
#pragma unroll
                    for (int e = 0; e < UNR; ++e) {
                        const int item = base + e * NPROD;
                        if (item < NPIX2 * NQ) {
                            const float wgt[4] = {wv[e].x, wv[e].y, wv[e].z, wv[e].w};
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                v.x = fmaf(wgt[k], u[e][k].x, v.x); v.y = fmaf(wgt[k], u[e][k].y, v.y);
                                v.z = fmaf(wgt[k], u[e][k].z, v.z); v.w = fmaf(wgt[k], u[e][k].w, v.w);
                            }
                            *reinterpret_cast<float4*>(&s2[(item >> 2) * PITCH + (item & (NQ - 1)) * 4]) = v;
                        }
                    }
                }
                // ---- f1 tile: 256 pixels x 4 quads (P16 -> fp32) ----
                for (int item = pt; item < NPIX1 * NQ; item += NPROD) {
                    const int q = item & (NQ - 1), p = item >> 2;
                    const int px = x0 + (p & (TX - 1)), py = y0 + (p >> 5);
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (px < Wo && py < Ho) v = ld_f1_quad_p16(f1 + (img + (size_t)(py * s) * W + px * s) * f1_ld, c0 + q * 4);
                    *reinterpret_cast<float4*>(&s1[p * PITCH + q * 4]) = v;
                }
                __syncwarp();
                if (lane == 0) mb_arrive(&full[st]);          // release: this warp's stores are visible to whoever acquires
                if (++st == NST) { st = 0; ++use; }
            }
        }
    } else {
        // ======================================= consumers =======================================
        const int dy = warp >> 1, half = warp & 1;
        const int seg = lane & 7, tyy = (lane >> 3) + 4 * half;
        const int rot = seg >> 1;
        const float inv = 1.f / (float)C;
        int st = 0;
        uint32_t use = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
            float acc[4][7];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int d = 0; d < 7; ++d) acc[i][d] = 0.f;
            for (int ch = 0; ch < nch; ++ch) {
                mb_wait(&full[st], use & 1);
                const float* a0 = sbuf + st * STAGE_FLOATS + (tyy * TX + seg * 4) * PITCH;
                const float* b0 = sbuf + st * STAGE_FLOATS + S1_FLOATS + ((tyy + dy) * SW + seg * 4) * PITCH;
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
                __syncwarp();
                if (lane == 0) mb_arrive(&empty[st]);
                if (++st == NST) { st = 0; ++use; }
            }
            // ---- results straight from registers: 7 consecutive floats per pixel (the dy row of its 49-vector) ----
            const int oy = ty * TY + tyy;
            if (oy < Ho) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ox = tx * TX + seg * 4 + i;
                    if (ox < Wo) {
                        float* o = out + ((size_t)n * Ho * Wo + (size_t)oy * Wo + ox) * out_ld + dy * 7;
#pragma unroll
                        for (int d = 0; d < 7; ++d) {
                            const float v = acc[i][d] * inv;
                            o[d] = lrelu ? lrelu_f(v) : v;
                        }
                    }
                }
            }
        }
    }
}

}  // namespace

// fp32 output rows; f1 P16; f2 fp32 NHWC or P16; C % 16 == 0.  Returns PIVLFN_EUNSUPPORTED for other combinations (the caller
// falls back to corr_nhwc_kernel).
int pivlfn_corr_ps_launch(const void* f1, int f1_ld, const void* f2, int f2_ld, int f2_p16, const float* flow, float flow_scale,
                          float* out, int out_ld, int N, int H, int W, int C, int stride, int lrelu, cudaStream_t st) {
    if (C % CK) return PIVLFN_EUNSUPPORTED;
    const int Ho = cdiv(H, stride), Wo = cdiv(W, stride);
    const int tiles_x = cdiv(Wo, TX), tiles_y = cdiv(Ho, TY);
    const long long total = (long long)tiles_x * tiles_y * N;
    if (total > 0x7FFFFFFFLL) return PIVLFN_EINVAL;
    const int nsm = pivlfn_num_sms();
    const int grid = (int)(total < nsm ? total : nsm);
    static unsigned long long cfg0 = 0, cfg1 = 0;
    const float* a = reinterpret_cast<const float*>(f1);
    const float* b = reinterpret_cast<const float*>(f2);
    if (f2_p16) {
        cudaError_t e = pivlfn_optin_smem(corr_ps_kernel<true>, SMEM_BYTES, cfg1);
        if (e != cudaSuccess) return (int)e;
        corr_ps_kernel<true><<<grid, NTHREADS, SMEM_BYTES, st>>>(a, f1_ld, b, f2_ld, flow, flow_scale, out, out_ld, N, C, H, W, Ho, Wo,
                                                                 stride, lrelu, tiles_x, tiles_y, (int)total);
    } else {
        cudaError_t e = pivlfn_optin_smem(corr_ps_kernel<false>, SMEM_BYTES, cfg0);
        if (e != cudaSuccess) return (int)e;
        corr_ps_kernel<false><<<grid, NTHREADS, SMEM_BYTES, st>>>(a, f1_ld, b, f2_ld, flow, flow_scale, out, out_ld, N, C, H, W, Ho, Wo,
                                                                  stride, lrelu, tiles_x, tiles_y, (int)total);
    }
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}
