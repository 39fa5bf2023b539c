// Generic NHWC convolution on the CUDA cores (fp32 FFMA), implicit GEMM:
//   M = output pixels (N*Ho*Wo), N = Cout, K = KH*KW*Cin.
// This is the exact-fp32 path for every layer shape the tcgen05 kernel does not take
// (7x7 stem with Cin=3, stride-2 convs, 1x1 convs, KxK -> 2 flow heads, separable (K,1)/(1,K)
// distance convs) and the cross-check for the tensor-core kernel in the tests.
// Replaces torch.nn.Conv2d + LeakyReLU of src/models.py:70-106,123-126,154-163,197-207,228-272.
#include "common.cuh"

namespace {

constexpr int BK = 16;       // K-chunk: channels of one filter tap per stage
constexpr int NTHREADS = 256;

struct ConvArgs {
    const float* x; int x_ld;
    int N, H, W, Cin;
    const float* w; int CoutP;
    const float* bias;
    float* y; int y_ld; int Cout;
    int KH, KW, stride, lrelu;
    const float* res; int res_ld;
    int Ho, Wo;
    long long M;
};

// BM x BN output tile per CTA, TM x TN outputs per thread.  (BM/TM)*(BN/TN) == NTHREADS.
template <int BM, int BN, int TM, int TN, bool VEC>
__global__ void __launch_bounds__(NTHREADS)
conv_simt_kernel(const ConvArgs a) {
    static_assert((BM / TM) * (BN / TN) == NTHREADS, "thread tiling");
    static_assert(TM % 4 == 0, "TM multiple of 4");
    constexpr int BMP = BM + 4;                 // row pitch of As (keeps float4 alignment)
    __shared__ __align__(16) float As[BK][BMP]; // k-major: As[k][m]
    __shared__ __align__(16) float Bs[BK][BN];

    const int tid = threadIdx.x;
    const long long m_base = (long long)blockIdx.x * BM;
    const int n_base = blockIdx.y * BN;

    // ---- A-load assignment: BM pixels x (BK/4) channel quads --------------------------------
    constexpr int A_ITEMS = BM * (BK / 4);
    constexpr int A_PER_THREAD = (A_ITEMS + NTHREADS - 1) / NTHREADS;
    int a_n[A_PER_THREAD], a_oy[A_PER_THREAD], a_ox[A_PER_THREAD];
    bool a_ok[A_PER_THREAD];
#pragma unroll
    for (int i = 0; i < A_PER_THREAD; ++i) {
        int item = tid + i * NTHREADS;
        int ml = item >> 2;                     // BK/4 == 4 quads per pixel
        long long m = m_base + ml;
        a_ok[i] = (item < A_ITEMS) && (m < a.M);
        long long mm = a_ok[i] ? m : 0;
        int ox = (int)(mm % a.Wo);
        long long t = mm / a.Wo;
        a_oy[i] = (int)(t % a.Ho);
        a_n[i] = (int)(t / a.Ho);
        a_ox[i] = ox;
    }
    // ---- B-load assignment: BK rows x BN/4 quads -----------------------------------------------
    constexpr int B_ITEMS = BK * (BN / 4);
    constexpr int B_PER_THREAD = (B_ITEMS + NTHREADS - 1) / NTHREADS;

    const int tm = (tid / (BN / TN)) * TM;      // first pixel of this thread inside the tile
    const int tn = (tid % (BN / TN)) * TN;      // first output channel inside the tile

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int ph = a.KH / 2, pw = a.KW / 2;
    const int nchunk = (a.Cin + BK - 1) / BK;

    for (int ky = 0; ky < a.KH; ++ky) {
        for (int kx = 0; kx < a.KW; ++kx) {
            const int tap = ky * a.KW + kx;
            for (int ch = 0; ch < nchunk; ++ch) {
                const int c0 = ch * BK;
                // ---- stage A ---------------------------------------------------------------
#pragma unroll
                for (int i = 0; i < A_PER_THREAD; ++i) {
                    int item = tid + i * NTHREADS;
                    if (item < A_ITEMS) {
                        int ml = item >> 2, q = item & 3;
                        int c = c0 + q * 4;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        int iy = a_oy[i] * a.stride + ky - ph;
                        int ix = a_ox[i] * a.stride + kx - pw;
                        if (a_ok[i] && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W && c < a.Cin) {
                            const float* p = a.x + ((long long)(a_n[i] * a.H + iy) * a.W + ix) * a.x_ld + c;
                            if (VEC && c + 3 < a.Cin) {
                                v = __ldg(reinterpret_cast<const float4*>(p));
                            } else {
                                v.x = __ldg(p);
                                if (c + 1 < a.Cin) v.y = __ldg(p + 1);
                                if (c + 2 < a.Cin) v.z = __ldg(p + 2);
                                if (c + 3 < a.Cin) v.w = __ldg(p + 3);
                            }
                        }
                        As[q * 4 + 0][ml] = v.x;
                        As[q * 4 + 1][ml] = v.y;
                        As[q * 4 + 2][ml] = v.z;
                        As[q * 4 + 3][ml] = v.w;
                    }
                }
                // ---- stage B ---------------------------------------------------------------
#pragma unroll
                for (int i = 0; i < B_PER_THREAD; ++i) {
                    int item = tid + i * NTHREADS;
                    if (item < B_ITEMS) {
                        int kk = item / (BN / 4), nq = item % (BN / 4);
                        int c = c0 + kk;
                        int n = n_base + nq * 4;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c < a.Cin && n < a.CoutP)
                            v = __ldg(reinterpret_cast<const float4*>(
                                a.w + ((long long)tap * a.Cin + c) * a.CoutP + n));
                        *reinterpret_cast<float4*>(&Bs[kk][nq * 4]) = v;
                    }
                }
                __syncthreads();
                // ---- multiply ----------------------------------------------------------------
#pragma unroll
                for (int kk = 0; kk < BK; ++kk) {
                    float av[TM], bv[TN];
#pragma unroll
                    for (int i = 0; i < TM; i += 4) {
                        float4 t = *reinterpret_cast<const float4*>(&As[kk][tm + i]);
                        av[i] = t.x; av[i + 1] = t.y; av[i + 2] = t.z; av[i + 3] = t.w;
                    }
                    if constexpr (TN % 4 == 0) {
#pragma unroll
                        for (int j = 0; j < TN; j += 4) {
                            float4 t = *reinterpret_cast<const float4*>(&Bs[kk][tn + j]);
                            bv[j] = t.x; bv[j + 1] = t.y; bv[j + 2] = t.z; bv[j + 3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tn + j];
                    }
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
                }
                __syncthreads();
            }
        }
    }

    // ---- epilogue: bias, LeakyReLU, residual, store ----------------------------------------------
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        long long m = m_base + tm + i;
        if (m >= a.M) continue;
        float* yp = a.y + m * a.y_ld;
        const float* rp = a.res ? a.res + m * a.res_ld : nullptr;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n_base + tn + j;
            if (n < a.Cout) {
                float v = acc[i][j] + (a.bias ? __ldg(a.bias + n) : 0.f);
                if (a.lrelu) v = lrelu_f(v);
                if (rp) v += rp[n];
                yp[n] = v;
            }
        }
    }
}

template <int BM, int BN, int TM, int TN>
int launch(const ConvArgs& a, bool vec, cudaStream_t st) {
    dim3 grid((unsigned)cdivll(a.M, BM), (unsigned)cdiv(a.Cout, BN));
    if (vec)
        conv_simt_kernel<BM, BN, TM, TN, true><<<grid, NTHREADS, 0, st>>>(a);
    else
        conv_simt_kernel<BM, BN, TM, TN, false><<<grid, NTHREADS, 0, st>>>(a);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

}  // namespace

extern "C" int pivlfn_conv_simt(const float* x, int x_ld, int N, int H, int W, int Cin,
                                const float* w, const float* bias, float* y, int y_ld, int Cout,
                                int KH, int KW, int stride, int lrelu,
                                const float* res, int res_ld, void* stream) {
    if (!x || !w || !y || N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return PIVLFN_EINVAL;
    if (x_ld < Cin || y_ld < Cout || (res && res_ld < Cout)) return PIVLFN_EINVAL;
    if (KH < 1 || KW < 1 || !(KH & 1) || !(KW & 1) || (stride != 1 && stride != 2)) return PIVLFN_EINVAL;
    ConvArgs a;
    a.x = x; a.x_ld = x_ld; a.N = N; a.H = H; a.W = W; a.Cin = Cin;
    a.w = w; a.CoutP = (Cout + 3) & ~3; a.bias = bias;
    a.y = y; a.y_ld = y_ld; a.Cout = Cout;
    a.KH = KH; a.KW = KW; a.stride = stride; a.lrelu = lrelu;
    a.res = res; a.res_ld = res_ld;
    a.Ho = (H + 2 * (KH / 2) - KH) / stride + 1;
    a.Wo = (W + 2 * (KW / 2) - KW) / stride + 1;
    a.M = (long long)N * a.Ho * a.Wo;
    if (((uintptr_t)w & 15) != 0) return PIVLFN_EINVAL;
    bool vec = ((uintptr_t)x % 16 == 0) && (x_ld % 4 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (Cout <= 8) return launch<256, 8, 4, 2>(a, vec, st);
    if (Cout <= 32) return launch<128, 32, 4, 4>(a, vec, st);
    return launch<128, 64, 8, 4>(a, vec, st);
}
