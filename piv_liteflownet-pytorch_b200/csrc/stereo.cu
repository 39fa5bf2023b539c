// Stereo-PIV post-processing of two camera flows, elementwise on the GPU so that the
// stereo_run.py workload needs no device -> host round trip between estimate() and the 3-component result:
//
//   nl_trans   stereo/dewarp.py:255-270: 24-coefficient rational-quadratic map, applied by stereo_run._stereo_cal (:153-163)
//              to the two FLOW components (x = u, y = v), then "* calibrate * fps" when a calibration is given;
//   willert    stereo/vel3d.py:4-24: Willert (1997) recombination of the left / right camera flows into (U, V, W).
//
// Arithmetic follows the reference operation by operation, in float32 with every product and sum rounded separately
// (numpy evaluates the expressions term by term, left to right).  Scalars: under the reference's pinned numpy 1.17
// (requirements.txt:9, value-based casting) a Python float or an np.float64 scalar that meets a float32 array is rounded to
// float32 and the operation runs in float32; scalar-with-scalar arithmetic (tan(theta0) - tan(theta1)) stays float64 and is
// rounded when it meets an array.  The results are therefore bit-identical to the reference's numpy path in its own
// environment (numpy >= 2 would promote willert to float64: same values to ~1e-7 relative).  Fused-multiply-add
// contraction is suppressed with the round-to-nearest intrinsics.
#include "common.cuh"

namespace {

struct NlCoef { float a[24]; };

__device__ __forceinline__ float nl_poly(const float* a, float x, float y) {
    // ((((a0*x + a1*y) + a2) + a3*x^2) + a4*y^2) + (a5*x)*y          (stereo/dewarp.py:263-264)
    float t = __fadd_rn(__fmul_rn(a[0], x), __fmul_rn(a[1], y));
    t = __fadd_rn(t, a[2]);
    t = __fadd_rn(t, __fmul_rn(a[3], __fmul_rn(x, x)));
    t = __fadd_rn(t, __fmul_rn(a[4], __fmul_rn(y, y)));
    t = __fadd_rn(t, __fmul_rn(__fmul_rn(a[5], x), y));
    return t;
}

__device__ __forceinline__ void nl_trans_dev(const NlCoef& A, float x, float y, float& nx, float& ny) {
    nx = __fdiv_rn(nl_poly(A.a, x, y), nl_poly(A.a + 6, x, y));
    ny = __fdiv_rn(nl_poly(A.a + 12, x, y), nl_poly(A.a + 18, x, y));
}

__global__ void nl_trans_kernel(const float* __restrict__ x, const float* __restrict__ y, NlCoef A,
                                float* __restrict__ nx, float* __restrict__ ny, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        nl_trans_dev(A, __ldg(x + i), __ldg(y + i), nx[i], ny[i]);
}

struct StereoArgs {
    NlCoef AL, AR;
    int use_map;                 // 0: the flows are combined as they are (willert only)
    int use_calib;
    float calib, fps;            // flow * calib * fps, two float32 products (stereo_run.py:160-161)
    float tan_t0, tan_t1;        // f32(tan(theta0)), f32(tan(theta1))
    float dt, db;                // f32(tan(theta0) - tan(theta1)), f32(tan(beta1) - tan(beta0)) (differences taken in float64)
};

// fl, fr: [B, 2, H, W] (the layout estimate(..., tensor=True) returns); out: [B, H, W, 3] (the 3-band .flo layout)
__global__ void stereo_2d3c_kernel(const float* __restrict__ fl, const float* __restrict__ fr, StereoArgs a,
                                   float* __restrict__ out, int B, long long HW) {
    const long long total = (long long)B * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / HW, p = i - b * HW;
        float u0 = __ldg(fl + (2 * b) * HW + p), v0 = __ldg(fl + (2 * b + 1) * HW + p);
        float u1 = __ldg(fr + (2 * b) * HW + p), v1 = __ldg(fr + (2 * b + 1) * HW + p);
        if (a.use_map) {
            float nx, ny;
            nl_trans_dev(a.AL, u0, v0, nx, ny); u0 = nx; v0 = ny;
            nl_trans_dev(a.AR, u1, v1, nx, ny); u1 = nx; v1 = ny;
            if (a.use_calib) {
                u0 = __fmul_rn(__fmul_rn(u0, a.calib), a.fps); v0 = __fmul_rn(__fmul_rn(v0, a.calib), a.fps);
                u1 = __fmul_rn(__fmul_rn(u1, a.calib), a.fps); v1 = __fmul_rn(__fmul_rn(v1, a.calib), a.fps);
            }
        }
        // stereo/vel3d.py:19-22
        const float u3 = __fdiv_rn(__fsub_rn(__fmul_rn(u1, a.tan_t0), __fmul_rn(u0, a.tan_t1)), a.dt);
        const float du = __fsub_rn(u1, u0);
        const float v3 = __fadd_rn(__fdiv_rn(__fadd_rn(v0, v1), 2.f),
                                   __fdiv_rn(__fdiv_rn(__fmul_rn(du, a.db), a.dt), 2.f));
        const float w3 = __fdiv_rn(du, a.dt);
        out[3 * i] = u3;
        out[3 * i + 1] = v3;
        out[3 * i + 2] = w3;
    }
}

inline int grid_1d(long long total) {
    long long g = (total + 255) / 256;
    return (int)(g < 148LL * 16 ? (g > 0 ? g : 1) : 148LL * 16);
}

}  // namespace

extern "C" int pivlfn_nl_trans(const float* x, const float* y, const float* A24, float* new_x, float* new_y, long long n,
                               void* stream) {
    if (!x || !y || !A24 || !new_x || !new_y || n <= 0) return PIVLFN_EINVAL;
    NlCoef A;
    for (int i = 0; i < 24; ++i) A.a[i] = A24[i];
    nl_trans_kernel<<<grid_1d(n), 256, 0, (cudaStream_t)stream>>>(x, y, A, new_x, new_y, n);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}

extern "C" int pivlfn_stereo_2d3c(const float* flow_left, const float* flow_right, const float* A_left, const float* A_right,
                                  int use_calib, float calib, float fps, double tan_theta0, double tan_theta1,
                                  double tan_beta0, double tan_beta1, float* out, int B, int H, int W, void* stream) {
    if (!flow_left || !flow_right || !out || B <= 0 || H <= 0 || W <= 0) return PIVLFN_EINVAL;
    if ((A_left == nullptr) != (A_right == nullptr)) return PIVLFN_EINVAL;
    if (tan_theta0 == tan_theta1) return PIVLFN_EINVAL;
    StereoArgs a;
    a.use_map = A_left != nullptr;
    for (int i = 0; i < 24; ++i) { a.AL.a[i] = a.use_map ? A_left[i] : 0.f; a.AR.a[i] = a.use_map ? A_right[i] : 0.f; }
    a.use_calib = use_calib; a.calib = calib; a.fps = fps;
    a.tan_t0 = (float)tan_theta0; a.tan_t1 = (float)tan_theta1;
    a.dt = (float)(tan_theta0 - tan_theta1); a.db = (float)(tan_beta1 - tan_beta0);
    const long long HW = (long long)H * W;
    stereo_2d3c_kernel<<<grid_1d((long long)B * HW), 256, 0, (cudaStream_t)stream>>>(flow_left, flow_right, a, out, B, HW);
    PIVLFN_LAUNCHED();
    return pivlfn_last_error();
}
