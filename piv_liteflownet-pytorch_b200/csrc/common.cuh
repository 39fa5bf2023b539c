// Shared helpers for the pivlfn sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pivlfn.h"

#define PIVLFN_LRELU_SLOPE 0.1f

extern long long g_pivlfn_launches;   // defined in misc.cu

#define PIVLFN_LAUNCHED() (g_pivlfn_launches++)

static inline int pivlfn_last_error() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PIVLFN_OK : (int)e;
}

// Per-device one-time opt-in to > 48 KB of dynamic shared memory.  The attribute belongs to the (function, device) pair, so
// the "done" state is a bit per device ordinal: a process that drives several GPUs configures each of them.
template <typename K>
static inline cudaError_t pivlfn_optin_smem(K kern, int bytes, unsigned long long& done_mask) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && ((done_mask >> dev) & 1ull)) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev >= 0 && dev < 64) done_mask |= 1ull << dev;
    return e;
}

// SM count of the current device (cached per device ordinal)
static inline int pivlfn_num_sms() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cache[dev] = n;
    return n;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ float lrelu_f(float v) { return v >= 0.f ? v : PIVLFN_LRELU_SLOPE * v; }

// Bilinear sampling set-up shared by every backwarp consumer (src/models.py:20-35):
// sample position (sx, sy) in pixels of an H x W image, taps outside contribute zero.
struct BilinearTaps {
    int x0, y0;          // top-left tap
    float w00, w01, w10, w11;  // weights of (y0,x0), (y0,x0+1), (y0+1,x0), (y0+1,x0+1), zeroed when outside
};

__device__ __forceinline__ BilinearTaps make_taps(float sx, float sy, int H, int W) {
    BilinearTaps t;
    float fx = floorf(sx), fy = floorf(sy);
    float ax = sx - fx, ay = sy - fy;
    // clamp before the int conversion so that huge / non-finite flows cannot overflow
    fx = fminf(fmaxf(fx, -2.f), (float)W);
    fy = fminf(fmaxf(fy, -2.f), (float)H);
    t.x0 = (int)fx;
    t.y0 = (int)fy;
    bool vx0 = t.x0 >= 0 && t.x0 < W, vx1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
    bool vy0 = t.y0 >= 0 && t.y0 < H, vy1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
    t.w00 = (vx0 && vy0) ? (1.f - ax) * (1.f - ay) : 0.f;
    t.w01 = (vx1 && vy0) ? ax * (1.f - ay) : 0.f;
    t.w10 = (vx0 && vy1) ? (1.f - ax) * ay : 0.f;
    t.w11 = (vx1 && vy1) ? ax * ay : 0.f;
    return t;
}
