// Microbenchmark (results: profiles/r1_ubench_store_rate.txt): how fast can the epilogue warps of ONE SM store an NHWC fp32 output,
// by store pattern?  The convolution's epilogue delivers ~10 B/clk of output per SM; about half of
// that time is the stores.  Here 8 warps per CTA (one CTA per SM, like the epilogue's two warp groups) write register
// data only -- no TMEM, no math -- to an output of [pixels][C] floats, each warp 32 consecutive pixels x 16 or 32
// channels per step, walking the tensor like the conv kernel's work items.
//   pattern 0: lane = pixel, four float4 stores at channel offsets 0/4/8/12      (32 lines x 16 B per instruction)
//   pattern 1: quad-transposed, lane quad = 4 pixels, one 64-byte run per quad    (8 x 64 B per instruction)   [shipped]
//   pattern 2: lane = pixel, two st.global.v8.f32 (32 B per lane)                 (32 x 32 B per instruction)
//   pattern 3: octet-transposed over 32 channels: 8 lanes write one pixel's 128 B (4 full lines per instruction)
// for C = 32 / 64 / 128 and for 148 or 8 resident CTAs (per-SM limit vs HBM limit).  Prints B/clk per SM and GB/s.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/store_rate tools/ubench/store_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void st_v8(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e),
                 "f"(f), "f"(g), "f"(h) : "memory");
}

template <int PATTERN>
__global__ void __launch_bounds__(256, 1) bench(float* y, int C, long long npix, long long* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int groups = C / (PATTERN == 3 ? 32 : 16);                 // channel groups per pixel block
    // work unit = (32-pixel block, channel group); the 8 warps of a CTA take consecutive units, CTAs stride over the tensor
    const long long units = (npix / 32) * groups;
    const float4 val = make_float4((float)lane, 1.f, 2.f, 3.f);
    const long long t0 = clock64();
    for (long long u = (long long)blockIdx.x * 8 + warp; u < units; u += (long long)gridDim.x * 8) {
        const long long blk = u / groups;
        const int g = (int)(u % groups);
        float* base = y + blk * 32 * C;
        if (PATTERN == 0) {
            float* p = base + (long long)lane * C + g * 16;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(p + 4 * j) = val;
        } else if (PATTERN == 1) {
            const int lq = lane & 3;
            float* p = base + (long long)(lane - lq) * C + g * 16 + lq * 4;
#pragma unroll
            for (int r = 0; r < 4; ++r) *reinterpret_cast<float4*>(p + (long long)r * C) = val;
        } else if (PATTERN == 2) {
            float* p = base + (long long)lane * C + g * 16;
            st_v8(p, val.x, val.y, val.z, val.w, val.x, val.y, val.z, val.w);
            st_v8(p + 8, val.x, val.y, val.z, val.w, val.x, val.y, val.z, val.w);
        } else {
            const int lo = lane & 7;
            float* p = base + (long long)(lane - lo) * C + g * 32 + lo * 4;
#pragma unroll
            for (int r = 0; r < 8; ++r) *reinterpret_cast<float4*>(p + (long long)r * C) = val;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

template <int PATTERN>
void run(float* y, int C, long long npix, int grid, long long* out, const char* name) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<PATTERN><<<grid, 256>>>(y, C, npix, out);
    cudaEventRecord(e0);
    bench<PATTERN><<<grid, 256>>>(y, C, npix, out);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cyc;
    cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)npix * C * 4;
    printf("C=%3d grid=%3d %-34s %6.1f B/clk/SM  %7.0f GB/s  (%.3f ms)\n", C, grid, name, bytes / grid / (double)cyc, bytes / ms / 1e6, ms);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const long long bytes_max = 2LL << 30;
    float* y;
    cudaMalloc(&y, bytes_max);
    long long* out;
    cudaMalloc(&out, 8);
    const int Cs[3] = {32, 64, 128};
    for (int ci = 0; ci < 3; ++ci) {
        const int C = Cs[ci];
        for (int gi = 0; gi < 2; ++gi) {
            const int grid = gi == 0 ? sms : 8;
            // the same bytes per SM in both cases, so that the 8-CTA run is not 18x longer
            const long long npix = (gi == 0 ? bytes_max : bytes_max * 8 / sms) / (C * 4) / 32 * 32;
            run<0>(y, C, npix, grid, out, "lane = pixel, 4 x float4");
            run<1>(y, C, npix, grid, out, "quad-transposed, 64 B runs");
            run<2>(y, C, npix, grid, out, "lane = pixel, 2 x st.v8");
            run<3>(y, C, npix, grid, out, "octet-transposed, 128 B lines");
        }
    }
    return 0;
}
