// Microbenchmark: how fast does the TMA engine of one SM deliver the convolution's activation tiles?
// Tensor x[N=64][256][256][C] fp32 (NHWC, 1 GB at C = 64: far larger than L2); every CTA (one per SM) walks the conv
// kernel's work items (16 x 16 output pixels + halo -> box 32 channels x 18 x 18 pixels = 41,472 bytes, 128B swizzle) and
// keeps DEPTH loads in flight (ring of DEPTH shared-memory slots, one mbarrier each).  Nobody reads the tiles.
//   variant 0: the conv's box {32, 18, 18, 1}          variant 1: two boxes {32, 18, 9, 1} per tile
//   variant 2: no swizzle                              variant 3: L2 promotion 256 B
//   variant 4: C = 32 (pixel pitch 128 B: tile rows contiguous in memory)
//   variant 5: box {32, 16, 16, 1} without halo (aligned, every byte read once)
// Prints cycles per tile, bytes/clk per SM and the aggregate GB/s for DEPTH = 1, 2, 3, 4.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/tma_tile_rate tools/ubench/tma_tile_rate.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int SLOT = 43008;
constexpr int MAXD = 5;

__global__ void __launch_bounds__(32, 1)
bench(const __grid_constant__ CUtensorMap tm, int depth, int nchunk, int tiles_x, int tiles_y, int total, int pieces, int rows_per_piece,
      int bytes_per_tile, int halo, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar[MAXD];
    if (threadIdx.x == 0) {
        for (int i = 0; i < MAXD; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const long long t0 = clock64();
    long long g = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int tx = w % tiles_x, ty = (w / tiles_x) % tiles_y, n = w / (tiles_x * tiles_y);
        for (int c = 0; c < nchunk; ++c, ++g) {
            const int s = (int)(g % depth);
            const long long use = g / depth;                       // how many times slot s was loaded before
            if (use > 0) mbar_wait(&bar[s], (uint32_t)((use - 1) & 1));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(bytes_per_tile) : "memory");
            for (int p = 0; p < pieces; ++p)
                asm volatile(
                    "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                    ::"r"(smem_u32(smem + (size_t)s * SLOT + (size_t)p * (bytes_per_tile / pieces))), "l"(&tm), "r"(smem_u32(&bar[s])),
                      "r"(c * 32), "r"(tx * 16 - halo), "r"(ty * 16 - halo + p * rows_per_piece), "r"(n) : "memory");
        }
    }
    // drain
    for (long long k = (g > depth ? g - depth : 0); k < g; ++k) mbar_wait(&bar[k % depth], (uint32_t)((k / depth) & 1));
    if (blockIdx.x == 0) { out[0] = clock64() - t0; out[1] = g; }
}

// Weight boxes of one ring stage (f16c, Cout = 32, 3 taps x [hi | lo] x 32 rows of 64 bytes = 12,288 bytes, L2-resident):
// mode 0 = one 4-D tensor box per stage (what the conv kernel issues), mode 1 = the same bytes as ONE contiguous bulk copy
__global__ void __launch_bounds__(32, 1)
bench_w(const __grid_constant__ CUtensorMap tm, const uint8_t* img, int mode, int depth, int iters, int bytes, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar[MAXD];
    if (threadIdx.x == 0) {
        for (int i = 0; i < MAXD; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const long long t0 = clock64();
    for (int g = 0; g < iters; ++g) {
        const int s = g % depth, use = g / depth;
        if (use > 0) mbar_wait(&bar[s], (uint32_t)((use - 1) & 1));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(bytes) : "memory");
        const int c = (g / 3) & 1, t = (g % 3) * 3;
        if (mode == 0)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(smem_u32(smem + (size_t)s * 16384)), "l"(&tm), "r"(smem_u32(&bar[s])), "r"(c * 32), "r"(0), "r"(0), "r"(t) : "memory");
        else
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem + (size_t)s * 16384)), "l"(img + (size_t)((c * 3 + g % 3) * bytes)), "r"(bytes), "r"(smem_u32(&bar[s])) : "memory");
    }
    for (int k = (iters > depth ? iters - depth : 0); k < iters; ++k) mbar_wait(&bar[k % depth], (uint32_t)((k / depth) & 1));
    if (blockIdx.x == 0) { out[0] = clock64() - t0; out[1] = iters; }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no encode\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int N = 64, H = 256, W = 256;
    float* x;
    cudaMalloc(&x, (size_t)N * H * W * 64 * 4);
    cudaMemset(x, 0, (size_t)N * H * W * 64 * 4);
    long long* out;
    cudaMalloc(&out, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXD * SLOT + 1024);
    cudaFuncSetAttribute(bench_w, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXD * SLOT + 1024);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int variant = 0; variant < 6; ++variant) {
        const int C = variant == 4 ? 32 : 64;
        const int halo = variant == 5 ? 0 : 1;
        const int side = 16 + 2 * halo;
        const int pieces = variant == 1 ? 2 : 1;
        CUtensorMap tm;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)side, (cuuint32_t)(side / pieces), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         variant == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                         variant == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("variant %d: encode failed %d\n", variant, (int)r); continue; }
        const int bytes = side * side * 128;
        const int nchunk = C / 32, tiles = (W / 16) * (H / 16) * N;
        for (int depth = 1; depth <= 4; ++depth) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            bench<<<sms, 32, MAXD * SLOT + 1024>>>(tm, depth, nchunk, W / 16, H / 16, tiles, pieces, side / pieces, bytes, halo, out);
            cudaEventRecord(e0);
            bench<<<sms, 32, MAXD * SLOT + 1024>>>(tm, depth, nchunk, W / 16, H / 16, tiles, pieces, side / pieces, bytes, halo, out);
            cudaEventRecord(e1);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("variant %d depth %d: %s\n", variant, depth, cudaGetErrorString(cudaGetLastError())); return 1; }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            long long h[2];
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("variant %d depth %d: %6.0f cycles/tile (%d B)  %5.1f B/clk/SM  %6.0f GB/s fetched  (%.3f ms)\n", variant, depth,
                   (double)h[0] / (double)h[1], bytes, (double)bytes * h[1] / h[0], (double)bytes * tiles * nchunk / ms / 1e6, ms);
        }
    }
    {
        // weights: [parts = 2][CoutP = 32][ntaps = 9][CinP = 64] fp16
        const int CinP = 64, CoutP = 32, ntaps = 9, parts = 2, tps = 3;
        uint8_t* w;
        cudaMalloc(&w, 1 << 20);
        cudaMemset(w, 0, 1 << 20);
        CUtensorMap tm;
        cuuint64_t dims[4] = {(cuuint64_t)CinP, (cuuint64_t)CoutP, (cuuint64_t)parts, (cuuint64_t)ntaps};
        cuuint64_t strides[3] = {(cuuint64_t)ntaps * CinP * 2, (cuuint64_t)CoutP * ntaps * CinP * 2, (cuuint64_t)CinP * 2};
        cuuint32_t box[4] = {32, (cuuint32_t)CoutP, (cuuint32_t)parts, (cuuint32_t)tps};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("weight map: encode failed %d\n", (int)r); return 1; }
        const int bytes = 32 * 2 * CoutP * parts * tps;
        for (int mode = 0; mode < 2; ++mode)
            for (int depth = 1; depth <= 4; ++depth) {
                bench_w<<<sms, 32, MAXD * SLOT + 1024>>>(tm, w, mode, depth, 6000, bytes, out);
                bench_w<<<sms, 32, MAXD * SLOT + 1024>>>(tm, w, mode, depth, 6000, bytes, out);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("weights mode %d: %s\n", mode, cudaGetErrorString(cudaGetLastError())); return 1; }
                long long h[2];
                cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                printf("weight stage (%d B, L2-resident) %s depth %d: %6.0f cycles/stage  %5.1f B/clk/SM\n", bytes,
                       mode == 0 ? "4-D tensor box, 192 rows of 64 B" : "one contiguous bulk copy      ", depth,
                       (double)h[0] / h[1], (double)bytes * h[1] / h[0]);
            }
    }
    return 0;
}
