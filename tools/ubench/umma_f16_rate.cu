// Microbenchmark: cost of tcgen05.mma.kind::f16 (fp16 operands, K = 16) with the operand layouts of the f16c convolution:
// 64-byte rows, SWIZZLE_64B, A = windows into a halo tile (8-row core groups `pitch * 64` bytes apart), B = dense weight
// tile (groups 512 bytes apart).  One CTA per SM, ONE issuing thread, M = 128.
//   mode 0: the same A window for every MMA           mode 1: a different window (tap shift) for every MMA
//   mode 2: like 1, two MMAs per window (a_hi*[w_hi|w_lo] then a_lo*w_hi of the same tap: the conv's mode-4 pattern)
//   mode 3: like 1 with pitch 8 (dense, SBO 512)      (modes 0-2: pitch 10, SBO 640)
//   mode 4: mode 2 + a tcgen05.commit to an mbarrier after every tap (the conv frees one weight stage per tap)
//   mode 7: mode 2 with both K steps of a 32-channel chunk (operand start + 0 / + 32 bytes)
//   mode 8: mode 7 with a different weight tile for every tap (stage ring of 4 tiles, 8 KB apart)
//   mode 9: mode 8 with the window shift and the weight-stage offset read from shared memory per tap (the operands then
//           sit in VECTOR registers and need an R2UR per MMA, like in the convolution kernel's issue loop)
//   mode 5: mode 2 from TWO issuing threads (warps 1 and 2, own accumulators), mode 6: mode 5 + the per-tap commits
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/umma_f16_rate tools/ubench/umma_f16_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// K-major, SWIZZLE_64B (layout type 4), 8-row group stride sbo bytes
__device__ __forceinline__ uint64_t make_desc64(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

#ifndef RANDOM_DATA
#define RANDOM_DATA 0
#endif
template <int MODE>
__global__ void __launch_bounds__(96, 1) bench(int N, int iters, long long* out) {
    constexpr int mode = MODE;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, bar2[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile uint32_t tab[16];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < 9) tab[threadIdx.x] = (uint32_t)(((threadIdx.x / 3) * 10 + threadIdx.x % 3) * 4);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(MODE >= 5 ? 2 : 1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // operand values: constant 1.0h, or (RANDOM_DATA) pseudo-random fp16 in about [-2, 2] -- the tensor pipe's power draw, and
    // with it any power management, depends on how the operand bits toggle
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) {
        uint32_t v = 0x3C003C00u;
        if (RANDOM_DATA) {
            uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            v = (h & 0x83FF83FFu) | 0x3C003C00u | ((h >> 3) & 0x04000400u);      // sign + mantissa random, exponent 15 or 16
        }
        reinterpret_cast<uint32_t*>(smem)[i] = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if ((warp == 1 || (warp == 2 && MODE >= 5)) && elect_one()) {
        const uint32_t tmem_i = tmem + (warp == 2 ? 256u : 0u);
        const uint32_t a_off = warp == 2 ? (10u * 16u * 64u) >> 4 : 0u;      // second issuer: the stacked tile 16 rows further
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);        // F16 x F16 -> F32
        const uint32_t idesc_half = (1u << 4) | ((uint32_t)((N / 2) >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t pitch = MODE == 3 ? 8u : 10u;
        const uint32_t sbo = pitch * 64u;
        const uint32_t sA = smem_u32(smem), sB = sA + 49152;                // halo tile 34 x 10 x 64 B = 21.8 KB (x2 hi / lo)
        long long t0 = clock64();
        // descriptors as (constant high word, low word) pairs and compile-time tap shifts: the issuing thread does one add
        // per MMA, like the convolution kernel
        const uint64_t hiA = make_desc64(0, sbo) & 0xFFFFFFFF00000000ull, hiB = make_desc64(0, 512) & 0xFFFFFFFF00000000ull;
        const uint32_t loA = (uint32_t)make_desc64(sA, sbo), loAlo = (uint32_t)make_desc64(sA + 24576, sbo), loB = (uint32_t)make_desc64(sB, 512);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const uint32_t shift16 = mode == 0 ? 0u : (uint32_t)(((tap / 3) * (MODE == 3 ? 8 : 10) + tap % 3) * 4);   // 64 B >> 4
                const uint64_t da = hiA | (uint64_t)(loA + shift16);
                const uint64_t db = hiB | (uint64_t)loB;
                if (mode == 9) {
                    const uint32_t sh = tab[tap];                                           // vector register
                    const uint32_t bo = (tab[(tap + it) % 9] & 3u) * 512u;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        umma_f16(tmem_i, (hiA | (uint64_t)(loA + sh)) + 2 * kk, (hiB | (uint64_t)(loB + bo)) + 2 * kk, idesc, 1);
                        umma_f16(tmem_i + N / 2, (hiA | (uint64_t)(loAlo + sh)) + 2 * kk, (hiB | (uint64_t)(loB + bo)) + 2 * kk, idesc_half, 1);
                    }
                } else if (mode == 7 || mode == 8) {
                    const uint64_t dbt = db + (mode == 8 ? (uint64_t)((tap & 3) * 512) : 0ull);
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        umma_f16(tmem_i, da + 2 * kk, dbt + 2 * kk, idesc, 1);
                        umma_f16(tmem_i + N / 2, (hiA | (uint64_t)(loAlo + shift16)) + 2 * kk, dbt + 2 * kk, idesc_half, 1);
                    }
                } else if (mode == 2 || mode >= 4) {
                    umma_f16(tmem_i, da + a_off, db, idesc, 1);                           // a_hi * [w_hi | w_lo]: N columns
                    umma_f16(tmem_i + N / 2, (hiA | (uint64_t)(loAlo + shift16)) + a_off, db, idesc_half, 1);   // a_lo * w_hi: N/2
                    if (mode == 4 || mode == 6)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[warp - 1])) : "memory");
                } else {
                    umma_f16(tmem, da, db, idesc, 1);
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        } while (!done);
        long long t1 = clock64();
        if (blockIdx.x == 0 && warp == 1) *out = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(bench<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(bench<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(bench<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(bench<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(bench<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(bench<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(bench<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    const int iters = 2000;
    for (int mode = 0; mode < 10; ++mode)
        for (int N : {32, 64, 128, 256}) {
            if (mode == 0) bench<0><<<148, 96, 96 * 1024>>>(N, iters, d);
            else if (mode == 1) bench<1><<<148, 96, 96 * 1024>>>(N, iters, d);
            else if (mode == 2) bench<2><<<148, 96, 96 * 1024>>>(N, iters, d);
            else if (mode == 3) bench<3><<<148, 96, 96 * 1024>>>(N, iters, d);
            else if (mode == 4) bench<4><<<148, 96, 96 * 1024>>>(N, iters, d);
            else if (mode == 5) bench<5><<<148, 96, 96 * 1024>>>(N, iters, d);
            else if (mode == 6) bench<6><<<148, 96, 96 * 1024>>>(N, iters, d);
            else if (mode == 7) bench<7><<<148, 96, 96 * 1024>>>(N, iters, d);
            else if (mode == 8) bench<8><<<148, 96, 96 * 1024>>>(N, iters, d);
            else bench<9><<<148, 96, 96 * 1024>>>(N, iters, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long c = 0;
            cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            const int per = mode >= 7 ? 36 : ((mode == 2 || mode >= 4) ? 18 : 9);
            printf("kind=f16 mode=%d N=%3d : %.1f cycles per MMA (%.1f per tap) (%s)\n", mode, N, (double)c / (iters * per),
                   (double)c / (iters * 9), cudaGetErrorString(e));
        }
    return 0;
}
