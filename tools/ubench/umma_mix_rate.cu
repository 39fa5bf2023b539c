// Microbenchmark: one f16 MMA (K = 16) + one fp8 MMA (kind::f8f6f4, e5m2 x e5m2, K = 32) per 16-channel K step against the three f16
// MMAs of the 3-product split, with the operand layouts of conv_p16: A = windows into a halo tile of 128-byte rows (SWIZZLE_128B,
// 8-row groups one halo row = pitch * 128 B apart), B = 64-byte rows (SWIZZLE_64B, groups 512 B apart), M = 128, one CTA per SM.
//   scheme 0: f16 x3 per K step (a_hi*W_hi, a_hi*W_lo, a_lo*W_hi)          -- what mode 5 issues
//   scheme 1: f16 + fp8 per K step (a_hi*W_hi, [lo8|hi8]*[W_hi8;W_lo8])    -- the candidate
//   scheme 2: fp8 only, 2 per K step (rate check of kind::f8f6f4 alone)
//   issuers 1 / 2: one thread, or two threads with their own stacked tile and accumulator (the conv kernel's arrangement)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/umma_mix_rate tools/ubench/umma_mix_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int F8>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    if (F8)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n\t}"
                     ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                     ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}

template <int SCHEME, int ISS>
__global__ void __launch_bounds__(96, 1) bench(int N, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(ISS));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // pseudo-random operand bits with small exponents: valid (finite) as fp16 pairs and as e5m2 bytes
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x83838383u) | 0x3C3C3C3Cu;
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
    if ((warp == 1 || (warp == 2 && ISS == 2)) && elect_one()) {
        const int issuer = warp - 1;
        const uint32_t pitch = 10;
        const uint32_t lbo_bits = 1u << 16;
        const uint32_t hiA = (uint32_t)((((uint64_t)((pitch * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61)) >> 32);
        const uint32_t hiB = (uint32_t)((((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61)) >> 32);
        const uint32_t idesc16 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);                              // F16 x F16 -> F32
        const uint32_t idesc8 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);     // E5M2 x E5M2 -> F32
        const uint32_t tile16 = 16 * pitch * 8;
        const uint32_t A00 = ((smem_u32(smem) >> 4) | lbo_bits) + (uint32_t)issuer * tile16;       // halo tile 34 x 10 x 128 B = 43.5 KB
        const uint32_t b0 = (smem_u32(smem + 64 * 1024) >> 4) | lbo_bits;                         // weight ring
        const uint32_t part16 = (uint32_t)(N * 64) >> 4;
        const uint32_t t_main = tmem + (uint32_t)issuer * 256u;
        const uint32_t row_step = (pitch - 3) * 8;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            uint32_t A0 = A00;
            int kx = 0;
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
                const uint32_t b = b0 + (uint32_t)(tap & 3) * 3 * part16;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const uint32_t A = A0 + 4 * kk, bb = b + 2 * kk;
                    if (SCHEME == 0) {
                        umma<0>(t_main, A, hiA, bb, hiB, idesc16, 1);
                        umma<0>(t_main, A, hiA, bb + part16, hiB, idesc16, 1);
                        umma<0>(t_main, A + 2, hiA, bb + 2 * part16, hiB, idesc16, 1);
                    } else if (SCHEME == 1) {
                        umma<0>(t_main, A, hiA, bb, hiB, idesc16, 1);
                        umma<1>(t_main, A + 2, hiA, bb + part16, hiB, idesc8, 1);
                    } else {
                        umma<1>(t_main, A, hiA, bb, hiB, idesc8, 1);
                        umma<1>(t_main, A + 2, hiA, bb + part16, hiB, idesc8, 1);
                    }
                }
                A0 += 8;
                if (++kx == 3) { kx = 0; A0 += row_step; }
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

template <int S, int I>
static void run(int N, long long* d) {
    const int iters = 2000, smem = 200 * 1024;
    cudaFuncSetAttribute(bench<S, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    bench<S, I><<<148, 96, smem>>>(N, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    const double per_step = (double)c / (iters * 18.0);            // 9 taps x 2 K steps of 16 channels
    // useful flops of one K step of one tile: 2 * 128 * N * 16; I tiles in flight
    printf("scheme=%d issuers=%d N=%3d : %.1f cycles per 16-channel K step -> %.0f useful MAC/clk/SM (%s)\n", S, I, N, per_step,
           128.0 * N * 16 * I / per_step, cudaGetErrorString(e));
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    for (int N : {32, 64, 96, 128}) {
        run<0, 1>(N, d); run<1, 1>(N, d); run<2, 1>(N, d);
        run<0, 2>(N, d); run<1, 2>(N, d); run<2, 2>(N, d);
    }
    return 0;
}
