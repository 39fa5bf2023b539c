// Microbenchmark: issue rate of tcgen05.mma (kind::tf32 / kind::f16) from shared memory, M = 128.
// One CTA per SM, one issuing thread; operands are whatever is in shared memory (values irrelevant).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo = 1024) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
template <int KIND>  // 0 = tf32, 1 = f16 (bf16 inputs)
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: one accumulator; 1: alternate two accumulators; 2: A fixed, B varies (different smem tiles)
// mode 3: A with SBO = 1280 (halo pitch 10); 4: SBO = 1280 and start + 128 B; 5: SBO = 1024, start + 128 B;
// mode 6: like 4 but a different window start for every MMA (tap shifts)
template <int KIND>
__global__ void __launch_bounds__(64, 1) bench(int N, int iters, int mode, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (warp == 1 && elect_one()) {
        const uint32_t fmt = KIND == 0 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t sA = smem_u32(smem), sB = sA + 24576;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t d = tmem + ((mode == 1 && (k & 1)) ? 256u : 0u);
                uint32_t sb = sB + ((mode == 2) ? (uint32_t)(k * 8192 % 32768) : 0u);
                uint32_t sa = sA + ((mode == 4 || mode == 5) ? 128u : 0u) + (mode == 6 ? (uint32_t)(((it + k) % 3) * 128 + ((it / 3) % 3) * 1280) : 0u);
                uint32_t sbo = (mode == 3 || mode == 4 || mode == 6) ? 1280u : 1024u;
                umma<KIND>(d, make_desc(sa, sbo) + k * 2, make_desc(sb) + k * 2, idesc, 1);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        } while (!done);
        long long t1 = clock64();
        if (blockIdx.x == 0) *out = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 2000;
    for (int kind = 0; kind < 1; ++kind)
        for (int mode = 0; mode < 7; ++mode)
            for (int N : {16, 64, 128}) {
                for (int grid : {148}) {
                    if (kind == 0) bench<0><<<grid, 64, 64 * 1024>>>(N, iters, mode, d);
                    else bench<1><<<grid, 64, 64 * 1024>>>(N, iters, mode, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long c = 0;
                    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                    printf("kind=%s mode=%d N=%3d grid=%3d : %.1f cycles/MMA (%s)\n", kind ? "f16 " : "tf32", mode, N, grid,
                           (double)c / (iters * 4), cudaGetErrorString(e));
                }
            }
    return 0;
}
