// P16: the activation format of the fp16-split ("f16c") pipeline.
//
// An fp32 activation x is kept in HBM as the PAIR the tensor cores consume,
//     hi = f16(x),   lo = f16((x - f16(x)) * 2^11),     x = hi + 2^-11 * lo   (22 significant bits, |x| < 65504),
// 4 bytes per element like fp32.  Channels are grouped by 16: one group of one pixel is 64 contiguous bytes,
//     [ hi(c0) .. hi(c15) | lo(c0) .. lo(c15) ]  =  16 words (8 words of packed f16x2 hi, then 8 of lo),
// so a View (pointer, channel words, pixel pitch in words) addresses a P16 tensor exactly like an fp32 NHWC tensor of
// 16 * ngroups channels, channel slices start at multiples of 16, and a 32-channel chunk of a pixel is one 128-byte row
// [hi0 | lo0 | hi1 | lo1]: TMA drops it into shared memory as an MMA-ready K-major tile (128B swizzle) whose four 32-byte
// K = 16 steps are hi(0..15), lo(0..15), hi(16..31), lo(16..31).  No kernel ever splits operands in shared memory.
// Pad channels of the last group hold hi = lo = 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace p16 {

constexpr float LO_SCALE = 2048.f;            // 2^11
constexpr float LO_INV = 1.f / 2048.f;

// two fp32 -> packed f16x2 (round to nearest even), low half = first value
__device__ __forceinline__ uint32_t pack_hi(float a, float b) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ void unpack2(uint32_t h, float& a, float& b) {
    asm("{\n\t.reg .b16 l, u;\n\tmov.b32 {l, u}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, u;\n\t}" : "=f"(a), "=f"(b) : "r"(h));
}
// the scaled residuals of (a, b) against their packed hi halves
__device__ __forceinline__ uint32_t pack_lo(float a, float b, uint32_t h) {
    float ha, hb;
    unpack2(h, ha, hb);
    return pack_hi((a - ha) * LO_SCALE, (b - hb) * LO_SCALE);
}
// fp32 value of a (hi, lo) word pair: two channels
__device__ __forceinline__ void decode2(uint32_t h, uint32_t l, float& a, float& b) {
    float ha, hb, la, lb;
    unpack2(h, ha, hb);
    unpack2(l, la, lb);
    a = fmaf(la, LO_INV, ha);
    b = fmaf(lb, LO_INV, hb);
}
// 8 channels (half a group): hi and lo as one 16-byte vector each
__device__ __forceinline__ void decode8(const uint4& h, const uint4& l, float* v) {
    decode2(h.x, l.x, v[0], v[1]);
    decode2(h.y, l.y, v[2], v[3]);
    decode2(h.z, l.z, v[4], v[5]);
    decode2(h.w, l.w, v[6], v[7]);
}
__device__ __forceinline__ void encode8(const float* v, uint4& h, uint4& l) {
    h.x = pack_hi(v[0], v[1]); h.y = pack_hi(v[2], v[3]); h.z = pack_hi(v[4], v[5]); h.w = pack_hi(v[6], v[7]);
    l.x = pack_lo(v[0], v[1], h.x); l.y = pack_lo(v[2], v[3], h.y); l.z = pack_lo(v[4], v[5], h.z); l.w = pack_lo(v[6], v[7], h.w);
}
// exponent field all ones in either half of a packed f16x2 word <=> inf or NaN: bit 15 / 31 of the result
__device__ __forceinline__ uint32_t nonfinite_bits(uint32_t h) { return (h & 0x7C007C00u) + 0x04000400u; }
__device__ __forceinline__ bool any_nonfinite(uint32_t acc) { return (acc & 0x80008000u) != 0u; }

// byte offset inside a pixel row of the hi vector of 8-channel unit u (u = channel / 8); the lo vector is 32 bytes further
__device__ __forceinline__ int unit_off_bytes(int u) { return (u >> 1) * 64 + (u & 1) * 16; }

}  // namespace p16
