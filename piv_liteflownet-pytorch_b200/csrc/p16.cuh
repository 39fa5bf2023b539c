// P16: the activation format of the split-operand ("f16c") pipeline.
//
// An fp32 activation x is kept in HBM as the operands the tensor cores consume, 4 bytes per element like fp32:
//     hi  = f16(x)                         2 bytes   the main product's operand (kind::f16, K = 16)
//     lo8 = e5m2((x - hi) * 2^11)          1 byte    } the two correction products a_lo * W_hi + a_hi * W_lo run as ONE fp8 MMA
//     hi8 = e5m2(x)                        1 byte    } (kind::f8f6f4, K = 32 = [16 x lo8 | 16 x hi8]) at twice the f16 rate;
//                                                       the weight side of that MMA is e4m3 (per-layer scale: pivlfn.model._pack_f8)
// x = hi + 2^-11 * lo8 up to 2^-14 |x| (the correction terms carry 3 significant bits: tools/sim_precision.py measures the
// flow error of the whole network at 3e-4 px max against 7e-3 px for single-pass TF32; the 3-product fp16 split it replaces
// spent three full-rate MMAs per product for 1e-6 px).  e5m2 has the exponent range of fp16, so the range check stays the
// fp16 one (|x| < 65504; hi8 saturates at 57344, which only touches a correction term).
// Channels are grouped by 16: one group of one pixel is 64 contiguous bytes,
//     [ hi(c0) .. hi(c15) | lo8(c0) .. lo8(c15) | hi8(c0) .. hi8(c15) ]  =  32 + 16 + 16 bytes = 16 words,
// so a View (pointer, channel words, pixel pitch in words) addresses a P16 tensor exactly like an fp32 NHWC tensor of
// 16 * ngroups channels, channel slices start at multiples of 16, and a 32-channel chunk of a pixel is one 128-byte row
// [hi0 | c0 | hi1 | c1] (c = lo8 | hi8): TMA drops it into shared memory as an MMA-ready K-major tile (128B swizzle) whose four
// 32-byte K steps are hi(0..15) [f16], c(0..15) [fp8], hi(16..31) [f16], c(16..31) [fp8].  No kernel ever splits operands in
// shared memory.  Pad channels of the last group hold zeros.
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
// four fp32 -> four e5m2 bytes (round to nearest even, saturating at the largest finite value), byte 0 = first value
__device__ __forceinline__ uint32_t pack_e5m2x4(float a, float b, float c, float d) {
    uint32_t r;
    asm("{\n\t.reg .b16 l, u;\n\tcvt.rn.satfinite.e5m2x2.f32 l, %2, %1;\n\tcvt.rn.satfinite.e5m2x2.f32 u, %4, %3;\n\tmov.b32 %0, {l, u};\n\t}"
        : "=r"(r) : "f"(a), "f"(b), "f"(c), "f"(d));
    return r;
}
// an e5m2 byte is the upper byte of the fp16 number of the same value
__device__ __forceinline__ void unpack_e5m2x4(uint32_t w, float& a, float& b, float& c, float& d) {
    unpack2(__byte_perm(w, 0u, 0x1404), a, b);
    unpack2(__byte_perm(w, 0u, 0x3424), c, d);
}
// the scaled residuals of four values against their packed hi halves, as e5m2 bytes
__device__ __forceinline__ uint32_t pack_lo4(float a, float b, float c, float d, uint32_t h01, uint32_t h23) {
    float ha, hb, hc, hd;
    unpack2(h01, ha, hb);
    unpack2(h23, hc, hd);
    return pack_e5m2x4((a - ha) * LO_SCALE, (b - hb) * LO_SCALE, (c - hc) * LO_SCALE, (d - hd) * LO_SCALE);
}
// four channels: hi words (h01, h23) and the four lo8 bytes -> fp32
__device__ __forceinline__ void decode4(uint32_t h01, uint32_t h23, uint32_t lo4, float* v) {
    float h[4], l[4];
    unpack2(h01, h[0], h[1]);
    unpack2(h23, h[2], h[3]);
    unpack_e5m2x4(lo4, l[0], l[1], l[2], l[3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaf(l[j], LO_INV, h[j]);
}
// 8 channels (half a group): hi as one 16-byte vector, lo8 as 8 bytes
__device__ __forceinline__ void decode8(const uint4& h, const uint2& lo8, float* v) {
    decode4(h.x, h.y, lo8.x, v);
    decode4(h.z, h.w, lo8.y, v + 4);
}
__device__ __forceinline__ void encode8(const float* v, uint4& h, uint2& lo8, uint2& hi8) {
    h.x = pack_hi(v[0], v[1]); h.y = pack_hi(v[2], v[3]); h.z = pack_hi(v[4], v[5]); h.w = pack_hi(v[6], v[7]);
    lo8.x = pack_lo4(v[0], v[1], v[2], v[3], h.x, h.y);
    lo8.y = pack_lo4(v[4], v[5], v[6], v[7], h.z, h.w);
    hi8.x = pack_e5m2x4(v[0], v[1], v[2], v[3]);
    hi8.y = pack_e5m2x4(v[4], v[5], v[6], v[7]);
}
// a whole 16-channel group: the four 16-byte vectors of its 64 bytes in memory order (hi 0-7, hi 8-15, lo8 0-15, hi8 0-15)
__device__ __forceinline__ void encode16(const float* r, uint4& h0, uint4& h1, uint4& lo8, uint4& hi8) {
    uint2 l, g;
    encode8(r, h0, l, g);
    lo8.x = l.x; lo8.y = l.y; hi8.x = g.x; hi8.y = g.y;
    encode8(r + 8, h1, l, g);
    lo8.z = l.x; lo8.w = l.y; hi8.z = g.x; hi8.w = g.y;
}
// exponent field all ones in either half of a packed f16x2 word <=> inf or NaN: bit 15 / 31 of the result
__device__ __forceinline__ uint32_t nonfinite_bits(uint32_t h) { return (h & 0x7C007C00u) + 0x04000400u; }
__device__ __forceinline__ uint32_t nonfinite_bits(const uint4& h) {
    return nonfinite_bits(h.x) | nonfinite_bits(h.y) | nonfinite_bits(h.z) | nonfinite_bits(h.w);
}
__device__ __forceinline__ bool any_nonfinite(uint32_t acc) { return (acc & 0x80008000u) != 0u; }

// byte offsets inside a pixel row of the three vectors of 8-channel unit u (u = channel / 8): hi 16 bytes, lo8 / hi8 8 bytes each
__device__ __forceinline__ int unit_off_bytes(int u) { return (u >> 1) * 64 + (u & 1) * 16; }
__device__ __forceinline__ int unit_lo8_bytes(int u) { return (u >> 1) * 64 + 32 + (u & 1) * 8; }
__device__ __forceinline__ int unit_hi8_bytes(int u) { return (u >> 1) * 64 + 48 + (u & 1) * 8; }

}  // namespace p16
