// BLAKE3 compression on the device: state in 16 registers, 7 fully unrolled rounds, the message
// schedule resolved at compile time. Replaces p3-blake3 `Blake3` as used by the reference's
// `SerializingHasher<Blake3>` and `CompressionFunctionFromHasher<Blake3, 2, 32>` (src/types.rs:83,199).
// Known-answer vector: src/test_circuits/blake3.rs:2646-2746 (tests/test_gpu_merkle.py).
#pragma once
#include "gl.cuh"

namespace b3 {

#define B3_IV0 0x6A09E667u
#define B3_IV1 0xBB67AE85u
#define B3_IV2 0x3C6EF372u
#define B3_IV3 0xA54FF53Au
#define B3_IV4 0x510E527Fu
#define B3_IV5 0x9B05688Cu
#define B3_IV6 0x1F83D9ABu
#define B3_IV7 0x5BE0CD19u

enum : u32 { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

__device__ __forceinline__ u32 rotr(u32 x, int n) { return __funnelshift_r(x, x, n); }

// Pipe balance: a G function is 4 XOR (LOP3) + 4 rotates (SHF) + 6 additions. ptxas puts the two 3-input additions on the
// ALU pipe as IADD3 next to the LOP3/SHF (10 ALU : 2 FMA-pipe instructions per G, the ALU pipe is the bound). Written as
// multiply-adds by a value the compiler cannot see is 1 (gridDim.z of our 1-D / 2-D launches) all six additions go to the
// FMA pipe as IMAD: 8 ALU : 6 FMA per G (leaf hashing: 0.63 -> 0.53 ms per bench step). The message word is added first,
// off the a -> d -> c -> b dependency chain. `one` = 1 as a literal gives the plain form, which has the shorter dependency
// chain and is kept where a CTA has little parallelism (the upper levels of k_merkle_subtree).
#define B3_G(a, b, c, d, mx, my) \
    a = (mx) * one + a;          \
    a = b * one + a;             \
    d = rotr(d ^ a, 16);         \
    c = d * one + c;             \
    b = rotr(b ^ c, 12);         \
    a = (my) * one + a;          \
    a = b * one + a;             \
    d = rotr(d ^ a, 8);          \
    c = d * one + c;             \
    b = rotr(b ^ c, 7);

#define B3_ROUND(m0, m1, m2, m3, m4, m5, m6, m7, m8, m9, m10, m11, m12, m13, m14, m15) \
    B3_G(s0, s4, s8, s12, m[m0], m[m1])                                                  \
    B3_G(s1, s5, s9, s13, m[m2], m[m3])                                                  \
    B3_G(s2, s6, s10, s14, m[m4], m[m5])                                                 \
    B3_G(s3, s7, s11, s15, m[m6], m[m7])                                                 \
    B3_G(s0, s5, s10, s15, m[m8], m[m9])                                                 \
    B3_G(s1, s6, s11, s12, m[m10], m[m11])                                               \
    B3_G(s2, s7, s8, s13, m[m12], m[m13])                                                \
    B3_G(s3, s4, s9, s14, m[m14], m[m15])

// cv (in/out) <- first 8 words of compress(cv, m, counter, block_len, flags). `m` must be indexed
// with compile-time constants only (it lives in registers).
template <bool kFmaAdds = true>
__device__ __forceinline__ void compress(u32 cv[8], const u32 m[16], u32 counter_lo, u32 counter_hi, u32 block_len,
                                         u32 flags) {
    u32 s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    u32 s8 = B3_IV0, s9 = B3_IV1, s10 = B3_IV2, s11 = B3_IV3;
    u32 s12 = counter_lo, s13 = counter_hi, s14 = block_len, s15 = flags;
    const u32 one = kFmaAdds ? gridDim.z : 1u;
    B3_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B3_ROUND(2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
    B3_ROUND(3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1)
    B3_ROUND(10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6)
    B3_ROUND(12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4)
    B3_ROUND(9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7)
    B3_ROUND(11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13)
    cv[0] = s0 ^ s8;
    cv[1] = s1 ^ s9;
    cv[2] = s2 ^ s10;
    cv[3] = s3 ^ s11;
    cv[4] = s4 ^ s12;
    cv[5] = s5 ^ s13;
    cv[6] = s6 ^ s14;
    cv[7] = s7 ^ s15;
}

// Raw 16-word form (state in, full 16 words out) for the reference's known-answer vector.
__device__ __forceinline__ void compress_raw(const u32 st[16], const u32 m[16], u32 out[16]) {
    u32 s0 = st[0], s1 = st[1], s2 = st[2], s3 = st[3], s4 = st[4], s5 = st[5], s6 = st[6], s7 = st[7];
    u32 s8 = st[8], s9 = st[9], s10 = st[10], s11 = st[11], s12 = st[12], s13 = st[13], s14 = st[14], s15 = st[15];
    const u32 one = gridDim.z;
    B3_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B3_ROUND(2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
    B3_ROUND(3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1)
    B3_ROUND(10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6)
    B3_ROUND(12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4)
    B3_ROUND(9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7)
    B3_ROUND(11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13)
    out[0] = s0 ^ s8; out[1] = s1 ^ s9; out[2] = s2 ^ s10; out[3] = s3 ^ s11;
    out[4] = s4 ^ s12; out[5] = s5 ^ s13; out[6] = s6 ^ s14; out[7] = s7 ^ s15;
    out[8] = s8 ^ st[0]; out[9] = s9 ^ st[1]; out[10] = s10 ^ st[2]; out[11] = s11 ^ st[3];
    out[12] = s12 ^ st[4]; out[13] = s13 ^ st[5]; out[14] = s14 ^ st[6]; out[15] = s15 ^ st[7];
}

__device__ __forceinline__ void set_iv(u32 cv[8]) {
    cv[0] = B3_IV0; cv[1] = B3_IV1; cv[2] = B3_IV2; cv[3] = B3_IV3;
    cv[4] = B3_IV4; cv[5] = B3_IV5; cv[6] = B3_IV6; cv[7] = B3_IV7;
}

// out = BLAKE3(l || r): one 64-byte block, CHUNK_START | CHUNK_END | ROOT.
template <bool kFmaAdds = false>
__device__ __forceinline__ void hash_pair(const u32 l[8], const u32 r[8], u32 out[8]) {
    u32 m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { m[i] = l[i]; m[8 + i] = r[i]; }
    set_iv(out);
    compress<kFmaAdds>(out, m, 0, 0, 64, CHUNK_START | CHUNK_END | ROOT);
}

}  // namespace b3
