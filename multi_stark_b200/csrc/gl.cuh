// Goldilocks (p = 2^64 - 2^32 + 1) and its quadratic extension X^2 = 7 on the device.
//
// Replaces, for the device side, the arithmetic the reference gets from p3-goldilocks / p3-field
// (`Val`, `ExtVal` at src/types.rs:24-27). INVARIANT: every value held in HBM and every value
// returned by the functions below is CANONICAL (in [0, p)); p3 hashes, observes and serialises
// the canonical representative, so device buffers can be hashed / copied out as they are.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define GLD_P 0xFFFFFFFF00000001ull
#define GLD_EPS 0xFFFFFFFFull

namespace gl {

// All arithmetic is written as 32-bit carry chains (add.cc / addc / madc): a conditional
// "+- p" becomes an add of the carry mask (2^64 - p = 2^32 - 1 has a single non-zero word), so there are
// no 64-bit compares or selects on the hot path.

__device__ __forceinline__ u64 pack(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }

// x any u64 -> canonical
__device__ __forceinline__ u64 canon(u64 x) {
    u32 x0 = (u32)x, x1 = (u32)(x >> 32), r0, r1;
    asm("{\n\t"
        ".reg .u32 t, k;\n\t"
        "add.cc.u32 t, %2, 0xffffffff;\n\t"   // x + (2^32 - 1) carries out of 64 bits iff x >= p
        "addc.cc.u32 t, %3, 0;\n\t"
        "addc.u32 k, 0, 0;\n\t"
        "neg.s32 k, k;\n\t"                   // mask
        "add.cc.u32 %0, %2, k;\n\t"           // x - p == x + (2^32 - 1) mod 2^64
        "addc.u32 %1, %3, 0;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(x0), "r"(x1));
    return pack(r0, r1);
}

// a any u64, b canonical (< p) -> some u64 congruent to a + b (NOT necessarily canonical)
__device__ __forceinline__ u64 add_lazy(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), r0, r1;
    asm("{\n\t"
        ".reg .u32 k;\n\t"
        "add.cc.u32 %0, %2, %4;\n\t"
        "addc.cc.u32 %1, %3, %5;\n\t"
        "addc.u32 k, 0, 0;\n\t"
        "neg.s32 k, k;\n\t"
        "add.cc.u32 %0, %0, k;\n\t"   // + (2^32 - 1) when the 64-bit sum wrapped; cannot wrap again since b < p
        "addc.u32 %1, %1, 0;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return pack(r0, r1);
}
// a any u64, b canonical (< p) -> some u64 congruent to a - b
__device__ __forceinline__ u64 sub_lazy(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), r0, r1;
    asm("{\n\t"
        ".reg .u32 m;\n\t"
        "sub.cc.u32 %0, %2, %4;\n\t"
        "subc.cc.u32 %1, %3, %5;\n\t"
        "subc.u32 m, 0, 0;\n\t"       // 0xffffffff on borrow
        "sub.cc.u32 %0, %0, m;\n\t"   // - (2^32 - 1) == + p; cannot borrow again since b < p
        "subc.u32 %1, %1, 0;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return pack(r0, r1);
}

// a, b canonical -> canonical: a + b = a - (p - b), one borrow fix-up (7 instructions; p - b is in (0, p], which sub_lazy
// tolerates)
__device__ __forceinline__ u64 add(u64 a, u64 b) { return sub_lazy(a, GLD_P - b); }
// a, b canonical -> canonical (a - b + p < 2^64 and < p after the fix, so sub_lazy is already canonical)
__device__ __forceinline__ u64 sub(u64 a, u64 b) { return sub_lazy(a, b); }
__device__ __forceinline__ u64 neg(u64 a) { return a ? GLD_P - a : 0; }

// 128-bit (hi:lo) -> u64 congruent mod p (not canonical). 2^64 = 2^32 - 1, 2^96 = -1 (mod p).
__device__ __forceinline__ u64 reduce128_lazy(u64 lo, u64 hi) {
    u32 l0 = (u32)lo, l1 = (u32)(lo >> 32), hl = (u32)hi, hh = (u32)(hi >> 32), r0, r1;
    asm("{\n\t"
        ".reg .u32 m, k, t0, t1;\n\t"
        // t = lo - hh (+p on borrow)
        "sub.cc.u32 %0, %2, %5;\n\t"
        "subc.cc.u32 %1, %3, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\t"
        "subc.u32 %1, %1, 0;\n\t"
        // u = hl * (2^32 - 1) = (hl << 32) - hl
        "sub.cc.u32 t0, 0, %4;\n\t"
        "subc.u32 t1, %4, 0;\n\t"
        // r = t + u (+ (2^32 - 1) on carry; cannot carry twice, u <= (2^32-1)^2)
        "add.cc.u32 %0, %0, t0;\n\t"
        "addc.cc.u32 %1, %1, t1;\n\t"
        "addc.u32 k, 0, 0;\n\t"
        "neg.s32 k, k;\n\t"
        "add.cc.u32 %0, %0, k;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(l0), "r"(l1), "r"(hl), "r"(hh));
    return pack(r0, r1);
}
namespace gf {
__device__ __forceinline__ u64 reduce128(u64 lo, u64 hi);
}
// any u64 operands (not necessarily canonical) -> canonical (26 instructions: one 128-bit product, lean reduction)
__device__ __forceinline__ u64 mul(u64 a, u64 b) {
    unsigned __int128 pr = (unsigned __int128)a * b;
    return gf::reduce128((u64)pr, (u64)(pr >> 64));
}
// any u64 operands -> some congruent u64
__device__ __forceinline__ u64 mul_lazy(u64 a, u64 b) { return reduce128_lazy(a * b, __umul64hi(a, b)); }
__device__ __forceinline__ u64 sqr(u64 a) { return mul(a, a); }

__device__ __forceinline__ u64 pow(u64 a, u64 e) {
    u64 acc = 1;
    while (e) {
        if (e & 1) acc = mul(acc, a);
        a = sqr(a);
        e >>= 1;
    }
    return acc;
}
__device__ __forceinline__ u64 inv(u64 a) { return pow(a, GLD_P - 2); }
__device__ __forceinline__ u64 halve(u64 a) { return (a & 1) ? (a >> 1) + (GLD_P >> 1) + 1 : a >> 1; }

// ---- lean canonical arithmetic for the NTT butterflies (canonical in, canonical out) ---------------------------
// Instruction counts (sm_100a SASS): add 7, sub 5, mul 26 (4 IMAD.WIDE + carries + reduction), against 14 / 5 / 34 of
// the general-purpose forms above.
namespace gf {
__device__ __forceinline__ u64 sub(u64 a, u64 b) { return sub_lazy(a, b); }
// a + b = a - (p - b): one borrow fix-up makes the result canonical (p - b is in (0, p], which sub_lazy tolerates)
__device__ __forceinline__ u64 add(u64 a, u64 b) { return sub_lazy(a, GLD_P - b); }
// 128-bit (hi:lo) -> canonical. 2^64 = 2^32 - 1 =: eps, 2^96 = -1 (mod p).
__device__ __forceinline__ u64 reduce128(u64 lo, u64 hi) {
    u32 l0 = (u32)lo, l1 = (u32)(lo >> 32), h0 = (u32)hi, h1 = (u32)(hi >> 32);
    u64 r;
    asm("{\n\t"
        ".reg .u32 t0, t1, m, k, z, r0, r1;\n\t"
        ".reg .u64 rr;\n\t"
        "sub.cc.u32 t0, %1, %4;\n\t"          // t = lo - h1 (- eps on borrow)
        "subc.cc.u32 t1, %2, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 t0, t0, m;\n\t"
        "subc.u32 t1, t1, 0;\n\t"
        "mad.lo.cc.u32 r0, %3, 0xffffffff, t0;\n\t"   // r = t + h0 * eps (+ eps on carry)
        "madc.hi.cc.u32 r1, %3, 0xffffffff, t1;\n\t"
        "addc.u32 k, 0, 0;\n\t"
        "mov.b64 rr, {r0, r1};\n\t"
        "mad.wide.u32 rr, k, 0xffffffff, rr;\n\t"
        "mov.b64 {r0, r1}, rr;\n\t"
        "add.cc.u32 z, r0, 0xffffffff;\n\t"   // r >= p  <=>  r + eps carries; then r - p = r + eps (mod 2^64)
        "addc.cc.u32 z, r1, 0;\n\t"
        "addc.u32 k, 0, 0;\n\t"
        "mad.wide.u32 %0, k, 0xffffffff, rr;\n\t"
        "}"
        : "=l"(r)
        : "r"(l0), "r"(l1), "r"(h0), "r"(h1));
    return r;
}
__device__ __forceinline__ u64 mul(u64 a, u64 b) {
    unsigned __int128 pr = (unsigned __int128)a * b;
    return reduce128((u64)pr, (u64)(pr >> 64));
}
// 2^e mod p for 0 <= e < 96 (folds to a constant when e is known at compile time)
__host__ __device__ constexpr u64 pow2_mod_p(int e) {
    return e < 64 ? (1ull << e) : ((1ull << (e - 32)) - (1ull << (e - 64)));  // 2^64 = 2^32 - 1
}
// x * 2^e mod p for 0 < e < 96 with shifts instead of a multiplication: x << (e % 32) is three 32-bit limbs y0, y1, y2 at
// limb position q = e / 32, and 2^64 = eps, 2^96 = -1, 2^128 = -2^32. x may be any u64; the result is canonical.
// `e` must be a compile-time constant after inlining (11 to 17 instructions against 22 for a constant multiplier).
__device__ __forceinline__ u64 mul_pow2(u64 x, int e) {
    const int q = e / 32, r = e % 32;
    const u32 x0 = (u32)x, x1 = (u32)(x >> 32);
    u32 y0, y1, y2;
    if (r == 0) {
        y0 = x0; y1 = x1; y2 = 0;
    } else {
        y0 = x0 << r; y1 = __funnelshift_l(x0, x1, r); y2 = x1 >> (32 - r);
    }
    if (q == 2) return sub_lazy((u64)y0 * GLD_EPS, ((u64)y2 << 32) | y1);          // y0 eps - (y1 + y2 2^32)
    if (q == 0) return sub_lazy(((u64)y1 << 32) | y0, GLD_P - (u64)y2 * GLD_EPS);  // (y0 + y1 2^32) + y2 eps
    return sub_lazy(sub_lazy((u64)y0 << 32, (u64)y2), GLD_P - (u64)y1 * GLD_EPS);  // y0 2^32 - y2 + y1 eps
}
}  // namespace gf

// ---- lazy multiply-accumulate: s += a * b as a 160-bit integer, reduced once at the end ---------------------------------
// For the dot products of the opening phase (sum over columns of alpha^c * M[c], barycentric sums): 13 instructions per
// term instead of a full modular multiplication and addition (about 33). Exact for up to 2^32 terms.
struct Acc160 {
    u32 l0, l1, l2, l3, l4;
};
__device__ __forceinline__ Acc160 acc_zero() {
    Acc160 s;
    s.l0 = s.l1 = s.l2 = s.l3 = s.l4 = 0;
    return s;
}
__device__ __forceinline__ void acc_mac(Acc160& s, u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t"
        "mad.lo.cc.u32 %0, %5, %7, %0;\n\t"    // a0 b0 at limb 0
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %8, %2;\n\t"   // a1 b1 at limb 2
        "madc.hi.cc.u32 %3, %6, %8, %3;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %8, %1;\n\t"    // a0 b1 at limb 1
        "madc.hi.cc.u32 %2, %5, %8, %2;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %1, %6, %7, %1;\n\t"    // a1 b0 at limb 1
        "madc.hi.cc.u32 %2, %6, %7, %2;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "}"
        : "+r"(s.l0), "+r"(s.l1), "+r"(s.l2), "+r"(s.l3), "+r"(s.l4)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}
// canonical value of the accumulator: 2^64 = eps, 2^96 = -1, 2^128 = -2^32 (mod p)
__device__ __forceinline__ u64 acc_reduce(const Acc160& s) {
    u64 r = gf::reduce128(pack(s.l0, s.l1), pack(s.l2, s.l3));
    return sub_lazy(r, (u64)s.l4 << 32);
}

// ---- extension field F_p[X]/(X^2 - 7) -------------------------------------------------------
struct e2 {
    u64 a, b;  // a + b X
};
__device__ __forceinline__ e2 e2_make(u64 a, u64 b) { e2 r; r.a = a; r.b = b; return r; }
__device__ __forceinline__ e2 e2_add(e2 x, e2 y) { return e2_make(add(x.a, y.a), add(x.b, y.b)); }
__device__ __forceinline__ e2 e2_sub(e2 x, e2 y) { return e2_make(sub(x.a, y.a), sub(x.b, y.b)); }
__device__ __forceinline__ u64 mul7(u64 x) {
    u64 x2 = add(x, x), x4 = add(x2, x2);
    return sub(add(x4, x4), x);
}
__device__ __forceinline__ e2 e2_mul(e2 x, e2 y) {
    u64 v0 = mul(x.a, y.a), v1 = mul(x.b, y.b);
    u64 cross = sub(sub(mul(add(x.a, x.b), add(y.a, y.b)), v0), v1);
    return e2_make(add(v0, mul7(v1)), cross);
}
__device__ __forceinline__ e2 e2_mul_base(e2 x, u64 s) { return e2_make(mul(x.a, s), mul(x.b, s)); }
__device__ __forceinline__ e2 e2_inv(e2 x) {
    u64 norm = sub(sqr(x.a), mul7(sqr(x.b)));
    u64 ni = inv(norm);
    return e2_make(mul(x.a, ni), mul(neg(x.b), ni));
}

// Two-level power table: value(e) = c * g^e = hi[e >> h1] * lo[e & (2^h1 - 1)], lo carries c.
struct PowTable {
    const u64* lo;
    const u64* hi;
    u32 h1;
};
__device__ __forceinline__ u64 pow_lookup(const PowTable& t, u64 e) {
    return mul(__ldg(t.hi + (e >> t.h1)), __ldg(t.lo + (e & ((1ull << t.h1) - 1))));
}

__device__ __forceinline__ u32 rev_bits(u32 x, u32 bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

}  // namespace gl
