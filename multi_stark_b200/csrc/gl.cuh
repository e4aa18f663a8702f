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

__device__ __forceinline__ u64 canon(u64 x) { return x >= GLD_P ? x - GLD_P : x; }

// a, b canonical -> canonical.
__device__ __forceinline__ u64 add(u64 a, u64 b) {
    u64 s = a + b;
    u64 t = s + GLD_EPS;  // == s - p (mod 2^64); carries out iff s >= p
    // a + b < 2p, so at most one subtraction of p is needed
    return (s < a || t < s) ? t : s;
}
__device__ __forceinline__ u64 sub(u64 a, u64 b) {
    u64 d = a - b;
    return a < b ? d - GLD_EPS : d;  // d + p (mod 2^64)
}
__device__ __forceinline__ u64 neg(u64 a) { return a ? GLD_P - a : 0; }

// 128-bit (hi:lo) -> canonical. 2^64 = 2^32 - 1, 2^96 = -1 (mod p).
__device__ __forceinline__ u64 reduce128(u64 lo, u64 hi) {
    u32 hh = (u32)(hi >> 32), hl = (u32)hi;
    u64 t0 = lo - hh;
    if (lo < hh) t0 -= GLD_EPS;
    u64 t1 = (u64)hl * GLD_EPS;
    u64 r = t0 + t1;
    if (r < t1) r += GLD_EPS;
    return canon(r);
}
// any u64 operands (not necessarily canonical) -> canonical
__device__ __forceinline__ u64 mul(u64 a, u64 b) { return reduce128(a * b, __umul64hi(a, b)); }
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
