// Host-side Goldilocks field and its degree-2 binomial extension.
//
// Mirrors the types the reference wires at src/types.rs:24-27
// (`Val = Goldilocks`, `ExtVal = BinomialExtensionField<Val, 2>`); the arithmetic itself lives
// in Plonky3 (p3-goldilocks / p3-field 0.5.1, rev e9d75614, not vendored in the reference).
// Published facts restated here: p = 2^64 - 2^32 + 1, GENERATOR = 7, TWO_ADICITY = 32,
// two_adic_generator(32) = 1753635133440165772, extension X^2 = 7, basis [1, X].
//
// Values are kept canonical (in [0, p)) at all times on the host, which is the representative
// p3 uses for equality, serde and byte serialisation.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <cassert>

namespace msh {

using u8 = uint8_t;
using u32 = uint32_t;
using u64 = uint64_t;
using u128 = unsigned __int128;

constexpr u64 GL_P = 0xFFFFFFFF00000001ULL;
constexpr u64 GL_EPS = 0xFFFFFFFFULL;  // 2^64 mod p
constexpr u64 GL_GENERATOR = 7;
constexpr unsigned GL_TWO_ADICITY = 32;
constexpr u64 GL_TWO_ADIC_ROOT_32 = 1753635133440165772ULL;  // generator of the 2^32 subgroup
constexpr u64 GL_EXT_W = 7;                                   // X^2 = W

static inline u64 gl_reduce128(u128 x) {
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    u64 t0;
    if (__builtin_sub_overflow(lo, hi_hi, &t0)) t0 -= GL_EPS;
    u64 t1 = hi_lo * GL_EPS;
    u64 r;
    if (__builtin_add_overflow(t0, t1, &r)) r += GL_EPS;
    if (r >= GL_P) r -= GL_P;
    return r;
}

struct Fp {
    u64 v;
    constexpr Fp() : v(0) {}
    constexpr explicit Fp(u64 x) : v(x >= GL_P ? x - GL_P : x) {}
    static constexpr Fp zero() { return Fp(); }
    static constexpr Fp one() { return Fp(1); }
    static Fp from_i64(int64_t x) { return x >= 0 ? Fp((u64)x) : -Fp((u64)(-x)); }
    bool is_zero() const { return v == 0; }
    friend bool operator==(Fp a, Fp b) { return a.v == b.v; }
    friend bool operator!=(Fp a, Fp b) { return a.v != b.v; }
    friend bool operator<(Fp a, Fp b) { return a.v < b.v; }
    friend Fp operator+(Fp a, Fp b) {
        u64 s = a.v + b.v;  // a.v, b.v < p so at most one wrap
        if (s < a.v || s >= GL_P) s -= GL_P;
        Fp r; r.v = s; return r;
    }
    friend Fp operator-(Fp a, Fp b) {
        Fp r; r.v = a.v >= b.v ? a.v - b.v : a.v + (GL_P - b.v); return r;
    }
    Fp operator-() const { Fp r; r.v = v ? GL_P - v : 0; return r; }
    friend Fp operator*(Fp a, Fp b) { Fp r; r.v = gl_reduce128((u128)a.v * b.v); return r; }
    Fp& operator+=(Fp o) { return *this = *this + o; }
    Fp& operator-=(Fp o) { return *this = *this - o; }
    Fp& operator*=(Fp o) { return *this = *this * o; }
    Fp square() const { return *this * *this; }
    Fp pow(u64 e) const {
        Fp base = *this, acc = one();
        while (e) { if (e & 1) acc *= base; base = base.square(); e >>= 1; }
        return acc;
    }
    Fp exp_power_of_2(unsigned k) const { Fp r = *this; while (k--) r = r.square(); return r; }
    Fp inverse() const { assert(v != 0); return pow(GL_P - 2); }
    Fp halve() const { Fp r; r.v = (v & 1) ? (v >> 1) + (GL_P >> 1) + 1 : v >> 1; return r; }
};

static_assert(sizeof(Fp) == 8, "Fp must be layout-compatible with uint64_t");

// two_adic_generator(bits): generator of the order-2^bits subgroup (p3 TwoAdicField).
static inline Fp two_adic_generator(unsigned bits) {
    assert(bits <= GL_TWO_ADICITY);
    return Fp(GL_TWO_ADIC_ROOT_32).exp_power_of_2(GL_TWO_ADICITY - bits);
}

// Montgomery batch inversion (p3_field::batch_multiplicative_inverse semantics; all inputs non-zero).
template <class T>
static inline void batch_inverse(const T* in, T* out, size_t n) {
    if (n == 0) return;
    std::vector<T> pref(n);
    T acc = T::one();
    for (size_t i = 0; i < n; i++) { pref[i] = acc; acc = acc * in[i]; }
    T inv = acc.inverse();
    for (size_t i = n; i-- > 0;) { T x = in[i]; out[i] = inv * pref[i]; inv = inv * x; }
}

// BinomialExtensionField<Goldilocks, 2>: a0 + a1*X, X^2 = 7.
struct Fp2 {
    Fp c[2];
    constexpr Fp2() : c{Fp(), Fp()} {}
    constexpr Fp2(Fp a0, Fp a1) : c{a0, a1} {}
    explicit constexpr Fp2(Fp a0) : c{a0, Fp()} {}
    static constexpr Fp2 zero() { return Fp2(); }
    static constexpr Fp2 one() { return Fp2(Fp(1), Fp()); }
    bool is_zero() const { return c[0].is_zero() && c[1].is_zero(); }
    friend bool operator==(const Fp2& a, const Fp2& b) { return a.c[0] == b.c[0] && a.c[1] == b.c[1]; }
    friend bool operator!=(const Fp2& a, const Fp2& b) { return !(a == b); }
    friend Fp2 operator+(const Fp2& a, const Fp2& b) { return Fp2(a.c[0] + b.c[0], a.c[1] + b.c[1]); }
    friend Fp2 operator-(const Fp2& a, const Fp2& b) { return Fp2(a.c[0] - b.c[0], a.c[1] - b.c[1]); }
    Fp2 operator-() const { return Fp2(-c[0], -c[1]); }
    friend Fp2 operator*(const Fp2& a, const Fp2& b) {
        Fp v0 = a.c[0] * b.c[0], v1 = a.c[1] * b.c[1];
        Fp cross = (a.c[0] + a.c[1]) * (b.c[0] + b.c[1]) - v0 - v1;
        return Fp2(v0 + v1 * Fp(GL_EXT_W), cross);
    }
    friend Fp2 operator*(const Fp2& a, Fp b) { return Fp2(a.c[0] * b, a.c[1] * b); }
    friend Fp2 operator*(Fp b, const Fp2& a) { return a * b; }
    friend Fp2 operator+(const Fp2& a, Fp b) { return Fp2(a.c[0] + b, a.c[1]); }
    friend Fp2 operator-(const Fp2& a, Fp b) { return Fp2(a.c[0] - b, a.c[1]); }
    Fp2& operator+=(const Fp2& o) { return *this = *this + o; }
    Fp2& operator-=(const Fp2& o) { return *this = *this - o; }
    Fp2& operator*=(const Fp2& o) { return *this = *this * o; }
    Fp2 square() const { return *this * *this; }
    Fp2 pow(u64 e) const {
        Fp2 base = *this, acc = one();
        while (e) { if (e & 1) acc *= base; base = base.square(); e >>= 1; }
        return acc;
    }
    Fp2 exp_power_of_2(unsigned k) const { Fp2 r = *this; while (k--) r = r.square(); return r; }
    Fp2 inverse() const {
        // (a0 - a1 X) / (a0^2 - W a1^2)
        Fp norm = c[0] * c[0] - Fp(GL_EXT_W) * c[1] * c[1];
        Fp ni = norm.inverse();
        return Fp2(c[0] * ni, (-c[1]) * ni);
    }
    Fp2 halve() const { return Fp2(c[0].halve(), c[1].halve()); }
};

static inline unsigned log2_strict(size_t n) {
    assert(n && (n & (n - 1)) == 0);
    unsigned l = 0;
    while ((size_t(1) << l) < n) l++;
    return l;
}
static inline size_t reverse_bits_len(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | ((x >> i) & 1); }
    return r;
}

// Row-major matrix of base-field elements (p3_matrix::dense::RowMajorMatrix<Val>).
struct Matrix {
    std::vector<Fp> values;
    size_t width = 0;
    Matrix() = default;
    Matrix(std::vector<Fp> v, size_t w) : values(std::move(v)), width(w) {
        assert(width == 0 ? values.empty() : values.size() % width == 0);
    }
    Matrix(size_t h, size_t w) : values(h * w), width(w) {}
    size_t height() const { return width ? values.size() / width : 0; }
    Fp* row(size_t r) { return values.data() + r * width; }
    const Fp* row(size_t r) const { return values.data() + r * width; }
};

// Borrowed row-major matrix (the caller's memory; `Fp` is layout-compatible with a canonical uint64_t).
struct MatrixView {
    const Fp* data = nullptr;
    size_t h = 0, width = 0;
    MatrixView() = default;
    MatrixView(const Fp* d, size_t height, size_t w) : data(d), h(height), width(w) {}
    MatrixView(const Matrix& m) : data(m.values.data()), h(m.height()), width(m.width) {}
    size_t height() const { return h; }
    const Fp* row(size_t r) const { return data + r * width; }
    Matrix to_matrix() const { return Matrix(std::vector<Fp>(data, data + h * width), width); }
};

// Borrowed claims: `values` flat, claim i = values[offsets[i] .. offsets[i+1])  (src/prover.rs:289-294 `claims: &[&[Val]]`).
// offsets == nullptr: every claim has `stride` values (one call shape, the common case: no offsets array to build or scan).
struct ClaimsView {
    const Fp* values = nullptr;
    const u64* offsets = nullptr;
    size_t n = 0;
    size_t stride = 0;
    size_t size() const { return n; }
    size_t len(size_t i) const { return offsets ? (size_t)(offsets[i + 1] - offsets[i]) : stride; }
    const Fp* at(size_t i) const { return values + (offsets ? (size_t)offsets[i] : i * stride); }
    size_t total() const { return n == 0 ? 0 : (offsets ? (size_t)offsets[n] : n * stride); }
    // all claims of one length, 0 if empty or ragged
    size_t uniform_len() const {
        if (n == 0) return 0;
        if (!offsets) return stride;
        size_t l = len(0);
        for (size_t i = 1; i < n; i++)
            if (len(i) != l) return 0;
        return l;
    }
};

}  // namespace msh
