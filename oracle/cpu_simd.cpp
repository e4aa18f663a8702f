// ORACLE (test infrastructure, not product code): the vectorised kernels declared in cpu_simd.hpp, one definition per
// instruction-set clone (GCC function multi-versioning; the dynamic loader picks the clone for the CPU it runs on).
#include "cpu_simd.hpp"
#include <cstdlib>

#if defined(__x86_64__) && defined(__GNUC__) && !defined(ORC_NO_SIMD_CLONES)
#define ORC_SIMD __attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))
#else
#define ORC_SIMD
#endif

namespace orc {
namespace simd {

// a[c], b[c] <- a[c] + b[c], (a[c] - b[c]) * t     (one decimation-in-frequency butterfly per column)
ORC_SIMD static void dif_butterfly_row_auto(u64* __restrict a, u64* __restrict b, u64 t, size_t w) {
#pragma omp simd
    for (size_t c = 0; c < w; c++) {
        u64 x = a[c], y = b[c];
        a[c] = gl_add(x, y);
        b[c] = gl_mul(gl_sub(x, y), t);
    }
}
// the same without the multiplication (twiddle 1)
ORC_SIMD static void dif_butterfly_row_notw_auto(u64* __restrict a, u64* __restrict b, size_t w) {
#pragma omp simd
    for (size_t c = 0; c < w; c++) {
        u64 x = a[c], y = b[c];
        a[c] = gl_add(x, y);
        b[c] = gl_sub(x, y);
    }
}
// dst[c] = src[c] * s
ORC_SIMD static void scale_row_auto(u64* __restrict dst, const u64* __restrict src, u64 s, size_t w) {
#pragma omp simd
    for (size_t c = 0; c < w; c++) dst[c] = gl_mul(src[c], s);
}


#define ORC_B3_G(a, b, c, d, mx, my)                                                                              \
    for (int l = 0; l < kLanes; l++) {                                                                            \
        u32 va = s[a][l], vb = s[b][l], vc = s[c][l], vd = s[d][l];                                               \
        va = va + vb + m[mx][l]; vd ^= va; vd = (vd >> 16) | (vd << 16);                                          \
        vc = vc + vd;            vb ^= vc; vb = (vb >> 12) | (vb << 20);                                          \
        va = va + vb + m[my][l]; vd ^= va; vd = (vd >> 8) | (vd << 24);                                           \
        vc = vc + vd;            vb ^= vc; vb = (vb >> 7) | (vb << 25);                                           \
        s[a][l] = va; s[b][l] = vb; s[c][l] = vc; s[d][l] = vd;                                                   \
    }

// cv[i][lane] (in/out) <- compress(cv, m, counter = 0, block_len[lane], flags[lane]) for every lane
ORC_SIMD void b3_compress_lanes(lane_t* cv, const lane_t* m, const u32* block_len, const u32* flags, const u32* counter_lo) {
    lane_t s[16];
    for (int i = 0; i < 8; i++)
        for (int l = 0; l < kLanes; l++) s[i][l] = cv[i][l];
    for (int i = 0; i < 4; i++)
        for (int l = 0; l < kLanes; l++) s[8 + i][l] = msh::b3::IV[i];
    for (int l = 0; l < kLanes; l++) { s[12][l] = counter_lo[l]; s[13][l] = 0; s[14][l] = block_len[l]; s[15][l] = flags[l]; }
    static const unsigned char sched[7][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8},
        {3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1}, {10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6},
        {12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4}, {9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7},
        {11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13}};
    for (int r = 0; r < 7; r++) {
        const unsigned char* p = sched[r];
        ORC_B3_G(0, 4, 8, 12, p[0], p[1])
        ORC_B3_G(1, 5, 9, 13, p[2], p[3])
        ORC_B3_G(2, 6, 10, 14, p[4], p[5])
        ORC_B3_G(3, 7, 11, 15, p[6], p[7])
        ORC_B3_G(0, 5, 10, 15, p[8], p[9])
        ORC_B3_G(1, 6, 11, 12, p[10], p[11])
        ORC_B3_G(2, 7, 8, 13, p[12], p[13])
        ORC_B3_G(3, 4, 9, 14, p[14], p[15])
    }
    for (int i = 0; i < 8; i++)
        for (int l = 0; l < kLanes; l++) cv[i][l] = s[i][l] ^ s[i + 8][l];
}
#undef ORC_B3_G
ORC_SIMD static void scale_row_inplace_auto(u64* row, u64 s, size_t w) {
#pragma omp simd
    for (size_t c = 0; c < w; c++) row[c] = gl_mul(row[c], s);
}


// ---- AVX-512 row kernels with MASKED tails: trace rows are 14 / 26 / 2 columns wide, so a scalar remainder loop would take
// as long as the vector body. 8 Goldilocks lanes per vector; the product is four vpmuludq, the reduction uses 2^64 = 2^32 - 1.
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define ORC_AVX512 __attribute__((target("avx512f,avx512dq,avx512vl,avx512bw")))
ORC_AVX512 static inline __m512i v_add(__m512i x, __m512i y, __m512i p) {
    __m512i s = _mm512_add_epi64(x, y);
    __mmask8 k = _mm512_cmplt_epu64_mask(s, x) | _mm512_cmpge_epu64_mask(s, p);
    return _mm512_mask_sub_epi64(s, k, s, p);
}
ORC_AVX512 static inline __m512i v_sub(__m512i x, __m512i y, __m512i p) {
    __m512i d = _mm512_sub_epi64(x, y);
    return _mm512_mask_add_epi64(d, _mm512_cmplt_epu64_mask(x, y), d, p);
}
ORC_AVX512 static inline __m512i v_mul(__m512i a, __m512i b, __m512i p, __m512i eps) {
    __m512i a1 = _mm512_srli_epi64(a, 32), b1 = _mm512_srli_epi64(b, 32);
    __m512i ll = _mm512_mul_epu32(a, b), lh = _mm512_mul_epu32(a, b1), hl = _mm512_mul_epu32(a1, b), hh = _mm512_mul_epu32(a1, b1);
    __m512i mid = _mm512_add_epi64(lh, _mm512_srli_epi64(ll, 32));
    __m512i mid2 = _mm512_add_epi64(hl, _mm512_and_si512(mid, eps));
    __m512i hi = _mm512_add_epi64(_mm512_add_epi64(hh, _mm512_srli_epi64(mid, 32)), _mm512_srli_epi64(mid2, 32));
    __m512i lo = _mm512_or_si512(_mm512_and_si512(ll, eps), _mm512_slli_epi64(mid2, 32));
    __m512i hi_hi = _mm512_srli_epi64(hi, 32), hi_lo = _mm512_and_si512(hi, eps);
    __m512i t0 = _mm512_sub_epi64(lo, hi_hi);
    t0 = _mm512_mask_sub_epi64(t0, _mm512_cmplt_epu64_mask(lo, hi_hi), t0, eps);
    __m512i t1 = _mm512_sub_epi64(_mm512_slli_epi64(hi_lo, 32), hi_lo);
    __m512i r = _mm512_add_epi64(t0, t1);
    r = _mm512_mask_add_epi64(r, _mm512_cmplt_epu64_mask(r, t1), r, eps);
    return _mm512_mask_sub_epi64(r, _mm512_cmpge_epu64_mask(r, p), r, p);
}
ORC_AVX512 static void dif_butterfly_row_512(u64* a, u64* b, u64 t, size_t w) {
    const __m512i p = _mm512_set1_epi64((long long)msh::GL_P), eps = _mm512_set1_epi64((long long)msh::GL_EPS), tv = _mm512_set1_epi64((long long)t);
    for (size_t c = 0; c < w; c += 8) {
        const __mmask8 k = w - c >= 8 ? (__mmask8)0xff : (__mmask8)((1u << (w - c)) - 1);
        __m512i x = _mm512_maskz_loadu_epi64(k, a + c), y = _mm512_maskz_loadu_epi64(k, b + c);
        _mm512_mask_storeu_epi64(a + c, k, v_add(x, y, p));
        _mm512_mask_storeu_epi64(b + c, k, v_mul(v_sub(x, y, p), tv, p, eps));
    }
}
ORC_AVX512 static void dif_butterfly_row_notw_512(u64* a, u64* b, size_t w) {
    const __m512i p = _mm512_set1_epi64((long long)msh::GL_P);
    for (size_t c = 0; c < w; c += 8) {
        const __mmask8 k = w - c >= 8 ? (__mmask8)0xff : (__mmask8)((1u << (w - c)) - 1);
        __m512i x = _mm512_maskz_loadu_epi64(k, a + c), y = _mm512_maskz_loadu_epi64(k, b + c);
        _mm512_mask_storeu_epi64(a + c, k, v_add(x, y, p));
        _mm512_mask_storeu_epi64(b + c, k, v_sub(x, y, p));
    }
}
ORC_AVX512 static void scale_row_512(u64* dst, const u64* src, u64 s, size_t w) {
    const __m512i p = _mm512_set1_epi64((long long)msh::GL_P), eps = _mm512_set1_epi64((long long)msh::GL_EPS), sv = _mm512_set1_epi64((long long)s);
    for (size_t c = 0; c < w; c += 8) {
        const __mmask8 k = w - c >= 8 ? (__mmask8)0xff : (__mmask8)((1u << (w - c)) - 1);
        _mm512_mask_storeu_epi64(dst + c, k, v_mul(_mm512_maskz_loadu_epi64(k, src + c), sv, p, eps));
    }
}
static const bool g_avx512 = [] {
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl") &&
           __builtin_cpu_supports("avx512bw") && !getenv("ORC_NO_AVX512");
}();
#else
static const bool g_avx512 = false;
static void dif_butterfly_row_512(u64*, u64*, u64, size_t) {}
static void dif_butterfly_row_notw_512(u64*, u64*, size_t) {}
static void scale_row_512(u64*, const u64*, u64, size_t) {}
#endif

void dif_butterfly_row(u64* __restrict a, u64* __restrict b, u64 t, size_t w) {
    if (g_avx512) dif_butterfly_row_512(a, b, t, w);
    else dif_butterfly_row_auto(a, b, t, w);
}
void dif_butterfly_row_notw(u64* __restrict a, u64* __restrict b, size_t w) {
    if (g_avx512) dif_butterfly_row_notw_512(a, b, w);
    else dif_butterfly_row_notw_auto(a, b, w);
}
void scale_row(u64* __restrict dst, const u64* __restrict src, u64 s, size_t w) {
    if (g_avx512) scale_row_512(dst, src, s, w);
    else scale_row_auto(dst, src, s, w);
}
void scale_row_inplace(u64* row, u64 s, size_t w) {
    if (g_avx512) scale_row_512(row, row, s, w);
    else scale_row_inplace_auto(row, s, w);
}

}  // namespace simd
}  // namespace orc

extern "C" {
// which clone the loader picked: 4 = x86-64-v4 (AVX-512), 3 = x86-64-v3 (AVX2), 1 = baseline
int orc_simd_level() {
#if defined(__x86_64__) && defined(__GNUC__)
    __builtin_cpu_init();
    if (__builtin_cpu_supports("x86-64-v4")) return 4;
    if (__builtin_cpu_supports("x86-64-v3")) return 3;
#endif
    return 1;
}
}
