// ORACLE (test infrastructure, not product code): CPU restatement of the two-adic DFT / coset LDE
// the reference obtains from Plonky3's `p3_dft::Radix2DitParallel` (type alias src/types.rs:200;
// call sites src/prover.rs:440,650,716 and, inside `TwoAdicFriPcs::commit`, src/prover.rs:350,419,
// src/system.rs:193). p3-dft 0.5.1 (rev e9d75614) is not vendored in the reference, so this restates
// its published conventions:
//   dft(f)_k = sum_j f_j w^{jk},  w = two_adic_generator(log n)
//   coset_lde_batch(evals, added_bits, shift) = idft -> *shift^j -> zero-pad -> dft
//   Radix2DitParallel returns a bit-reversed view; `.bit_reverse_rows()` is the raw storage in which
//   natural index k lives at row rev(k).
// Field arithmetic is exact, so any correct DFT is bit-identical to the reference's; only the
// storage order is a convention, and it is pinned by the reference's own relational test
// (src/prover.rs:975-999), mirrored in tests/test_oracle_dft.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file.
#pragma once
#include "../multi_stark_b200/host/goldilocks.hpp"
#include "cpu_simd.hpp"
#include <map>
#include <mutex>
#include <memory>
#include <algorithm>

namespace orc {
using namespace msh;

// Twiddle table w_N^i, i < N/2, cached per log N.
inline const std::vector<Fp>& twiddles(unsigned log_n) {
    static std::mutex mu;
    static std::map<unsigned, std::unique_ptr<std::vector<Fp>>> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto& slot = cache[log_n];
    if (!slot) {
        size_t half = log_n == 0 ? 0 : (size_t(1) << (log_n - 1));
        slot = std::make_unique<std::vector<Fp>>(half);
        Fp w = two_adic_generator(log_n), acc = Fp::one();
        for (size_t i = 0; i < half; i++) { (*slot)[i] = acc; acc *= w; }
    }
    return *slot;
}

// One radix-2 decimation-in-frequency layer on rows [base, base+2*half) of a row-major matrix.
static inline void dif_layer_block(Fp* m, size_t width, size_t base, size_t half, const Fp* tw, size_t tw_stride) {
    for (size_t j = 0; j < half; j++) {
        Fp t = tw[j * tw_stride];
        Fp* a = m + (base + j) * width;
        Fp* b = m + (base + j + half) * width;
        if (t.v == 1) simd::dif_butterfly_row_notw((u64*)a, (u64*)b, width);
        else simd::dif_butterfly_row((u64*)a, (u64*)b, t.v, width);
    }
}

// Size-n DIF of a cache-resident block of rows (no threading inside): natural rows in, bit-reversed rows out.
// tw = twiddle table of a transform of size n * tw_scale (w^i, i < n * tw_scale / 2).
static inline void dif_block_inplace(Fp* m, size_t n, size_t width, const Fp* tw, size_t tw_scale) {
    for (size_t blk = n; blk >= 2; blk >>= 1) {
        size_t half = blk >> 1, stride = (n / blk) * tw_scale;
        for (size_t base = 0; base < n; base += blk) dif_layer_block(m, width, base, half, tw, stride);
    }
}

// In-place forward DFT of every column; output left in BIT-REVERSED row order
// (row rev(k) holds dft_k) -- i.e. `dft.dft_batch(m).bit_reverse_rows()` (src/prover.rs:650,716).
// Transforms that do not fit the cache run as TWO passes over memory (the four-step split p3-dft's Radix2DitParallel makes
// with its mid-way bit reversal): n = T * S; pass 1 gathers, for every x < S, the T rows x, x + S, x + 2S, ... into a
// thread-local buffer, transforms them (size T, bit-reversed out), multiplies row r by w_n^{x * rev_T(r)} and scatters them
// back; pass 2 finishes every contiguous block of S rows on its own. One layer at a time over a 100 MB matrix is bound by
// DRAM bandwidth, not by the butterflies.
inline void dft_batch_bitrev_inplace(Fp* m, size_t n, size_t width) {
    if (n <= 1) return;
    unsigned log_n = log2_strict(n);
    const std::vector<Fp>& tw = twiddles(log_n);
    const size_t row_bytes = width * sizeof(Fp), cache = size_t(1) << 19;
    if (n * row_bytes <= cache || log_n < 6) {
        dif_block_inplace(m, n, width, tw.data(), 1);
        return;
    }
    unsigned log_s = 1;
    while (log_s + 1 < log_n && (size_t(2) << log_s) * row_bytes <= cache) log_s++;
    if (log_n - log_s > log_s + 2 && (n >> log_s) * row_bytes > 4 * cache) log_s = log_n / 2;  // very tall: balance the passes
    const size_t S = size_t(1) << log_s, T = n >> log_s;
    const unsigned log_t = log_n - log_s;
    const size_t half_n = n >> 1;
#pragma omp parallel
    {
        std::vector<Fp> buf(T * width);
#pragma omp for schedule(static)
        for (long long xx = 0; xx < (long long)S; xx++) {
            const size_t x = (size_t)xx;
            for (size_t j = 0; j < T; j++) std::copy(m + (j * S + x) * width, m + (j * S + x + 1) * width, buf.data() + j * width);
            dif_block_inplace(buf.data(), T, width, tw.data(), S);  // size-T twiddles are w_n^{i * S}
            for (size_t r = 0; r < T; r++) {
                Fp* dst = m + (r * S + x) * width;
                const Fp* src = buf.data() + r * width;
                size_t e = (x * reverse_bits_len(r, log_t)) & (n - 1);  // w_n^e, w_n^{n/2} = -1
                if (e == 0) { std::copy(src, src + width, dst); continue; }
                Fp t = e < half_n ? tw[e] : -tw[e - half_n];
                simd::scale_row((u64*)dst, (const u64*)src, t.v, width);
            }
        }
    }
    if (S * row_bytes <= 2 * cache) {
#pragma omp parallel for schedule(static)
        for (long long r = 0; r < (long long)T; r++) dif_block_inplace(m + (size_t)r * S * width, S, width, tw.data(), T);
    } else {  // still too large for the cache: recurse (three or more passes), one block at a time with the threads inside
        for (size_t r = 0; r < T; r++) dft_batch_bitrev_inplace(m + r * S * width, S, width);
    }
}

// Out-of-place row permutation by bit reversal.
inline Matrix bit_reverse_rows(const Matrix& in) {
    size_t n = in.height(), w = in.width;
    Matrix out(n, w);
    if (n == 0) return out;
    unsigned log_n = log2_strict(n);
    long long nn = (long long)n;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < nn; i++) {
        size_t r = reverse_bits_len((size_t)i, log_n);
        std::copy(in.row(i), in.row(i) + w, out.row(r));
    }
    return out;
}

// dft_batch in natural order.
inline Matrix dft_batch(Matrix m) {
    dft_batch_bitrev_inplace(m.values.data(), m.height(), m.width);
    return bit_reverse_rows(m);
}

// idft(f)_j = n^{-1} dft(f)_{(n-j) mod n}  (stated at src/prover.rs:620-622).
inline Matrix idft_batch(Matrix m) {
    size_t n = m.height(), w = m.width;
    if (n == 0) return m;
    unsigned log_n = log2_strict(n);
    dft_batch_bitrev_inplace(m.values.data(), n, w);
    Matrix out(n, w);
    Fp n_inv = Fp((u64)n).inverse();
    long long nn = (long long)n;
#pragma omp parallel for schedule(static)
    for (long long j = 0; j < nn; j++) {
        size_t src = reverse_bits_len((n - (size_t)j) & (n - 1), log_n);
        simd::scale_row((u64*)out.row(j), (const u64*)m.row(src), n_inv.v, w);
    }
    return out;
}

// Multiply row j by shift^j.
inline void scale_rows_by_powers(Matrix& m, Fp shift) {
    size_t n = m.height(), w = m.width;
    const size_t CH = 512;
    long long nch = (long long)((n + CH - 1) / CH);
#pragma omp parallel for schedule(static)
    for (long long ch = 0; ch < nch; ch++) {
        Fp weight = shift.pow((u64)ch * CH);
        for (size_t r = (size_t)ch * CH; r < std::min(n, (size_t)(ch + 1) * CH); r++) {
            Fp* row = m.row(r);
            simd::scale_row_inplace((u64*)row, weight.v, w);
            weight *= shift;
        }
    }
}

// coset_dft_batch(coeffs, shift): evaluations at shift * w^k, natural order.
inline Matrix coset_dft_batch(Matrix coeffs, Fp shift) {
    scale_rows_by_powers(coeffs, shift);
    return dft_batch(std::move(coeffs));
}
// coset_idft_batch(evals, shift): coefficients of the polynomial with those evaluations on shift*H.
inline Matrix coset_idft_batch(Matrix evals, Fp shift) {
    Matrix c = idft_batch(std::move(evals));
    scale_rows_by_powers(c, shift.inverse());
    return c;
}

// `lde_from_shifted_coefficients` (src/prover.rs:709-717): zero-pad to n << added_bits, one DFT,
// raw bit-reversed storage.
inline Matrix lde_from_shifted_coefficients(Matrix coeffs, unsigned added_bits) {
    size_t n = coeffs.height(), w = coeffs.width;
    coeffs.values.resize((n << added_bits) * w, Fp::zero());
    dft_batch_bitrev_inplace(coeffs.values.data(), n << added_bits, w);
    return coeffs;
}

// What `TwoAdicFriPcs::commit` stores per matrix (src/prover.rs:681-692):
// coset_lde_batch(evals, log_blowup, shift).bit_reverse_rows(): stored[i] = P(shift * w_{nB}^{rev(i)}).
// = idft -> * shift^j -> zero-pad -> dft, with the inverse's reordering, the 1/n and the shift powers written straight into
// the zero-padded matrix (one pass, no intermediate copies).
inline Matrix coset_lde_batch_bitrev(Matrix evals, unsigned added_bits, Fp shift) {
    size_t n = evals.height(), w = evals.width;
    if (n == 0) return evals;
    unsigned log_n = log2_strict(n);
    dft_batch_bitrev_inplace(evals.values.data(), n, w);
    Matrix out(n << added_bits, w);  // rows >= n stay zero
    Fp n_inv = Fp((u64)n).inverse();
    const size_t CH = 512;
    long long nch = (long long)((n + CH - 1) / CH);
#pragma omp parallel for schedule(static)
    for (long long ch = 0; ch < nch; ch++) {
        Fp weight = shift.pow((u64)ch * CH) * n_inv;  // idft(f)_j = n^-1 dft(f)_{(n-j) mod n}, then * shift^j
        for (size_t j = (size_t)ch * CH; j < std::min(n, (size_t)(ch + 1) * CH); j++) {
            size_t src = reverse_bits_len((n - j) & (n - 1), log_n);
            simd::scale_row((u64*)out.row(j), (const u64*)evals.row(src), weight.v, w);
            weight *= shift;
        }
    }
    dft_batch_bitrev_inplace(out.values.data(), n << added_bits, w);
    return out;
}

// `shifted_quotient_slices` (src/prover.rs:631-679): from the quotient's evaluations on the coset
// GENERATOR*H_{nq} (nq rows, d columns, natural order) to the n x (q*d) matrix of slice coefficients
// with the committed LDE's GENERATOR^r row pre-scale folded in:
//   out[r][k*d+c] = S[rev((N-(k*n+r)) mod N)][c] * N^{-1} * GENERATOR^{-k*n},  S = raw DFT storage.
inline Matrix shifted_quotient_slices(Matrix quotient_evals, size_t quotient_degree) {
    size_t d = quotient_evals.width, big = quotient_evals.height();
    unsigned log_big = log2_strict(big);
    size_t n = big / quotient_degree, width = quotient_degree * d;
    dft_batch_bitrev_inplace(quotient_evals.values.data(), big, d);
    const Matrix& storage = quotient_evals;
    Fp n_inv = Fp((u64)big).inverse();
    Fp step = Fp(GL_GENERATOR).pow((u64)n).inverse();
    std::vector<Fp> weights(quotient_degree);
    Fp acc = Fp::one();
    for (size_t k = 0; k < quotient_degree; k++) { weights[k] = acc * n_inv; acc *= step; }
    Matrix out(n, width);
    long long nn = (long long)n;
#pragma omp parallel for schedule(static)
    for (long long row = 0; row < nn; row++) {
        for (size_t k = 0; k < quotient_degree; k++) {
            size_t j = k * n + (size_t)row;
            size_t src = reverse_bits_len((big - j) & (big - 1), log_big);
            for (size_t c = 0; c < d; c++) out.row(row)[k * d + c] = storage.row(src)[c] * weights[k];
        }
    }
    return out;
}

}  // namespace orc
