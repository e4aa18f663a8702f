// Goldilocks NTT / coset low-degree extension kernels (sm_100a).
//
// Replaces, on the device, what the reference obtains from p3-dft `Radix2DitParallel`
// (src/types.rs:200; call sites src/prover.rs:440,650,716) and the `coset_lde_batch(...)
// .bit_reverse_rows()` inside `TwoAdicFriPcs::commit` (src/prover.rs:350,419; src/system.rs:193).
//
// Data layout: row-major n x w matrices of canonical u64, exactly the reference's RowMajorMatrix.
// A transform of size n = 2^L over the rows is split into passes of at most 10 bits. Each pass
// stages a tile of T (= 2^tb) points x up to 16 adjacent u64 (128 B of one row, or whole rows of
// several adjacent blocks) in shared memory, runs the tb butterfly stages as register-resident
// radix-16/8/4/2 rounds, applies the inter-pass twiddle and writes back:
//
//  * k_ntt_strided: decimation-in-frequency pass over bits [lo, lo+tb) of the row index, natural
//    order in, bit-reversed order out, in place (the 4-step decomposition with the transpose absorbed
//    into index arithmetic). A chain of these is a DFT whose output is stored bit-reversed, which is
//    the storage the reference gets from `dft_batch(..).bit_reverse_rows()`.
//  * k_ntt_block: first pass of the coset evaluation. It reads the bit-reversed coefficients that the
//    inverse transform left behind (contiguous blocks, fully coalesced), multiplies row j by
//    shift_s^j / n, runs a decimation-in-time tile (bit-reversed in, natural out) and scatters whole
//    128-byte row segments to their bit-reversed block. The zero-padded forward transform of size
//    n*B is never materialised: output block rev(s) of n rows is the size-n coset transform with
//    shift * w_{nB}^s, so the LDE reads n rows and writes n*B.
#include "internal.hpp"

namespace msg {

constexpr int kPitch = 17;        // u64 per tile row (16 data + 1 pad: conflict-free column writes)
constexpr int kMaxXt = 16;        // tile columns
constexpr int kMaxLogT = 10;      // tile points
constexpr int kTwSmallLog = 10;

// ------------------------------------------------------------------------------------------------
// Register radix rounds over a shared-memory tile [T][kPitch].
// ------------------------------------------------------------------------------------------------
template <int RB, bool DIT>
__device__ __forceinline__ void radix_round(u64* tile, const u64* twl, u32 log_t, u32 s0, u32 ncols, u32 tid,
                                            u32 nthr) {
    constexpr int R = 1 << RB;
    const u32 stride = 1u << s0;
    const u32 items = ((1u << log_t) >> RB) * ncols;
    for (u32 item = tid; item < items; item += nthr) {
        u32 q = item % ncols, g = item / ncols;
        u32 off = g & (stride - 1), blk = g >> s0;
        u32 row0 = (blk << (s0 + RB)) + off;
        u64* base = tile + (size_t)row0 * kPitch + q;
        u64 v[R];
#pragma unroll
        for (int m = 0; m < R; m++) v[m] = base[(size_t)(m << s0) * kPitch];
#pragma unroll
        for (int s = 0; s < RB; s++) {
            const int lb = DIT ? s : RB - 1 - s;
            const int half = 1 << lb;
            const u32 shift = log_t - 1 - s0 - lb;
#pragma unroll
            for (int el = 0; el < half; el++) {
                u64 t = twl[(((u32)el << s0) + off) << shift];
#pragma unroll
                for (int m0 = 0; m0 < R; m0 += 2 * half) {
                    int i = m0 + el, j = i + half;
                    if (DIT) {
                        u64 x = gl::mul(v[j], t), u = v[i];
                        v[i] = gl::add(u, x);
                        v[j] = gl::sub(u, x);
                    } else {
                        u64 u = v[i], x = v[j];
                        v[i] = gl::add(u, x);
                        v[j] = gl::mul(gl::sub(u, x), t);
                    }
                }
            }
        }
#pragma unroll
        for (int m = 0; m < R; m++) base[(size_t)(m << s0) * kPitch] = v[m];
    }
}

template <bool DIT>
__device__ __forceinline__ void radix_dispatch(int rb, u64* tile, const u64* twl, u32 log_t, u32 s0, u32 ncols,
                                               u32 tid, u32 nthr) {
    switch (rb) {
        case 4: radix_round<4, DIT>(tile, twl, log_t, s0, ncols, tid, nthr); break;
        case 3: radix_round<3, DIT>(tile, twl, log_t, s0, ncols, tid, nthr); break;
        case 2: radix_round<2, DIT>(tile, twl, log_t, s0, ncols, tid, nthr); break;
        default: radix_round<1, DIT>(tile, twl, log_t, s0, ncols, tid, nthr); break;
    }
}

// natural order in -> bit-reversed order out (rows of the tile)
__device__ __forceinline__ void tile_dif(u64* tile, const u64* twl, u32 log_t, u32 ncols) {
    int hi = (int)log_t;
    while (hi > 0) {
        int rb = hi >= 4 ? 4 : hi;
        if (hi > 4 && hi < 8) rb = (hi + 1) / 2;  // 7 -> 4+3, 6 -> 3+3, 5 -> 3+2
        radix_dispatch<false>(rb, tile, twl, log_t, (u32)(hi - rb), ncols, threadIdx.x, blockDim.x);
        __syncthreads();
        hi -= rb;
    }
}
// bit-reversed order in -> natural order out
__device__ __forceinline__ void tile_dit(u64* tile, const u64* twl, u32 log_t, u32 ncols) {
    int lo = 0;
    while (lo < (int)log_t) {
        int rem = (int)log_t - lo;
        int rb = rem >= 4 ? 4 : rem;
        if (rem > 4 && rem < 8) rb = (rem + 1) / 2;
        radix_dispatch<true>(rb, tile, twl, log_t, (u32)lo, ncols, threadIdx.x, blockDim.x);
        __syncthreads();
        lo += rb;
    }
}

__device__ __forceinline__ void load_local_twiddles(u64* twl, const u64* tw_small, u32 log_t) {
    u32 half = (1u << log_t) >> 1;
    for (u32 i = threadIdx.x; i < half; i += blockDim.x) twl[i] = __ldg(tw_small + ((size_t)i << (kTwSmallLog - log_t)));
}

// ------------------------------------------------------------------------------------------------
// Strided decimation-in-frequency pass.
// Element (a, j, x) of the pass lives at ((a*T + j)*I + x); the pass transforms j for every (a, x),
// then multiplies the output at tile row r (frequency k1 = rev(r)) by w_M^{(x / w) * k1}, M = T*I/w.
// ------------------------------------------------------------------------------------------------
struct StridedParams {
    const u64* src;
    u64* dst;
    const u64* tw_small;
    gl::PowTable tw;
    u64 I;        // inner elements per point: S * w
    u64 a_total;  // number of (independent) outer blocks
    u32 log_t, w, xt;
    u32 nchunks;  // wide mode: column chunks per outer block; 0 selects narrow mode
    u32 aa;       // narrow mode (I <= xt): outer blocks per tile
    int has_tw;
};

__global__ void __launch_bounds__(1024) k_ntt_strided(StridedParams p) {
    extern __shared__ u64 smem[];
    const u32 T = 1u << p.log_t;
    u64* tile = smem;
    u64* twl = smem + (size_t)T * kPitch;
    const u32 tid = threadIdx.x, nthr = blockDim.x;
    load_local_twiddles(twl, p.tw_small, p.log_t);

    u32 ncols;
    u64 base;  // element offset of (row 0, col 0)
    u64 x0 = 0;
    if (p.nchunks) {
        u64 a = blockIdx.x / p.nchunks;
        x0 = (u64)(blockIdx.x % p.nchunks) * p.xt;
        ncols = (u32)min((u64)p.xt, p.I - x0);
        base = a * T * p.I + x0;
        const u32 total = T * ncols;
        for (u32 e = tid; e < total; e += nthr) {
            u32 q = e % ncols, j = e / ncols;
            tile[(size_t)j * kPitch + q] = p.src[base + (u64)j * p.I + q];
        }
    } else {
        u64 a0 = (u64)blockIdx.x * p.aa;
        u32 na = (u32)min((u64)p.aa, p.a_total - a0);
        u32 I = (u32)p.I;
        ncols = na * I;
        base = a0 * T * I;
        const u32 per_a = T * I, total = na * per_a;
        for (u32 e = tid; e < total; e += nthr) {
            u32 aq = e / per_a, rem = e % per_a;
            u32 j = rem / I, x = rem % I;
            tile[(size_t)j * kPitch + aq * I + x] = p.src[base + e];
        }
    }
    __syncthreads();
    tile_dif(tile, twl, p.log_t, ncols);

    if (p.nchunks) {
        const u32 total = T * ncols;
        for (u32 e = tid; e < total; e += nthr) {
            u32 q = e % ncols, j = e / ncols;
            u64 v = tile[(size_t)j * kPitch + q];
            if (p.has_tw) {
                u64 b = (x0 + q) / p.w;
                u64 k1 = gl::rev_bits(j, p.log_t);
                v = gl::mul(v, gl::pow_lookup(p.tw, b * k1));
            }
            p.dst[base + (u64)j * p.I + q] = v;
        }
    } else {
        u32 I = (u32)p.I;
        const u32 per_a = T * I, total = (ncols / I) * per_a;
        for (u32 e = tid; e < total; e += nthr) {
            u32 aq = e / per_a, rem = e % per_a;
            u32 j = rem / I, x = rem % I;
            u64 v = tile[(size_t)j * kPitch + aq * I + x];
            if (p.has_tw) {
                u64 b = x / p.w;
                u64 k1 = gl::rev_bits(j, p.log_t);
                v = gl::mul(v, gl::pow_lookup(p.tw, b * k1));
            }
            p.dst[base + e] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Block pass: first pass of the coset evaluations from bit-reversed coefficients.
// src position p = (blk, t): coefficient j = rev_T(t) * S + rev_{L-tb}(blk). For low index b, the tile
// reads block rev(b), scales by scale[beta](j), runs a DIT tile (row s = frequency k1 = s), multiplies
// by w_n^{b*k1} and stores row s at destination row rev_T(s) * S + b of output block beta.
// ------------------------------------------------------------------------------------------------
struct BlockParams {
    const u64* src;
    u64* dst;
    const u64* tw_small;
    const gl::PowTable* scale;  // [B] device array, indexed by output block
    gl::PowTable tw;
    u64 S;                // n / T
    u64 out_block_elems;  // n * w
    u32 log_n, log_t, w, xt;
    u32 nchunks;  // w > xt: column chunks per b; 0: whole rows, nbq low indices per tile
    u32 nbq;
    int has_tw;
};

__global__ void __launch_bounds__(1024) k_ntt_block(BlockParams p) {
    extern __shared__ u64 smem[];
    const u32 T = 1u << p.log_t;
    u64* tile = smem;
    u64* twl = smem + (size_t)T * kPitch;
    const u32 tid = threadIdx.x, nthr = blockDim.x;
    load_local_twiddles(twl, p.tw_small, p.log_t);
    const gl::PowTable sc = p.scale[blockIdx.y];
    const u32 hi_bits = p.log_n - p.log_t;
    const u32 w = p.w;
    u32 ncols;
    u64 b0;
    u32 c0 = 0;
    if (p.nchunks == 0) {
        b0 = (u64)blockIdx.x * p.nbq;
        u32 nb = (u32)min((u64)p.nbq, p.S - b0);
        ncols = nb * w;
        const u32 per_b = T * w, total = nb * per_b;
        for (u32 e = tid; e < total; e += nthr) {
            u32 bq = e / per_b, rem = e % per_b;
            u32 t = rem / w, c = rem % w;
            u64 b = b0 + bq;
            u64 blk = gl::rev_bits((u32)b, hi_bits);
            u64 v = p.src[(blk * T + t) * w + c];
            u64 j = (u64)gl::rev_bits(t, p.log_t) * p.S + b;
            tile[(size_t)t * kPitch + bq * w + c] = gl::mul(v, gl::pow_lookup(sc, j));
        }
    } else {
        b0 = blockIdx.x / p.nchunks;
        c0 = (blockIdx.x % p.nchunks) * p.xt;
        ncols = min(p.xt, w - c0);
        u64 blk = gl::rev_bits((u32)b0, hi_bits);
        const u32 total = T * ncols;
        for (u32 e = tid; e < total; e += nthr) {
            u32 q = e % ncols, t = e / ncols;
            u64 v = p.src[(blk * T + t) * w + c0 + q];
            u64 j = (u64)gl::rev_bits(t, p.log_t) * p.S + b0;
            tile[(size_t)t * kPitch + q] = gl::mul(v, gl::pow_lookup(sc, j));
        }
    }
    __syncthreads();
    tile_dit(tile, twl, p.log_t, ncols);

    u64* out = p.dst + (u64)blockIdx.y * p.out_block_elems;
    const u32 total = T * ncols;
    for (u32 e = tid; e < total; e += nthr) {
        u32 q = e % ncols, s = e / ncols;
        u64 v = tile[(size_t)s * kPitch + q];
        u64 b = p.nchunks ? b0 : b0 + q / w;
        if (p.has_tw) v = gl::mul(v, gl::pow_lookup(p.tw, b * (u64)s));
        u64 drow = (u64)gl::rev_bits(s, p.log_t) * p.S;
        // (drow + b0) * w + c0 + q addresses (b, c) because q = bq * w + c in whole-row mode
        out[(drow + b0) * w + c0 + q] = v;
    }
}

// dst[beta][j][c] = src[j][c] * tab[beta](j)
__global__ void k_scale_rows(const u64* src, u64* dst, u64 n, u64 w, const gl::PowTable* tabs, u32 nb) {
    u64 total = n * w;
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (u64)gridDim.x * blockDim.x) {
        u64 j = e / w;
        u64 v = src[e];
        for (u32 beta = 0; beta < nb; beta++) dst[(u64)beta * total + e] = gl::mul(v, gl::pow_lookup(tabs[beta], j));
    }
}

// dst[r][c] = src[rev(r)][c] * scalar
__global__ void k_bitrev_rows_scale(const u64* src, u64* dst, u64 n, u64 w, u32 log_n, u64 scalar) {
    u64 total = n * w;
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (u64)gridDim.x * blockDim.x) {
        u64 r = e / w, c = e % w;
        u64 v = src[(u64)gl::rev_bits((u32)r, log_n) * w + c];
        dst[e] = scalar == 1 ? v : gl::mul(v, scalar);
    }
}

// ------------------------------------------------------------------------------------------------
// Host drivers
// ------------------------------------------------------------------------------------------------
static std::vector<u32> plan_passes(u32 log_n) {
    std::vector<u32> tb;
    if (log_n == 0) return tb;
    u32 np = (log_n + kMaxLogT - 1) / kMaxLogT;
    u32 base = log_n / np, rem = log_n % np;
    for (u32 i = 0; i < np; i++) tb.push_back(base + (i < rem ? 1 : 0));
    return tb;
}

static size_t tile_smem(u32 log_t) { return ((size_t)(1u << log_t) * kPitch + ((1u << log_t) >> 1)) * sizeof(u64); }
static u32 tile_threads(u32 log_t, u32 ncols) {
    u32 t = ((1u << log_t) * ncols) / 16;
    if (t < 64) t = 64;
    if (t > 1024) t = 1024;
    return (t + 31) / 32 * 32;
}

static void ensure_smem_attr() {
    static bool done = false;
    if (done) return;
    MSG_CUDA(cudaFuncSetAttribute(k_ntt_strided, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem(kMaxLogT)));
    MSG_CUDA(cudaFuncSetAttribute(k_ntt_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem(kMaxLogT)));
    done = true;
}

// One DIF pass over bits [lo_bit, lo_bit + tb) of a size-2^log_m transform, for `count` transforms
// stored back to back (count * 2^log_m rows in total).
static void launch_strided(Ctx& c, const u64* src, u64* dst, u32 log_m, u64 count, u64 w, u32 lo_bit, u32 tb, bool inverse) {
    ensure_smem_attr();
    StridedParams p{};
    p.src = src;
    p.dst = dst;
    p.tw_small = c.tw_small[inverse ? 1 : 0];
    p.log_t = tb;
    p.w = (u32)w;
    p.xt = kMaxXt;
    u64 S = 1ull << lo_bit;
    p.I = S * w;
    p.a_total = count << (log_m - lo_bit - tb);
    p.has_tw = lo_bit > 0;
    if (p.has_tw) {
        u32 log_big = lo_bit + tb;
        msh::Fp g = msh::two_adic_generator(log_big);
        if (inverse) g = g.inverse();
        p.tw = c.pow_table(g.v, 1, log_big).view();
    }
    u64 blocks;
    u32 ncols_max;
    if (p.I > (u64)p.xt) {
        p.nchunks = (u32)((p.I + p.xt - 1) / p.xt);
        blocks = p.a_total * p.nchunks;
        ncols_max = p.xt;
    } else {
        p.nchunks = 0;
        p.aa = (u32)(p.xt / p.I);
        blocks = (p.a_total + p.aa - 1) / p.aa;
        ncols_max = (u32)std::min<u64>(p.aa, p.a_total) * (u32)p.I;
    }
    MSG_REQUIRE(blocks < (1ull << 31), "ntt: grid too large");
    {
        KLaunch kl(c, "k_ntt_strided");
        k_ntt_strided<<<(unsigned)blocks, tile_threads(tb, ncols_max), tile_smem(tb), c.stream>>>(p);
    }
    MSG_CUDA(cudaGetLastError());
}

void ntt_dft_bitrev(Ctx& c, const u64* src, u64* dst, u64 n, u64 w, bool inverse, u64 batch) {
    MSG_REQUIRE(is_pow2(n), "ntt: height must be a power of two");
    if (w == 0 || batch == 0) return;
    u32 log_n = ilog2(n);
    MSG_REQUIRE(log_n <= msh::GL_TWO_ADICITY, "ntt: height exceeds the two-adicity of the field");
    auto plan = plan_passes(log_n);
    if (plan.empty()) {
        if (src != dst) MSG_CUDA(cudaMemcpyAsync(dst, src, batch * n * w * 8, cudaMemcpyDeviceToDevice, c.stream));
        return;
    }
    u32 hi = log_n;
    const u64* cur = src;
    for (u32 tb : plan) {
        launch_strided(c, cur, dst, log_n, batch, w, hi - tb, tb, inverse);
        cur = dst;
        hi -= tb;
    }
}

void ntt_coset_lde(Ctx& c, const u64* src, u64* dst, u64* tmp, u64 n, u64 w, u32 added_bits, u64 shift) {
    MSG_REQUIRE(is_pow2(n), "lde: height must be a power of two");
    if (w == 0) return;
    u32 log_n = ilog2(n);
    MSG_REQUIRE(log_n + added_bits <= msh::GL_TWO_ADICITY, "lde: extended height exceeds the two-adicity of the field");
    ensure_smem_attr();
    // 1. inverse transform (unnormalised): coefficients in bit-reversed order
    ntt_dft_bitrev(c, src, tmp, n, w, true);
    // 2. first forward pass per coset from the bit-reversed coefficients
    auto plan = plan_passes(log_n);
    u32 tb = plan.empty() ? 0 : plan.back();
    BlockParams p{};
    p.src = tmp;
    p.dst = dst;
    p.tw_small = c.tw_small[0];
    p.scale = c.coset_tables(log_n, added_bits, shift);
    p.log_n = log_n;
    p.log_t = tb;
    p.w = (u32)w;
    p.xt = kMaxXt;
    p.S = n >> tb;
    p.out_block_elems = n * w;
    p.has_tw = p.S > 1;
    if (p.has_tw) p.tw = c.pow_table(msh::two_adic_generator(log_n).v, 1, log_n).view();
    u64 blocks;
    u32 ncols_max;
    if (w <= (u64)p.xt) {
        p.nchunks = 0;
        p.nbq = (u32)(p.xt / w);
        blocks = (p.S + p.nbq - 1) / p.nbq;
        ncols_max = (u32)std::min<u64>(p.nbq, p.S) * (u32)w;
    } else {
        p.nchunks = (u32)((w + p.xt - 1) / p.xt);
        blocks = p.S * p.nchunks;
        ncols_max = p.xt;
    }
    MSG_REQUIRE(blocks < (1ull << 31), "lde: grid too large");
    dim3 grid((unsigned)blocks, 1u << added_bits);
    {
        KLaunch kl(c, "k_ntt_block");
        k_ntt_block<<<grid, tile_threads(tb, ncols_max), tile_smem(tb), c.stream>>>(p);
    }
    MSG_CUDA(cudaGetLastError());
    // 3. remaining forward passes: every block rev_T(k1) of S rows is an independent size-S DFT
    u32 log_s = log_n - tb;
    if (log_s > 0) {
        auto rest = plan_passes(log_s);
        u32 hi = log_s;
        u64 count = (n >> log_s) << added_bits;
        for (u32 t2 : rest) {
            launch_strided(c, dst, dst, log_s, count, w, hi - t2, t2, false);
            hi -= t2;
        }
    }
}

void ntt_lde_from_coeffs(Ctx& c, const u64* src, u64* dst, u64 n, u64 w, u32 added_bits) {
    MSG_REQUIRE(is_pow2(n), "lde: height must be a power of two");
    if (w == 0) return;
    u32 log_n = ilog2(n);
    MSG_REQUIRE(log_n + added_bits <= msh::GL_TWO_ADICITY, "lde: extended height exceeds the two-adicity of the field");
    const gl::PowTable* tabs = c.lde_coeff_tables(log_n, added_bits);
    u64 total = n * w;
    unsigned blocks = (unsigned)std::min<u64>((total + 255) / 256, (u64)c.sm_count * 16);
    {
        KLaunch kl(c, "k_scale_rows");
        k_scale_rows<<<blocks, 256, 0, c.stream>>>(src, dst, n, w, tabs, 1u << added_bits);
    }
    MSG_CUDA(cudaGetLastError());
    ntt_dft_bitrev(c, dst, dst, n, w, false, 1ull << added_bits);
}

void ntt_bit_reverse_rows(Ctx& c, const u64* src, u64* dst, u64 n, u64 w) {
    if (n * w == 0) return;
    u64 total = n * w;
    unsigned blocks = (unsigned)std::min<u64>((total + 255) / 256, (u64)c.sm_count * 16);
    {
        KLaunch kl(c, "k_bitrev_rows_scale");
        k_bitrev_rows_scale<<<blocks, 256, 0, c.stream>>>(src, dst, n, w, ilog2(n), 1);
    }
    MSG_CUDA(cudaGetLastError());
}

void ntt_idft_natural(Ctx& c, const u64* src, u64* dst, u64* tmp, u64 n, u64 w) {
    if (n * w == 0) return;
    ntt_dft_bitrev(c, src, tmp, n, w, true);
    u64 total = n * w;
    unsigned blocks = (unsigned)std::min<u64>((total + 255) / 256, (u64)c.sm_count * 16);
    {
        KLaunch kl(c, "k_bitrev_rows_scale");
        k_bitrev_rows_scale<<<blocks, 256, 0, c.stream>>>(tmp, dst, n, w, ilog2(n), msh::Fp((msh::u64)n).inverse().v);
    }
    MSG_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Tables
// ------------------------------------------------------------------------------------------------
DevPow Ctx::pow_table(u64 g, u64 cst, u32 bits) {
    auto key = std::make_tuple(g, cst, bits);
    auto it = pow_cache.find(key);
    if (it != pow_cache.end()) return it->second;
    DevPow t;
    t.h1 = (bits + 1) / 2;
    size_t nlo = (size_t)1 << t.h1, nhi = (size_t)1 << (bits - t.h1);
    std::vector<u64> lo(nlo), hi(nhi);
    msh::Fp G(g), acc(cst);
    for (size_t i = 0; i < nlo; i++) { lo[i] = acc.v; acc *= G; }
    msh::Fp Gh = G.exp_power_of_2(t.h1);
    acc = msh::Fp::one();
    for (size_t i = 0; i < nhi; i++) { hi[i] = acc.v; acc *= Gh; }
    MSG_CUDA(cudaMalloc(&t.lo, nlo * 8));
    MSG_CUDA(cudaMalloc(&t.hi, nhi * 8));
    owned.push_back(t.lo);
    owned.push_back(t.hi);
    MSG_CUDA(cudaMemcpyAsync(t.lo, lo.data(), nlo * 8, cudaMemcpyHostToDevice, stream));
    MSG_CUDA(cudaMemcpyAsync(t.hi, hi.data(), nhi * 8, cudaMemcpyHostToDevice, stream));
    MSG_CUDA(cudaStreamSynchronize(stream));  // the host vectors die here
    pow_cache[key] = t;
    return t;
}

static const gl::PowTable* upload_views(Ctx& c, const std::vector<gl::PowTable>& v) {
    gl::PowTable* d;
    MSG_CUDA(cudaMalloc(&d, v.size() * sizeof(gl::PowTable)));
    c.owned.push_back(d);
    MSG_CUDA(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(gl::PowTable), cudaMemcpyHostToDevice, c.stream));
    MSG_CUDA(cudaStreamSynchronize(c.stream));
    return d;
}

// scale[beta](j) = (shift * w_{nB}^{rev_b(beta)})^j / n
const gl::PowTable* Ctx::coset_tables(u32 log_n, u32 added_bits, u64 shift) {
    auto key = std::make_tuple(log_n, added_bits, shift);
    auto it = coset_cache.find(key);
    if (it != coset_cache.end()) return it->second;
    u32 B = 1u << added_bits;
    msh::Fp wN = msh::two_adic_generator(log_n + added_bits);
    msh::Fp ninv = msh::Fp((msh::u64)1 << log_n).inverse();
    std::vector<gl::PowTable> views(B);
    for (u32 beta = 0; beta < B; beta++) {
        u32 s = (u32)msh::reverse_bits_len(beta, added_bits);
        msh::Fp g = msh::Fp(shift) * wN.pow(s);
        views[beta] = pow_table(g.v, ninv.v, log_n).view();
    }
    gl::PowTable* d = const_cast<gl::PowTable*>(upload_views(*this, views));
    coset_cache[key] = d;
    return d;
}

// tab[beta](j) = w_{nB}^{rev_b(beta) * j}
const gl::PowTable* Ctx::lde_coeff_tables(u32 log_n, u32 added_bits) {
    auto key = std::make_tuple(log_n, added_bits | 0x80000000u, (u64)0);
    auto it = coset_cache.find(key);
    if (it != coset_cache.end()) return it->second;
    u32 B = 1u << added_bits;
    msh::Fp wN = msh::two_adic_generator(log_n + added_bits);
    std::vector<gl::PowTable> views(B);
    for (u32 beta = 0; beta < B; beta++) {
        u32 s = (u32)msh::reverse_bits_len(beta, added_bits);
        views[beta] = pow_table(wN.pow(s).v, 1, log_n).view();
    }
    gl::PowTable* d = const_cast<gl::PowTable*>(upload_views(*this, views));
    coset_cache[key] = d;
    return d;
}

void ctx_init_tables(Ctx& c) {
    const size_t half = (size_t)1 << (kTwSmallLog - 1);
    std::vector<u64> f(half), inv(half);
    msh::Fp w = msh::two_adic_generator(kTwSmallLog), wi = w.inverse();
    msh::Fp a = msh::Fp::one(), b = msh::Fp::one();
    for (size_t i = 0; i < half; i++) { f[i] = a.v; inv[i] = b.v; a *= w; b *= wi; }
    for (int d = 0; d < 2; d++) {
        MSG_CUDA(cudaMalloc(&c.tw_small[d], half * 8));
        c.owned.push_back(c.tw_small[d]);
    }
    MSG_CUDA(cudaMemcpyAsync(c.tw_small[0], f.data(), half * 8, cudaMemcpyHostToDevice, c.stream));
    MSG_CUDA(cudaMemcpyAsync(c.tw_small[1], inv.data(), half * 8, cudaMemcpyHostToDevice, c.stream));
    MSG_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace msg
