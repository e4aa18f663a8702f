// ORACLE (test infrastructure, not product code): vectorised inner loops of the CPU restatement, so that the CPU arm of
// bench.py is a fair stand-in for the reference's `--features parallel` build, which runs Plonky3's packed Goldilocks
// arithmetic (AVX2 / AVX-512 `PackedGoldilocks`, enabled by `target-cpu=native`, /root/reference/.cargo/config.toml:2 and the
// `Packing` types at src/types.rs:25-27) and the SIMD BLAKE3 of the `blake3` crate 1.8.5.
//
// Written as plain C over lane arrays; GCC vectorises the loops for every clone of `ORC_SIMD` (x86-64-v4 = AVX-512F/BW/DQ/VL,
// x86-64-v3 = AVX2, baseline) and the loader picks the best clone for the machine it runs on: the library is built in the
// build container and shipped to the GPU box, so `-march=native` at build time would be wrong.
//   * Goldilocks: 64 x 64 -> 128 from four 32 x 32 -> 64 products (vpmuludq), reduction with 2^64 = 2^32 - 1, 2^96 = -1;
//     8 butterflies per AVX-512 instruction stream, vectorised ACROSS THE COLUMNS of a row pair (one twiddle per row pair,
//     exactly the loop structure of p3-dft's `Radix2DitParallel` over `RowMajorMatrix` rows).
//   * BLAKE3: kLanes independent messages per pass, state transposed (word i of every lane in one vector), as the blake3
//     crate's hash_many does; used for the Merkle leaves (one row per lane) and node layers (one node per lane).
// Results are bit-identical to the scalar code (checked by tests/test_oracle_simd.py against the scalar paths and the
// official blake3 package).
#pragma once
#include "../multi_stark_b200/host/goldilocks.hpp"
#include "../multi_stark_b200/host/blake3_host.hpp"

namespace orc {
namespace simd {
using msh::u64;
using msh::u32;

static inline u64 gl_mul(u64 a, u64 b) {
    u64 a0 = (u32)a, a1 = a >> 32, b0 = (u32)b, b1 = b >> 32;
    u64 ll = a0 * b0, lh = a0 * b1, hl = a1 * b0, hh = a1 * b1;
    u64 mid = lh + (ll >> 32);
    u64 mid2 = hl + (u32)mid;
    u64 hi = hh + (mid >> 32) + (mid2 >> 32);
    u64 lo = (u32)ll | (mid2 << 32);
    u64 hi_hi = hi >> 32, hi_lo = (u32)hi;
    u64 t0 = lo - hi_hi;
    t0 -= (lo < hi_hi) ? msh::GL_EPS : 0;
    u64 t1 = (hi_lo << 32) - hi_lo;
    u64 r = t0 + t1;
    r += (r < t1) ? msh::GL_EPS : 0;
    r -= (r >= msh::GL_P) ? msh::GL_P : 0;
    return r;
}
static inline u64 gl_add(u64 x, u64 y) {
    u64 s = x + y;
    s -= (s < x || s >= msh::GL_P) ? msh::GL_P : 0;
    return s;
}
static inline u64 gl_sub(u64 x, u64 y) {
    u64 d = x - y;
    d += (x < y) ? msh::GL_P : 0;
    return d;
}

// a[c], b[c] <- a[c] + b[c], (a[c] - b[c]) * t     (one decimation-in-frequency butterfly per column)
void dif_butterfly_row(u64* __restrict a, u64* __restrict b, u64 t, size_t w);
// the same without the multiplication (twiddle 1)
void dif_butterfly_row_notw(u64* __restrict a, u64* __restrict b, size_t w);
// dst[c] = src[c] * s
void scale_row(u64* __restrict dst, const u64* __restrict src, u64 s, size_t w);
// row[c] *= s
void scale_row_inplace(u64* row, u64 s, size_t w);

// ---- BLAKE3, kLanes messages at a time --------------------------------------------------------------------------------
constexpr int kLanes = 16;
typedef u32 lane_t[kLanes] __attribute__((aligned(64)));
// cv[i][lane] (in/out) <- compress(cv, m, counter_lo[lane], block_len[lane], flags[lane]) for every lane
void b3_compress_lanes(lane_t* cv, const lane_t* m, const u32* block_len, const u32* flags, const u32* counter_lo);

// BLAKE3 of kLanes messages of the SAME length `len` <= 1024 bytes (one chunk): msgs[l] points to lane l's bytes (lanes >= n are
// ignored). out[l] receives the 32-byte digest.
inline void b3_hash_lanes_one_chunk(const uint8_t* const* msgs, int n, size_t len, msh::Digest* out) {
    lane_t cv[8], m[16];
    u32 bl[kLanes], fl[kLanes], ctr[kLanes];
    for (int i = 0; i < 8; i++)
        for (int l = 0; l < kLanes; l++) cv[i][l] = msh::b3::IV[i];
    for (int l = 0; l < kLanes; l++) ctr[l] = 0;
    const size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
    for (size_t b = 0; b < nblocks; b++) {
        const size_t off = b * 64, blen = len - off < 64 ? len - off : 64;
        for (int l = 0; l < kLanes; l++) {
            uint8_t buf[64] = {0};
            const uint8_t* src = msgs[l < n ? l : 0] + off;
            if (blen == 64) {
                for (int k = 0; k < 16; k++) { u32 w; memcpy(&w, src + 4 * k, 4); m[k][l] = w; }  // little-endian host
            } else {
                memcpy(buf, src, blen);
                for (int k = 0; k < 16; k++) { u32 w; memcpy(&w, buf + 4 * k, 4); m[k][l] = w; }
            }
            bl[l] = (u32)blen;
            fl[l] = (b == 0 ? msh::b3::CHUNK_START : 0) | (b + 1 == nblocks ? (msh::b3::CHUNK_END | msh::b3::ROOT) : 0);
        }
        b3_compress_lanes(cv, m, bl, fl, ctr);
    }
    for (int l = 0; l < n; l++)
        for (int i = 0; i < 8; i++) memcpy(out[l].data() + 4 * i, &cv[i][l], 4);
}

}  // namespace simd
}  // namespace orc
