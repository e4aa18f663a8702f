// Host-side BLAKE3 (unkeyed, 256-bit output) for the Fiat-Shamir transcript and small host hashes.
//
// The reference hashes through p3-blake3 `Blake3` (src/types.rs:9,83,199), which wraps the
// `blake3` crate 1.8.5 (Cargo.lock:71-72): standard BLAKE3 `hash()`. This is a from-scratch
// restatement of the published BLAKE3 algorithm (7-round compression, 1024-byte chunks, binary
// tree of parent nodes with the largest-power-of-two left subtree). Bulk Merkle hashing of the
// committed matrices does NOT go through this file: it runs in csrc/blake3.cu on the GPU.
#pragma once
#include <cstdint>
#include <cstring>
#include <array>
#include <vector>

namespace msh {

using Digest = std::array<uint8_t, 32>;

namespace b3 {
constexpr uint32_t IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A,
                            0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
constexpr int MSG_PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum : uint32_t { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline void g(uint32_t* s, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
    s[a] = s[a] + s[b] + mx; s[d] = rotr(s[d] ^ s[a], 16);
    s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 12);
    s[a] = s[a] + s[b] + my; s[d] = rotr(s[d] ^ s[a], 8);
    s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 7);
}
// Seven rounds over a raw 16-word state and 16 message words; output = the 16 feed-forwarded
// words (first 8 = chaining value). The reference's known-answer test
// (src/test_circuits/blake3.rs:2646-2746) is stated on exactly this raw form.
static inline void compress_raw(const uint32_t state_in[16], const uint32_t msg[16], uint32_t out[16]) {
    uint32_t s[16], m[16], t[16];
    memcpy(s, state_in, 64); memcpy(m, msg, 64);
    for (int r = 0; r < 7; r++) {
        g(s, 0, 4, 8, 12, m[0], m[1]);   g(s, 1, 5, 9, 13, m[2], m[3]);
        g(s, 2, 6, 10, 14, m[4], m[5]);  g(s, 3, 7, 11, 15, m[6], m[7]);
        g(s, 0, 5, 10, 15, m[8], m[9]);  g(s, 1, 6, 11, 12, m[10], m[11]);
        g(s, 2, 7, 8, 13, m[12], m[13]); g(s, 3, 4, 9, 14, m[14], m[15]);
        if (r < 6) { for (int i = 0; i < 16; i++) t[i] = m[MSG_PERM[i]]; memcpy(m, t, 64); }
    }
    for (int i = 0; i < 8; i++) { out[i] = s[i] ^ s[i + 8]; out[i + 8] = s[i + 8] ^ state_in[i]; }
}
static inline void compress(const uint32_t cv[8], const uint32_t msg[16], uint64_t counter,
                            uint32_t block_len, uint32_t flags, uint32_t out[16]) {
    uint32_t st[16] = {cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], cv[6], cv[7],
                       IV[0], IV[1], IV[2], IV[3],
                       (uint32_t)counter, (uint32_t)(counter >> 32), block_len, flags};
    compress_raw(st, msg, out);
}
static inline void load_words(const uint8_t* p, size_t len, uint32_t w[16]) {
    uint8_t buf[64] = {0};
    memcpy(buf, p, len);
    for (int i = 0; i < 16; i++)
        w[i] = (uint32_t)buf[4 * i] | ((uint32_t)buf[4 * i + 1] << 8) | ((uint32_t)buf[4 * i + 2] << 16) |
               ((uint32_t)buf[4 * i + 3] << 24);
}
// Chaining value of one chunk (<= 1024 bytes); `root` marks a single-chunk message.
static inline void chunk_cv(const uint8_t* p, size_t len, uint64_t chunk_idx, bool root, uint32_t cv_out[8]) {
    uint32_t cv[8]; memcpy(cv, IV, 32);
    size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
    for (size_t b = 0; b < nblocks; b++) {
        size_t off = b * 64, bl = len - off < 64 ? len - off : 64;
        uint32_t w[16], out[16];
        load_words(p + off, bl, w);
        uint32_t flags = (b == 0 ? CHUNK_START : 0) | (b + 1 == nblocks ? CHUNK_END | (root ? ROOT : 0) : 0);
        compress(cv, w, chunk_idx, (uint32_t)bl, flags, out);
        memcpy(cv, out, 32);
    }
    memcpy(cv_out, cv, 32);
}
static inline void parent_cv(const uint32_t l[8], const uint32_t r[8], bool root, uint32_t out8[8]) {
    uint32_t w[16], out[16];
    memcpy(w, l, 32); memcpy(w + 8, r, 32);
    compress(IV, w, 0, 64, PARENT | (root ? ROOT : 0), out);
    memcpy(out8, out, 32);
}
// Subtree over chunks [first, first+count) of the message (count >= 1).
static inline void subtree(const uint8_t* data, size_t len, size_t first, size_t count, bool root, uint32_t out[8]) {
    if (count == 1) {
        size_t off = first * 1024, cl = len - off < 1024 ? len - off : 1024;
        chunk_cv(data + off, cl, first, root, out);
        return;
    }
    size_t left = 1;
    while (left * 2 < count) left *= 2;  // largest power of two strictly less than count
    uint32_t l[8], r[8];
    subtree(data, len, first, left, false, l);
    subtree(data, len, first + left, count - left, false, r);
    parent_cv(l, r, root, out);
}
}  // namespace b3

static inline Digest blake3_hash(const uint8_t* data, size_t len) {
    size_t nchunks = len == 0 ? 1 : (len + 1023) / 1024;
    uint32_t cv[8];
    b3::subtree(data, len, 0, nchunks, true, cv);
    Digest d;
    for (int i = 0; i < 8; i++) {
        d[4 * i] = (uint8_t)cv[i]; d[4 * i + 1] = (uint8_t)(cv[i] >> 8);
        d[4 * i + 2] = (uint8_t)(cv[i] >> 16); d[4 * i + 3] = (uint8_t)(cv[i] >> 24);
    }
    return d;
}
static inline Digest blake3_hash(const std::vector<uint8_t>& v) { return blake3_hash(v.data(), v.size()); }

}  // namespace msh
