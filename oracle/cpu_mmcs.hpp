// ORACLE (test infrastructure, not product code): CPU restatement of the mixed-matrix Merkle
// commitment the reference configures at src/types.rs:82-84,199-207:
//   Mmcs = MerkleTreeMmcs<Val, u8, SerializingHasher<Blake3>, CompressionFunctionFromHasher<Blake3,2,32>, 2, 32>
// (p3-merkle-tree / p3-symmetric / p3-blake3 0.5.1, rev e9d75614, not vendored). Restated semantics
// (SURVEY.md Appendix A.4):
//   leaf  = BLAKE3( canonical u64 little-endian bytes of the row, rows of all matrices of that
//           height concatenated in the order the matrices were given )
//   node  = BLAKE3( left || right )                      (64 bytes -> 32 bytes)
//   matrices stably sorted tallest first; when a layer reaches the (power-of-two) height of the
//   next matrices: next[i] = compress( compress(prev[2i], prev[2i+1]), leaf(rows_i) )
//   open_batch(index): matrix m opens row index >> (log_max_height - log_height_m); the proof is
//   the sibling digests bottom-up; opened rows are listed in the ORIGINAL matrix order.
// PARITY UNPINNED against stored reference roots: the reference holds no golden roots
// (src/types.rs:246-319 only PRINTS them). Leaf and node hashes are pinned against the official
// `blake3` Python package in tests/test_oracle_hash.py.
#pragma once
#include "../multi_stark_b200/host/goldilocks.hpp"
#include "../multi_stark_b200/host/blake3_host.hpp"
#include "cpu_simd.hpp"
#include <numeric>
#include <stdexcept>

namespace orc {
using namespace msh;

struct MatView {
    const Fp* data;
    size_t height, width;
    const Fp* row(size_t r) const { return data + r * width; }
};

// SerializingHasher<Blake3>::hash_iter over the concatenated rows `r` of `mats`.
inline Digest hash_rows(const std::vector<MatView>& mats, size_t r) {
    size_t total = 0;
    for (auto& m : mats) total += m.width;
    std::vector<uint8_t> buf(total * 8);
    size_t o = 0;
    for (auto& m : mats) {
        const Fp* row = m.row(r);
        for (size_t c = 0; c < m.width; c++) {
            u64 v = row[c].v;
            for (int b = 0; b < 8; b++) buf[o++] = (uint8_t)(v >> (8 * b));
        }
    }
    return blake3_hash(buf);
}
// leaf digests of rows [r0, r0 + n), n <= simd::kLanes, one row per SIMD lane (rows of up to 1024 bytes = one BLAKE3 chunk;
// wider rows take the scalar multi-chunk path)
inline void hash_rows_many(const std::vector<MatView>& mats, size_t r0, int n, Digest* out) {
    size_t total = 0;
    for (auto& m : mats) total += m.width;
    if (total * 8 > 1024 || n < 2) {
        for (int l = 0; l < n; l++) out[l] = hash_rows(mats, r0 + l);
        return;
    }
    // canonical u64 values are their own little-endian bytes on this (little-endian) host: concatenate the rows
    std::vector<uint8_t> buf((size_t)simd::kLanes * total * 8 + 64);
    const uint8_t* msgs[simd::kLanes];
    for (int l = 0; l < n; l++) {
        uint8_t* dst = buf.data() + (size_t)l * total * 8;
        msgs[l] = dst;
        for (auto& m : mats) {
            memcpy(dst, m.row(r0 + l), m.width * 8);
            dst += m.width * 8;
        }
    }
    simd::b3_hash_lanes_one_chunk(msgs, n, total * 8, out);
}
// next[i] = compress(prev[2i], prev[2i+1]) for i in [i0, i0 + n), n <= simd::kLanes
inline void compress_many(const Digest* prev, size_t i0, int n, Digest* out) {
    const uint8_t* msgs[simd::kLanes];
    for (int l = 0; l < n; l++) msgs[l] = prev[2 * (i0 + l)].data();  // 64 contiguous bytes: two adjacent digests
    simd::b3_hash_lanes_one_chunk(msgs, n, 64, out);
}
inline void compress_pairs_many(const Digest* l, const Digest* r, int n, Digest* out) {
    uint8_t buf[simd::kLanes][64];
    const uint8_t* msgs[simd::kLanes];
    for (int k = 0; k < n; k++) {
        memcpy(buf[k], l[k].data(), 32);
        memcpy(buf[k] + 32, r[k].data(), 32);
        msgs[k] = buf[k];
    }
    simd::b3_hash_lanes_one_chunk(msgs, n, 64, out);
}
inline Digest hash_values(const Fp* vals, size_t n) {
    MatView m{vals, 1, n};
    return hash_rows({m}, 0);
}
inline Digest compress2(const Digest& l, const Digest& r) {
    uint8_t buf[64];
    memcpy(buf, l.data(), 32); memcpy(buf + 32, r.data(), 32);
    return blake3_hash(buf, 64);
}

struct MerkleTree {
    std::vector<MatView> leaves;                   // original order
    std::vector<std::vector<Digest>> digest_layers;  // layer 0 = leaf digests of the tallest matrices
    size_t max_height = 0;
    const Digest& root() const { return digest_layers.back()[0]; }
};

inline MerkleTree merkle_commit(const std::vector<MatView>& leaves) {
    if (leaves.empty()) throw std::runtime_error("No matrices given?");
    MerkleTree t;
    t.leaves = leaves;
    std::vector<size_t> order(leaves.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return leaves[a].height > leaves[b].height; });
    for (auto& m : leaves)
        if (m.height == 0 || (m.height & (m.height - 1))) throw std::runtime_error("matrix heights must be powers of two");
    size_t pos = 0;
    size_t max_h = leaves[order[0]].height;
    t.max_height = max_h;
    std::vector<MatView> group;
    while (pos < order.size() && leaves[order[pos]].height == max_h) group.push_back(leaves[order[pos++]]);
    std::vector<Digest> layer(max_h);
    {
        long long nn = (long long)max_h;
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < nn; i += simd::kLanes)
            hash_rows_many(group, (size_t)i, (int)std::min<long long>(simd::kLanes, nn - i), &layer[i]);
    }
    t.digest_layers.push_back(std::move(layer));
    while (t.digest_layers.back().size() > 1) {
        const std::vector<Digest>& prev = t.digest_layers.back();
        size_t next_len = prev.size() / 2;
        group.clear();
        while (pos < order.size() && leaves[order[pos]].height == next_len) group.push_back(leaves[order[pos++]]);
        std::vector<Digest> next(next_len);
        long long nn = (long long)next_len;
        bool inject = !group.empty();
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < nn; i += simd::kLanes) {
            const int cnt = (int)std::min<long long>(simd::kLanes, nn - i);
            Digest d[simd::kLanes];
            compress_many(prev.data(), (size_t)i, cnt, d);
            if (inject) {
                Digest leaf[simd::kLanes];
                hash_rows_many(group, (size_t)i, cnt, leaf);
                compress_pairs_many(d, leaf, cnt, &next[i]);
            } else {
                for (int k = 0; k < cnt; k++) next[i + k] = d[k];
            }
        }
        t.digest_layers.push_back(std::move(next));
    }
    if (pos != order.size()) throw std::runtime_error("matrix shorter than the tree root layer");
    return t;
}

struct BatchOpening {
    std::vector<std::vector<Fp>> opened_values;  // one row per matrix, original order
    std::vector<Digest> opening_proof;           // siblings, bottom-up
};

inline BatchOpening open_batch(const MerkleTree& t, size_t index) {
    BatchOpening bo;
    unsigned log_max = log2_strict(t.max_height);
    for (auto& m : t.leaves) {
        unsigned lh = log2_strict(m.height);
        size_t r = index >> (log_max - lh);
        bo.opened_values.emplace_back(m.row(r), m.row(r) + m.width);
    }
    for (unsigned i = 0; i < log_max; i++) bo.opening_proof.push_back(t.digest_layers[i][(index >> i) ^ 1]);
    return bo;
}

// MerkleTreeMmcs::verify_batch. `heights[i]` is the height of matrix i (original order).
inline bool verify_batch(const Digest& commit, const std::vector<size_t>& heights, size_t index,
                         const std::vector<std::vector<Fp>>& opened_values, const std::vector<Digest>& proof) {
    if (heights.size() != opened_values.size() || heights.empty()) return false;
    std::vector<size_t> order(heights.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return heights[a] > heights[b]; });
    size_t pos = 0;
    size_t cur_h = heights[order[0]];
    if (cur_h == 0 || (cur_h & (cur_h - 1))) return false;
    if (proof.size() != log2_strict(cur_h)) return false;
    if (index >= cur_h) return false;
    auto hash_group = [&](size_t h) {
        std::vector<MatView> g;
        while (pos < order.size() && heights[order[pos]] == h) {
            auto& v = opened_values[order[pos++]];
            g.push_back(MatView{v.data(), 1, v.size()});
        }
        return hash_rows(g, 0);
    };
    Digest root = hash_group(cur_h);
    for (const Digest& sib : proof) {
        root = (index & 1) == 0 ? compress2(root, sib) : compress2(sib, root);
        index >>= 1;
        cur_h >>= 1;
        if (pos < order.size() && heights[order[pos]] == cur_h) root = compress2(root, hash_group(cur_h));
    }
    if (pos != order.size()) return false;
    return root == commit;
}

}  // namespace orc
