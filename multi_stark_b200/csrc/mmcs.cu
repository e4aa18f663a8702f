// Mixed-matrix Merkle commitment on the device: tree construction with injection of shorter
// matrices and batched query openings. Semantics of p3-merkle-tree `MerkleTreeMmcs::{commit,
// open_batch}` as configured by the reference at src/types.rs:82-84,199-207.
#include "mmcs.hpp"
#include <algorithm>
#include <numeric>

namespace msg {

void mmcs_build(Ctx& c, msgpu_pdata* pd) {
    StageScope stage_scope(c, "merkle");
    MSG_REQUIRE(!pd->mats.empty(), "commit: no matrices given");
    std::vector<size_t> order(pd->mats.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](size_t a, size_t b) { return pd->mats[a].height > pd->mats[b].height; });
    pd->total_width = 0;
    for (auto& m : pd->mats) {
        MSG_REQUIRE(is_pow2(m.height), "commit: matrix heights must be powers of two");
        pd->total_width += m.width;
    }
    u64 max_h = pd->mats[order[0]].height;
    pd->max_height = max_h;
    u64 total = 2 * max_h - 1;
    pd->digests = (uint8_t*)c.alloc(total * 32);
    pd->layer_off.clear();
    pd->layer_len.clear();
    u64 off = 0;
    for (u64 len = max_h; len >= 1; len >>= 1) {
        pd->layer_off.push_back(off);
        pd->layer_len.push_back(len);
        off += len;
        if (len == 1) break;
    }
    size_t pos = 0;
    std::vector<MatRef> group;
    while (pos < order.size() && pd->mats[order[pos]].height == max_h) {
        auto& m = pd->mats[order[pos++]];
        group.push_back(MatRef{m.ptr, m.height, m.width});
    }
    b3_hash_rows(c, group, pd->digests);
    // leaf digests of the shorter matrices, per layer they are injected into
    size_t n_layers = pd->layer_len.size();
    std::vector<uint8_t*> inj(n_layers, nullptr);
    for (size_t l = 1; l < n_layers; l++) {
        u64 next_len = pd->layer_len[l];
        group.clear();
        while (pos < order.size() && pd->mats[order[pos]].height == next_len) {
            auto& m = pd->mats[order[pos++]];
            group.push_back(MatRef{m.ptr, m.height, m.width});
        }
        if (!group.empty()) {
            inj[l] = (uint8_t*)c.alloc(next_len * 32);
            b3_hash_rows(c, group, inj[l]);
        }
    }
    // node layers: 10 per launch while the layer is large, then everything that is left in one CTA
    size_t l = 0;
    while (l + 1 < n_layers) {
        u64 len = pd->layer_len[l];
        u32 levels = len > 4096 ? 10 : ilog2(len);
        uint8_t* outs[16];
        const uint8_t* injs[16];
        for (u32 k = 0; k < levels; k++) {
            outs[k] = pd->digests + pd->layer_off[l + 1 + k] * 32;
            injs[k] = inj[l + 1 + k];
        }
        b3_merkle_subtrees(c, pd->digests + pd->layer_off[l] * 32, len, levels, outs, injs);
        l += levels;
    }
    for (auto* p : inj)
        if (p) c.free(p);
    MSG_REQUIRE(pos == order.size(), "commit: internal error, matrix not placed in the tree");
    MSG_CUDA(cudaMemcpyAsync(pd->root, pd->digests + pd->layer_off.back() * 32, 32, cudaMemcpyDeviceToHost, c.stream));
    c.sync();
}

struct OpenMat {
    const u64* ptr;
    u32 width;
    u32 shift;    // log_max_height - log_height
    u64 out_off;  // column offset inside one query's opened block
};

__global__ void k_open_rows(const OpenMat* mats, u32 nmats, const u64* idx, u64 total_width, u64* out) {
    const u64 q = blockIdx.x;
    const u64 index = idx[q];
    for (u32 k = 0; k < nmats; k++) {
        const OpenMat m = mats[k];
        const u64* row = m.ptr + (index >> m.shift) * m.width;
        u64* o = out + q * total_width + m.out_off;
        for (u32 cidx = threadIdx.x; cidx < m.width; cidx += blockDim.x) o[cidx] = row[cidx];
    }
}

__global__ void k_open_proof(const uint4* digests, const u64* layer_off, u32 depth, const u64* idx, u64 n_idx, uint4* out) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 total = n_idx * depth * 2;
    if (t >= total) return;
    u32 half = (u32)(t & 1);
    u64 r = t >> 1;
    u32 lvl = (u32)(r % depth);
    u64 q = r / depth;
    u64 sib = (idx[q] >> lvl) ^ 1;
    out[t] = digests[(layer_off[lvl] + sib) * 2 + half];
}

void mmcs_open_batch(Ctx& c, const msgpu_pdata* pd, const u64* indices_host, u64 n_idx, u64* opened_host,
                     uint8_t* proof_host) {
    if (n_idx == 0) return;
    u32 log_max = ilog2(pd->max_height);
    for (u64 i = 0; i < n_idx; i++) MSG_REQUIRE(indices_host[i] < pd->max_height, "open_batch: index out of range");
    std::vector<OpenMat> om;
    u64 off = 0;
    for (auto& m : pd->mats) {
        om.push_back(OpenMat{m.ptr, (u32)m.width, log_max - ilog2(m.height), off});
        off += m.width;
    }
    u64 tw = pd->total_width;
    size_t sz_mats = om.size() * sizeof(OpenMat), sz_idx = n_idx * 8, sz_off = pd->layer_off.size() * 8;
    OpenMat* d_mats = (OpenMat*)c.alloc(sz_mats);
    u64* d_idx = (u64*)c.alloc(sz_idx);
    u64* d_off = (u64*)c.alloc(sz_off);
    u64* d_open = (u64*)c.alloc(std::max<u64>(n_idx * tw * 8, 8));
    uint8_t* d_proof = (uint8_t*)c.alloc(std::max<u64>(n_idx * log_max * 32, 32));
    MSG_CUDA(cudaMemcpyAsync(d_mats, om.data(), sz_mats, cudaMemcpyHostToDevice, c.stream));
    MSG_CUDA(cudaMemcpyAsync(d_idx, indices_host, sz_idx, cudaMemcpyHostToDevice, c.stream));
    MSG_CUDA(cudaMemcpyAsync(d_off, pd->layer_off.data(), sz_off, cudaMemcpyHostToDevice, c.stream));
    if (tw) {
        {
            KLaunch kl(c, "k_open_rows");
            k_open_rows<<<(unsigned)n_idx, 128, 0, c.stream>>>(d_mats, (u32)om.size(), d_idx, tw, d_open);
        }
        MSG_CUDA(cudaGetLastError());
        MSG_CUDA(cudaMemcpyAsync(opened_host, d_open, n_idx * tw * 8, cudaMemcpyDeviceToHost, c.stream));
    }
    if (log_max) {
        u64 total = n_idx * log_max * 2;
        {
            KLaunch kl(c, "k_open_proof");
            k_open_proof<<<(unsigned)((total + 127) / 128), 128, 0, c.stream>>>((const uint4*)pd->digests, d_off, log_max,
                                                                                d_idx, n_idx, (uint4*)d_proof);
        }
        MSG_CUDA(cudaGetLastError());
        MSG_CUDA(cudaMemcpyAsync(proof_host, d_proof, n_idx * log_max * 32, cudaMemcpyDeviceToHost, c.stream));
    }
    c.sync();
    c.free(d_mats);
    c.free(d_idx);
    c.free(d_off);
    c.free(d_open);
    c.free(d_proof);
}

void pdata_destroy(msgpu_pdata* pd) {
    if (!pd) return;
    Ctx& c = *pd->ctx;
    for (auto& m : pd->mats)
        if (m.owned && m.ptr) c.free(m.ptr);
    if (pd->digests) c.free(pd->digests);
    delete pd;
}

}  // namespace msg
