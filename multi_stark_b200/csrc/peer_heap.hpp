// The allocator inside one peer window (peer.cu): first fit at the LOWEST offset over an ordered free list, with coalescing.
// Deterministic by construction -- the result of a call depends only on the sequence of calls made so far -- which is what lets
// every rank of the row-sharded prover compute the same (segment, offset) for a block without exchanging addresses.
// No CUDA in this header: tests/test_peer_heap.py drives it on the CPU against a Python model.
#pragma once
#include <cstddef>
#include <map>

namespace msg {

class FirstFitHeap {
  public:
    static constexpr size_t kAlign = 512;
    static constexpr size_t npos = (size_t)-1;

    FirstFitHeap() = default;
    // manages offsets [first, bytes)
    FirstFitHeap(size_t first, size_t bytes) {
        if (bytes > first) free_[first] = bytes - first;
    }
    static size_t round_up(size_t bytes) {
        const size_t need = (bytes + kAlign - 1) / kAlign * kAlign;
        return need < kAlign ? kAlign : need;
    }
    // offset of a block of at least `bytes` (rounded up to kAlign), or npos
    size_t alloc(size_t bytes) {
        const size_t need = round_up(bytes);
        for (auto it = free_.begin(); it != free_.end(); ++it) {
            if (it->second < need) continue;
            const size_t o = it->first, sz = it->second;
            free_.erase(it);
            if (sz > need) free_[o + need] = sz - need;
            used_[o] = need;
            return o;
        }
        return npos;
    }
    // false: not the offset of a live block
    bool free(size_t off) {
        auto it = used_.find(off);
        if (it == used_.end()) return false;
        size_t o = it->first, sz = it->second;
        used_.erase(it);
        auto nx = free_.lower_bound(o);
        if (nx != free_.end() && o + sz == nx->first) {  // merge with the free block behind
            sz += nx->second;
            nx = free_.erase(nx);
        }
        if (nx != free_.begin()) {  // merge with the free block in front
            auto pv = std::prev(nx);
            if (pv->first + pv->second == o) {
                pv->second += sz;
                return true;
            }
        }
        free_[o] = sz;
        return true;
    }
    size_t live_blocks() const { return used_.size(); }
    size_t free_blocks() const { return free_.size(); }
    size_t free_bytes() const {
        size_t t = 0;
        for (auto& f : free_) t += f.second;
        return t;
    }

  private:
    std::map<size_t, size_t> free_;  // offset -> size
    std::map<size_t, size_t> used_;  // offset -> size
};

}  // namespace msg
