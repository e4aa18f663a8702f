// Host Fiat-Shamir transcript: the reference's
//   Challenger = DeterministicPow<SerializingChallenger64<Val, HashChallenger<u8, Blake3, 32>>>
// (src/types.rs:28-29,44-81), seeded by GoldilocksBlake3Config::new (src/types.rs:111-140).
// p3-challenger 0.5.1 is not vendored in the reference; published semantics restated:
//   HashChallenger: input buffer starts as the seed; observe(byte) clears the output buffer and
//     appends; sample() flushes when the output buffer is empty (out = BLAKE3(input); input := out;
//     output := out) and pops from the END of the output buffer.
//   SerializingChallenger64: a field element is observed as its canonical u64 LE bytes, a digest as its
//     bytes; a base sample is rejection sampling of u64::from_le_bytes(8 sampled bytes) < p; an extension
//     sample is D base samples, coordinate 0 first; sample_bits(b) = low b bits of one such u64.
//   DeterministicPow: grind(0) returns ZERO and touches nothing; grind(b > 0) returns the SMALLEST witness w
//     with sample_bits(b) == 0 after observe(w) (the serial reference's answer; src/types.rs:31-42).
#pragma once
#include "goldilocks.hpp"
#include "blake3_host.hpp"
#include <functional>
#include <vector>

namespace msh {

struct CommitmentParameters {
    size_t log_blowup = 1;
    size_t cap_height = 0;
};
struct FriParameters {
    size_t log_final_poly_len = 0;
    size_t max_log_arity = 1;
    size_t num_queries = 100;
    size_t commit_proof_of_work_bits = 0;
    size_t query_proof_of_work_bits = 0;
};

class Challenger {
  public:
    Challenger() = default;
    explicit Challenger(std::vector<u8> seed) : input_(std::move(seed)) {}

    // src/types.rs:118-130: tag, then the seven parameters as u64 LE
    static Challenger for_config(const CommitmentParameters& c, const FriParameters& f) {
        static const char tag[] = "multi-stark/v0";
        std::vector<u8> seed(tag, tag + sizeof(tag) - 1);
        for (u64 p : {(u64)c.log_blowup, (u64)c.cap_height, (u64)f.log_final_poly_len, (u64)f.max_log_arity,
                      (u64)f.num_queries, (u64)f.commit_proof_of_work_bits, (u64)f.query_proof_of_work_bits})
            for (int b = 0; b < 8; b++) seed.push_back((u8)(p >> (8 * b)));
        return Challenger(std::move(seed));
    }

    void observe_byte(u8 b) {
        output_.clear();
        input_.push_back(b);
    }
    void observe(Fp v) {  // 8 x observe_byte
        output_.clear();
        u8 le[8];
        for (int b = 0; b < 8; b++) le[b] = (u8)(v.v >> (8 * b));
        input_.insert(input_.end(), le, le + 8);
    }
    void observe_usize(size_t x) { observe(Fp((u64)x)); }
    void observe(const Fp2& v) {  // observe_algebra_element
        observe(v.c[0]);
        observe(v.c[1]);
    }
    void observe(const Digest& d) {
        output_.clear();
        input_.insert(input_.end(), d.begin(), d.end());
    }
    // BLAKE3 of a long input buffer may be computed elsewhere (the device: gpu_backend.hpp); same digest by definition.
    using BigHash = std::function<Digest(const u8*, size_t)>;
    void set_big_hash(BigHash f, size_t threshold) { big_hash_ = std::move(f); big_threshold_ = threshold; }
    void observe_slice(const Fp* v, size_t n) {
        for (size_t i = 0; i < n; i++) observe(v[i]);
    }

    u8 sample_byte() {
        if (output_.empty()) flush();
        u8 b = output_.back();
        output_.pop_back();
        return b;
    }
    u64 sample_u64() {
        u64 v = 0;
        for (int b = 0; b < 8; b++) v |= (u64)sample_byte() << (8 * b);
        return v;
    }
    Fp sample_base() {
        for (;;) {
            u64 v = sample_u64();
            if (v < GL_P) { Fp r; r.v = v; return r; }
        }
    }
    Fp2 sample_ext() {  // sample_algebra_element
        Fp a = sample_base();
        Fp b = sample_base();
        return Fp2(a, b);
    }
    size_t sample_bits(size_t bits) { return (size_t)(sample_u64() & ((bits >= 64) ? ~0ull : ((1ull << bits) - 1))); }

    bool check_witness(size_t bits, Fp witness) {
        if (bits == 0) return true;
        observe(witness);
        return sample_bits(bits) == 0;
    }
    Fp grind(size_t bits) {
        if (bits == 0) return Fp::zero();
        for (u64 i = 0;; i++) {
            Challenger trial = *this;
            Fp w; w.v = i;
            if (trial.check_witness(bits, w)) {
                check_witness(bits, w);
                return w;
            }
        }
    }

    const std::vector<u8>& input_buffer() const { return input_; }
    // The caller observed further bytes "virtually" and hashed input_buffer() || those bytes elsewhere (the device);
    // continue exactly as after the flush() that the next sample would have performed.
    void set_flushed(const Digest& d) {
        input_.assign(d.begin(), d.end());
        output_.assign(d.begin(), d.end());
    }

  private:
    void flush() {
        Digest d = (big_hash_ && input_.size() >= big_threshold_) ? big_hash_(input_.data(), input_.size()) : blake3_hash(input_);
        input_.assign(d.begin(), d.end());
        output_.assign(d.begin(), d.end());
    }
    std::vector<u8> input_, output_;
    BigHash big_hash_;
    size_t big_threshold_ = 0;
};

}  // namespace msh
