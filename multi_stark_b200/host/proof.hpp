// Proof objects and their wire format.
//
// `Proof` mirrors src/prover.rs:213-238 field for field; `to_bytes` follows the reference's bincode 2
// configuration `standard().with_little_endian().with_fixed_int_encoding()` (src/prover.rs:241-243):
// u64-LE length prefix per Vec, integers fixed-width LE, bool/u8 one byte, Option = tag byte + value,
// Goldilocks = canonical u64, extension element = its 2 coordinates, digest = 32 raw bytes.
// The FRI proof layout (`FriProof`, `QueryProof`, `BatchOpening`, `CommitPhaseProofStep`) belongs to p3-fri
// 0.5.1, which is not vendored in the reference: field ORDER here follows SURVEY Appendix A.6 and is
// PARITY UNPINNED (no golden proof bytes exist in the reference tree).
#pragma once
#include <cstdlib>
#include <string>
#include "challenger.hpp"
#include <optional>

namespace msh {

struct BatchOpening {
    std::vector<std::vector<Fp>> opened_values;  // one row per matrix, commit order
    std::vector<Digest> opening_proof;           // siblings, bottom-up
    bool operator==(const BatchOpening& o) const { return opened_values == o.opened_values && opening_proof == o.opening_proof; }
};
struct CommitPhaseProofStep {
    u8 log_arity = 1;
    std::vector<Fp2> sibling_values;  // arity - 1 values
    std::vector<Digest> opening_proof;
    bool operator==(const CommitPhaseProofStep& o) const {
        return log_arity == o.log_arity && sibling_values == o.sibling_values && opening_proof == o.opening_proof;
    }
};
struct QueryProof {
    std::vector<BatchOpening> input_proof;  // one per round
    std::vector<CommitPhaseProofStep> commit_phase_openings;
    bool operator==(const QueryProof& o) const { return input_proof == o.input_proof && commit_phase_openings == o.commit_phase_openings; }
};
struct FriProof {
    std::vector<Digest> commit_phase_commits;
    std::vector<Fp> commit_pow_witnesses;
    std::vector<QueryProof> query_proofs;
    std::vector<Fp2> final_poly;
    Fp query_pow_witness;
    bool operator==(const FriProof& o) const {
        return commit_phase_commits == o.commit_phase_commits && commit_pow_witnesses == o.commit_pow_witnesses &&
               query_proofs == o.query_proofs && final_poly == o.final_poly && query_pow_witness == o.query_pow_witness;
    }
};

// [matrix][point][column]
using OpenedValuesForRound = std::vector<std::vector<std::vector<Fp2>>>;

struct Proof {
    std::vector<bool> active;
    Digest stage_1_trace, stage_2_trace, quotient_chunks;  // Commitments
    std::vector<Fp2> intermediate_accumulators;
    std::vector<u8> log_degrees;
    FriProof opening_proof;
    OpenedValuesForRound quotient_opened_values;
    std::optional<OpenedValuesForRound> preprocessed_opened_values;
    OpenedValuesForRound stage_1_opened_values, stage_2_opened_values;
};

class ByteWriter {
  public:
    std::vector<u8> out;
    ByteWriter() { out.reserve(size_t(1) << 20); }  // a proof is ~1 MB: one allocation, appended 8 bytes at a time
    void u8_(u8 v) { out.push_back(v); }
    void u64_(u64 v) {
        u8 le[8];
        for (int b = 0; b < 8; b++) le[b] = (u8)(v >> (8 * b));
        out.insert(out.end(), le, le + 8);
    }
    void fp(Fp v) { u64_(v.v); }
    void fp2(const Fp2& v) { fp(v.c[0]); fp(v.c[1]); }
    void digest(const Digest& d) { out.insert(out.end(), d.begin(), d.end()); }
    template <class T, class F>
    void vec(const std::vector<T>& v, F&& f) { u64_(v.size()); for (auto& x : v) f(x); }
};

// ---- the one place that knows how a `Commitment` (src/types.rs:87) is laid out on the wire ------------------------------------
// The reference's commitments are p3-merkle-tree 0.5.1's (`Mmcs::new(.., cap_height)`, src/types.rs:199-207), whose serde
// shape is NOT verifiable in this tree (Plonky3 is an un-vendored git dependency and there is no cargo here):
//   Raw32  the 32 digest bytes as a fixed array (`Hash<F, u8, 32>`: serde arrays carry no length)            -- the default
//   CapVec a Merkle cap `Vec<[u8; 32]>`: u64 length prefix (2^cap_height = 1) followed by the digest bytes
// Every commitment in `Proof::to_bytes` -- the three trace commitments and the FRI commit-phase commitments -- goes through
// write_commitment / read_commitment, so `integration/golden_dump.rs` run against real Plonky3 settles it with one setting
// (MSH_COMMITMENT_WIRE=cap or set_commitment_wire). The transcript observes the digest bytes either way
// (`challenger.observe(commit)`, src/prover.rs:356).
enum class CommitmentWire { Raw32, CapVec };
inline CommitmentWire& commitment_wire_setting() {
    static CommitmentWire w = [] {
        const char* e = getenv("MSH_COMMITMENT_WIRE");
        return (e && std::string(e) == "cap") ? CommitmentWire::CapVec : CommitmentWire::Raw32;
    }();
    return w;
}
inline void set_commitment_wire(CommitmentWire w) { commitment_wire_setting() = w; }
inline void write_commitment(ByteWriter& w, const Digest& d) {
    if (commitment_wire_setting() == CommitmentWire::CapVec) w.u64_(1);
    w.digest(d);
}

inline void write_opened_round(ByteWriter& w, const OpenedValuesForRound& r) {
    w.vec(r, [&](const std::vector<std::vector<Fp2>>& m) {
        w.vec(m, [&](const std::vector<Fp2>& p) { w.vec(p, [&](const Fp2& v) { w.fp2(v); }); });
    });
}
inline void write_fri_proof(ByteWriter& w, const FriProof& f) {
    w.vec(f.commit_phase_commits, [&](const Digest& d) { write_commitment(w, d); });
    w.vec(f.commit_pow_witnesses, [&](Fp v) { w.fp(v); });
    w.vec(f.query_proofs, [&](const QueryProof& q) {
        w.vec(q.input_proof, [&](const BatchOpening& b) {
            w.vec(b.opened_values, [&](const std::vector<Fp>& row) { w.vec(row, [&](Fp v) { w.fp(v); }); });
            w.vec(b.opening_proof, [&](const Digest& d) { w.digest(d); });
        });
        w.vec(q.commit_phase_openings, [&](const CommitPhaseProofStep& s) {
            w.u8_(s.log_arity);
            w.vec(s.sibling_values, [&](const Fp2& v) { w.fp2(v); });
            w.vec(s.opening_proof, [&](const Digest& d) { w.digest(d); });
        });
    });
    w.vec(f.final_poly, [&](const Fp2& v) { w.fp2(v); });
    w.fp(f.query_pow_witness);
}
inline std::vector<u8> proof_to_bytes(const Proof& p) {
    ByteWriter w;
    w.u64_(p.active.size());
    for (bool b : p.active) w.u8_(b ? 1 : 0);
    write_commitment(w, p.stage_1_trace);
    write_commitment(w, p.stage_2_trace);
    write_commitment(w, p.quotient_chunks);
    w.vec(p.intermediate_accumulators, [&](const Fp2& v) { w.fp2(v); });
    w.vec(p.log_degrees, [&](u8 v) { w.u8_(v); });
    write_fri_proof(w, p.opening_proof);
    write_opened_round(w, p.quotient_opened_values);
    if (p.preprocessed_opened_values) { w.u8_(1); write_opened_round(w, *p.preprocessed_opened_values); }
    else w.u8_(0);
    write_opened_round(w, p.stage_1_opened_values);
    write_opened_round(w, p.stage_2_opened_values);
    return w.out;
}

// ---- reader (Proof::from_bytes, src/prover.rs:251-254) ------------------------------------------------
class ByteReader {
  public:
    ByteReader(const u8* p, size_t n) : p_(p), end_(p + n) {}
    bool ok = true;
    u8 u8_() { if (p_ >= end_) { ok = false; return 0; } return *p_++; }
    u64 u64_() { u64 v = 0; for (int b = 0; b < 8; b++) v |= (u64)u8_() << (8 * b); return v; }
    Fp fp() { u64 v = u64_(); if (v >= GL_P) ok = false; Fp r; r.v = v < GL_P ? v : 0; return r; }
    Fp2 fp2() { Fp a = fp(); Fp b = fp(); return Fp2(a, b); }
    Digest digest() { Digest d{}; for (auto& b : d) b = u8_(); return d; }
    size_t len() { u64 n = u64_(); if (n > (u64)(end_ - p_)) { ok = false; return 0; } return (size_t)n; }
    bool done() const { return p_ == end_; }
  private:
    const u8 *p_, *end_;
};
inline Digest read_commitment(ByteReader& r) {
    if (commitment_wire_setting() == CommitmentWire::CapVec && r.u64_() != 1) r.ok = false;  // cap_height 0: one digest
    return r.digest();
}
inline OpenedValuesForRound read_opened_round(ByteReader& r) {
    OpenedValuesForRound out(r.len());
    for (auto& m : out) { m.resize(r.len()); for (auto& p : m) { p.resize(r.len()); for (auto& v : p) v = r.fp2(); } }
    return out;
}
inline void read_fri_proof(ByteReader& r, FriProof& f) {
    f.commit_phase_commits.resize(r.len());
    for (auto& d : f.commit_phase_commits) d = read_commitment(r);
    f.commit_pow_witnesses.resize(r.len());
    for (auto& v : f.commit_pow_witnesses) v = r.fp();
    f.query_proofs.resize(r.len());
    for (auto& q : f.query_proofs) {
        q.input_proof.resize(r.len());
        for (auto& b : q.input_proof) {
            b.opened_values.resize(r.len());
            for (auto& row : b.opened_values) { row.resize(r.len()); for (auto& v : row) v = r.fp(); }
            b.opening_proof.resize(r.len());
            for (auto& d : b.opening_proof) d = r.digest();
        }
        q.commit_phase_openings.resize(r.len());
        for (auto& s : q.commit_phase_openings) {
            s.log_arity = r.u8_();
            s.sibling_values.resize(r.len());
            for (auto& v : s.sibling_values) v = r.fp2();
            s.opening_proof.resize(r.len());
            for (auto& d : s.opening_proof) d = r.digest();
        }
        if (!r.ok) return;
    }
    f.final_poly.resize(r.len());
    for (auto& v : f.final_poly) v = r.fp2();
    f.query_pow_witness = r.fp();
}
inline bool pcs_open_from_bytes(const u8* data, size_t n, std::vector<OpenedValuesForRound>& opened, FriProof& proof) {
    ByteReader r(data, n);
    opened.resize(r.len());
    for (auto& o : opened) o = read_opened_round(r);
    if (!r.ok) return false;
    read_fri_proof(r, proof);
    return r.ok && r.done();
}
inline bool proof_from_bytes(const u8* data, size_t n, Proof& p) {
    ByteReader r(data, n);
    p.active.resize(r.len());
    for (size_t i = 0; i < p.active.size(); i++) { u8 b = r.u8_(); if (b > 1) r.ok = false; p.active[i] = b == 1; }
    p.stage_1_trace = read_commitment(r); p.stage_2_trace = read_commitment(r); p.quotient_chunks = read_commitment(r);
    p.intermediate_accumulators.resize(r.len());
    for (auto& v : p.intermediate_accumulators) v = r.fp2();
    p.log_degrees.resize(r.len());
    for (auto& v : p.log_degrees) v = r.u8_();
    read_fri_proof(r, p.opening_proof);
    if (!r.ok) return false;
    p.quotient_opened_values = read_opened_round(r);
    u8 tag = r.u8_();
    if (tag > 1) r.ok = false;
    if (tag == 1) p.preprocessed_opened_values = read_opened_round(r);
    else p.preprocessed_opened_values.reset();
    p.stage_1_opened_values = read_opened_round(r);
    p.stage_2_opened_values = read_opened_round(r);
    return r.ok && r.done();
}

}  // namespace msh
