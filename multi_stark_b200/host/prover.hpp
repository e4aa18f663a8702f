// prove(): host orchestration of the proof, restating System::new's preprocessed commitment
// (src/system.rs:180-203) and System::prove_multiple_claims (src/prover.rs:289-603) step for step over a
// `ProverBackend` that owns the heavy arithmetic (device: gpu_backend.hpp over the libmsgpu C ABI; CPU: the
// oracle's backend, used by tests only). Span names follow the reference's tracing spans
// (src/prover.rs:289,336,391,413,437,538).
#pragma once
#include "pcs.hpp"
#include "system.hpp"
#include <chrono>
#include <map>

namespace msh {

struct QuotientJob {  // one active circuit of the quotient stage (src/prover.rs:443-519)
    size_t circuit;       // canonical index
    size_t pos;           // active position = matrix index inside the stage-1 / stage-2 commitments
    int preprocessed_idx; // matrix index inside the preprocessed commitment, -1 if none
    unsigned log_degree, log_quotient_degree;
    Fp publics[8];        // beta, gamma, acc, next_acc as base coordinates (src/prover.rs:472-480)
};

struct ProverBackend {
    virtual ~ProverBackend() {}
    // Pcs::commit on host matrices over their natural domains (setup: preprocessed traces)
    virtual PcsHandlePtr commit(const std::vector<const Matrix*>& evals, Digest& root) = 0;
    // stage 1: commit the active traces; the backend keeps them (natural order) for the stage-2 construction
    virtual PcsHandlePtr commit_stage1(const std::vector<size_t>& circuits, const std::vector<MatrixView>& traces, Digest& root) = 0;
    // sum over claims of 1 / (beta + fingerprint(gamma, claim))  (src/prover.rs:381-387)
    virtual Fp2 claims_accumulator(const ClaimsView& claims, Fp2 beta, Fp2 gamma) = 0;
    // Optional: the claims of the proof about to start; a device backend begins their upload once the traces are on the
    // device, under the stage-1 kernels (the claims are first needed after the stage-1 root, src/prover.rs:364-387).
    virtual void announce_claims(const ClaimsView&) {}
    // Optional: observe every claim (length-prefixed, src/prover.rs:369-372) into the transcript on the backend's side and
    // leave the challenger flushed (the next transcript operation is always a sample, src/prover.rs:376). false = not done.
    virtual bool observe_claims(Challenger&, const ClaimsView&) { return false; }
    // optional accelerator for the transcript's BLAKE3 over a long input buffer (2^20 claims = 42 MB of observations)
    virtual Challenger::BigHash big_hash() { return nullptr; }
    // stage 2: lookup values from the kept traces, stage-2 traces, their commitment; `intermediate` = running accumulator
    // after each active circuit (src/prover.rs:391-421, src/lookup.rs:472-555)
    virtual PcsHandlePtr commit_stage2(Fp2 beta, Fp2 gamma, Fp2 acc, std::vector<Fp2>& intermediate, Digest& root) = 0;
    // quotient stage for every active circuit + Pcs::commit_ldes (src/prover.rs:437-527)
    virtual PcsHandlePtr commit_quotient(const std::vector<QuotientJob>& jobs, PcsHandle* pre, PcsHandle* s1, PcsHandle* s2, Fp2 alpha,
                                         Digest& root) = 0;
    virtual std::unique_ptr<OpenDevice> open_begin(const std::vector<OpenRound>& rounds) = 0;
    virtual void end_proof() {}  // drop per-proof state (kept traces)
};

struct ProverKey {  // src/system.rs:104-107
    PcsHandlePtr preprocessed_data;
    bool has_preprocessed = false;
    Digest preprocessed_commit{};
};

struct ProveTimings {
    std::map<std::string, double> ms;
    TranscriptTrace trace;  // beta, gamma, alpha, zeta, alpha_pcs, FRI betas; query indices
};

class Prover {
  public:
    Prover(const SystemShape& shape, ProverBackend& backend) : shape_(shape), be_(backend) {
        // System::new tail: one commitment over all preprocessed traces (src/system.rs:180-196)
        std::vector<const Matrix*> pre;
        for (auto& c : shape_.circuits)
            if (c.has_preprocessed) pre.push_back(&c.preprocessed);
        if (!pre.empty()) {
            key_.preprocessed_data = be_.commit(pre, key_.preprocessed_commit);
            key_.has_preprocessed = true;
        }
    }
    const ProverKey& key() const { return key_; }

    // traces[i] = stage-1 trace of circuit i (height 0 = inactive)
    Proof prove(const std::vector<std::vector<Fp>>& claims, const std::vector<const Matrix*>& traces, ProveTimings* tm = nullptr) {
        std::vector<Fp> flat;
        std::vector<u64> offsets(1, 0);
        for (auto& c : claims) { flat.insert(flat.end(), c.begin(), c.end()); offsets.push_back(flat.size()); }
        ClaimsView cv;
        cv.values = flat.data(); cv.offsets = offsets.data(); cv.n = claims.size();
        std::vector<MatrixView> views;
        for (auto* m : traces) views.push_back(MatrixView(*m));
        return prove(cv, views, tm);
    }

    Proof prove(const ClaimsView& claims, const std::vector<MatrixView>& traces, ProveTimings* tm = nullptr) {
        using clk = std::chrono::steady_clock;
        auto t0 = clk::now();
        auto lap = [&](const char* name) {
            auto t1 = clk::now();
            if (tm) tm->ms[name] += std::chrono::duration<double, std::milli>(t1 - t0).count();
            t0 = t1;
        };
        if (traces.size() != shape_.circuits.size()) throw std::runtime_error("expected one trace per circuit");
        Challenger ch = Challenger::for_config(shape_.commitment, shape_.fri);
        if (auto bh = be_.big_hash()) ch.set_big_hash(bh, size_t(1) << 16);
        shape_.observe_shape(ch);
        Proof proof;
        std::vector<size_t> active_indices;
        for (size_t i = 0; i < traces.size(); i++) {
            bool a = traces[i].height() > 0;
            proof.active.push_back(a);
            ch.observe(Fp(a ? 1 : 0));
            if (a) active_indices.push_back(i);
            if (a && shape_.circuits[i].has_preprocessed && shape_.circuits[i].preprocessed_height != traces[i].height())
                throw std::runtime_error("main trace height must equal preprocessed trace height");
            if (a && traces[i].width != shape_.circuits[i].main_width) throw std::runtime_error("trace width does not match the circuit");
        }
        if (active_indices.empty()) throw std::runtime_error("cannot prove with every circuit deactivated (all traces empty)");
        std::vector<int> active_pos(traces.size(), -1);
        for (size_t p = 0; p < active_indices.size(); p++) active_pos[active_indices[p]] = (int)p;

        // stark/stage1_commit
        std::vector<unsigned> log_degrees;
        std::vector<MatrixView> active_traces;
        for (size_t ci : active_indices) {
            if (traces[ci].height() & (traces[ci].height() - 1)) throw std::runtime_error("trace height must be a power of two");
            log_degrees.push_back(log2_strict(traces[ci].height()));
            active_traces.push_back(traces[ci]);
        }
        be_.announce_claims(claims);
        PcsHandlePtr s1 = be_.commit_stage1(active_indices, active_traces, proof.stage_1_trace);
        lap("stark/stage1_commit");
        if (key_.has_preprocessed) ch.observe(key_.preprocessed_commit);
        ch.observe(proof.stage_1_trace);
        for (unsigned ld : log_degrees) ch.observe_usize(ld);
        ch.observe_usize(claims.size());
        if (!be_.observe_claims(ch, claims))
            for (size_t i = 0; i < claims.size(); i++) {
                ch.observe_usize(claims.len(i));
                ch.observe_slice(claims.at(i), claims.len(i));
            }
        Fp2 beta = ch.sample_ext();
        ch.observe(beta);
        Fp2 gamma = ch.sample_ext();
        ch.observe(gamma);
        Fp2 acc = be_.claims_accumulator(claims, beta, gamma);
        lap("stark/claims");

        // stark/lookup_construction + stark/stage2_commit
        PcsHandlePtr s2 = be_.commit_stage2(beta, gamma, acc, proof.intermediate_accumulators, proof.stage_2_trace);
        lap("stark/stage2_commit");
        ch.observe(proof.stage_2_trace);
        for (auto& a : proof.intermediate_accumulators) ch.observe(a);
        Fp2 alpha = ch.sample_ext();

        // stark/quotient
        std::vector<QuotientJob> jobs;
        for (size_t pos = 0; pos < active_indices.size(); pos++) {
            size_t ci = active_indices[pos];
            QuotientJob j;
            j.circuit = ci;
            j.pos = pos;
            j.preprocessed_idx = shape_.preprocessed_indices[ci];
            j.log_degree = log_degrees[pos];
            j.log_quotient_degree = log2_strict(shape_.circuits[ci].quotient_degree());
            Fp2 next_acc = proof.intermediate_accumulators[pos];
            const Fp2 pubs[4] = {beta, gamma, acc, next_acc};
            for (int k = 0; k < 4; k++) { j.publics[2 * k] = pubs[k].c[0]; j.publics[2 * k + 1] = pubs[k].c[1]; }
            acc = next_acc;
            jobs.push_back(j);
        }
        PcsHandlePtr qd = be_.commit_quotient(jobs, key_.preprocessed_data.get(), s1.get(), s2.get(), alpha, proof.quotient_chunks);
        ch.observe(proof.quotient_chunks);
        lap("stark/quotient");

        // stark/fri_open
        Fp2 zeta = ch.sample_ext();
        if (tm) tm->trace.challenges = {beta, gamma, alpha, zeta};
        std::vector<OpenRound> rounds(3);
        rounds[0].data = s1.get();
        rounds[1].data = s2.get();
        rounds[2].data = qd.get();
        for (unsigned ld : log_degrees) {
            Fp2 zeta_next = zeta * two_adic_generator(ld);  // trace_domain.next_point(zeta), src/prover.rs:546-548
            rounds[0].points.push_back({zeta, zeta_next});
            rounds[1].points.push_back({zeta, zeta_next});
            rounds[2].points.push_back({zeta});
        }
        if (key_.has_preprocessed) {
            OpenRound r0;
            r0.data = key_.preprocessed_data.get();
            for (size_t ci = 0; ci < shape_.circuits.size(); ci++) {
                if (shape_.preprocessed_indices[ci] < 0) continue;
                if (active_pos[ci] >= 0) {
                    Fp2 zeta_next = zeta * two_adic_generator(log_degrees[active_pos[ci]]);
                    r0.points.push_back({zeta, zeta_next});
                } else {
                    r0.points.push_back({});
                }
            }
            rounds.push_back(std::move(r0));
        }
        std::unique_ptr<OpenDevice> dev = be_.open_begin(rounds);
        std::vector<OpenedValuesForRound> opened;
        pcs_open(*dev, rounds, shape_.commitment, shape_.fri, ch, opened, proof.opening_proof, tm ? &tm->ms : nullptr,
                 tm ? &tm->trace : nullptr);
        dev.reset();
        proof.stage_1_opened_values = std::move(opened[0]);
        proof.stage_2_opened_values = std::move(opened[1]);
        proof.quotient_opened_values = std::move(opened[2]);
        if (opened.size() > 3) proof.preprocessed_opened_values = std::move(opened[3]);
        for (unsigned ld : log_degrees) proof.log_degrees.push_back((u8)ld);
        be_.end_proof();
        lap("stark/fri_open");
        return proof;
    }

  private:
    const SystemShape& shape_;
    ProverBackend& be_;
    ProverKey key_;
};

}  // namespace msh
