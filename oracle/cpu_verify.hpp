// ORACLE (test infrastructure, not product code): CPU restatement of the verifier -- the acceptance oracle for
// proofs made by the device prover. Follows, step for step:
//   System::verify_shape              src/verifier.rs:536-695
//   System::verify_multiple_claims    src/verifier.rs:208-532   (transcript replay :255-326, rounds :333-411,
//                                     pcs.verify :413, OOD check :419-530)
//   TwoAdicFriPcs::verify + verifier::verify_fri / verify_query + TwoAdicFriFolding::fold_row of p3-fri 0.5.1
//   (rev e9d75614, not vendored in the reference; call site src/verifier.rs:413), published semantics:
//     observe every opened value (round, matrix, point, column); sample alpha;
//     per commit-phase commitment: observe, check_witness(commit_pow_bits), sample beta;
//     observe the final polynomial; check_witness(query_pow_bits);
//     log_max_height = #commits + log_blowup + log_final_poly_len;
//     per query: index = sample_bits(log_max_height); per round verify_batch at index >> (log_max - log_h_round) and
//       accumulate per LDE height  ro += alpha_pow * (p(z) - p(x)) / (z - x), alpha_pow *= alpha  (matrix, point, column);
//       a height == log_blowup entry must be zero and is dropped;
//     folded = ro[log_max]; per layer (log_folded = log_max-1 ... log_final): evals[index&1] = folded,
//       evals[(index&1)^1] = sibling; index >>= 1; verify_batch(width 2, height 2^log_folded); folded = interpolation of
//       (x0, e0), (-x0, e1) at beta, x0 = w_{2^(log_folded+1)}^{rev(index)}; if an input of height 2^log_folded exists
//       folded += beta^2 * ro; finally folded == final_poly(w_{log_max}^{rev(index, log_max)}).
// This file shares NO code with the prover path except field / hash / transcript primitives, so prover/verifier
// agreement is a real check of the device prover's FRI rules (the reference's own end-to-end tests give the same kind
// of assurance, src/verifier.rs:783-826). PARITY UNPINNED against real p3-fri (no golden proofs in the reference tree).
#pragma once
#include "../multi_stark_b200/host/proof.hpp"
#include "../multi_stark_b200/host/system.hpp"
#include "cpu_eval.hpp"
#include "cpu_mmcs.hpp"
#include <map>

namespace orc {
using namespace msh;

enum class VerifyError : int {
    Ok = 0,
    InvalidClaim = 1,
    InvalidOpeningArgument = 2,
    InvalidProofShape = 3,
    InvalidSystem = 4,
    OodEvaluationMismatch = 5,
    UnbalancedChannel = 6,
};

// One matrix of an opening round as the verifier sees it: trace-domain log size and (point, claimed values) pairs.
struct VerifierMat {
    unsigned log_degree;
    std::vector<std::pair<Fp2, std::vector<Fp2>>> points;
};
struct VerifierRound {
    Digest commit;
    std::vector<VerifierMat> mats;
};

// Interpolate the line through (x0, e0), (x1, e1) and evaluate at beta (TwoAdicFriFolding::fold_row).
inline Fp2 fold_row(size_t index, unsigned log_height, Fp2 beta, Fp2 e0, Fp2 e1) {
    Fp x0 = two_adic_generator(log_height + 1).pow((u64)reverse_bits_len(index, log_height));
    Fp x1 = -x0;
    return e0 + (beta - x0) * (e1 - e0) * (x1 - x0).inverse();
}

// TwoAdicFriPcs::verify. Returns false on any failure.
inline bool pcs_verify(const std::vector<VerifierRound>& rounds, const FriProof& proof, const CommitmentParameters& cp,
                       const FriParameters& fp, Challenger& ch) {
    const unsigned log_blowup = (unsigned)cp.log_blowup;
    for (auto& r : rounds)
        for (auto& m : r.mats)
            for (auto& pt : m.points)
                for (auto& y : pt.second) ch.observe(y);
    Fp2 alpha = ch.sample_ext();

    if (proof.commit_pow_witnesses.size() != proof.commit_phase_commits.size()) return false;
    std::vector<Fp2> betas;
    for (size_t i = 0; i < proof.commit_phase_commits.size(); i++) {
        ch.observe(proof.commit_phase_commits[i]);
        if (!ch.check_witness(fp.commit_proof_of_work_bits, proof.commit_pow_witnesses[i])) return false;
        betas.push_back(ch.sample_ext());
    }
    if (proof.final_poly.size() != (size_t(1) << fp.log_final_poly_len)) return false;
    for (auto& c : proof.final_poly) ch.observe(c);
    if (proof.query_proofs.size() != fp.num_queries) return false;
    if (!ch.check_witness(fp.query_proof_of_work_bits, proof.query_pow_witness)) return false;

    const unsigned log_max_height = (unsigned)(proof.commit_phase_commits.size() + log_blowup + fp.log_final_poly_len);
    const unsigned log_final_height = (unsigned)(log_blowup + fp.log_final_poly_len);

    for (const QueryProof& qp : proof.query_proofs) {
        size_t index = ch.sample_bits(log_max_height);
        // ---- open_input: check the input openings and build the reduced openings per height ----
        if (qp.input_proof.size() != rounds.size()) return false;
        std::map<unsigned, std::pair<Fp2, Fp2>> reduced;  // log_height -> (alpha_pow, ro)
        for (size_t r = 0; r < rounds.size(); r++) {
            const msh::BatchOpening& bo = qp.input_proof[r];
            const VerifierRound& round = rounds[r];
            if (bo.opened_values.size() != round.mats.size()) return false;
            std::vector<size_t> heights;
            unsigned log_batch_max = 0;
            for (auto& m : round.mats) {
                heights.push_back(size_t(1) << (m.log_degree + log_blowup));
                log_batch_max = std::max(log_batch_max, m.log_degree + log_blowup);
            }
            if (log_batch_max > log_max_height) return false;
            if (!round.mats.empty()) {
                size_t reduced_index = index >> (log_max_height - log_batch_max);
                if (!verify_batch(round.commit, heights, reduced_index, bo.opened_values, bo.opening_proof)) return false;
            }
            for (size_t mi = 0; mi < round.mats.size(); mi++) {
                const VerifierMat& m = round.mats[mi];
                unsigned log_height = m.log_degree + log_blowup;
                size_t rev_reduced = reverse_bits_len(index >> (log_max_height - log_height), log_height);
                Fp x = Fp(GL_GENERATOR) * two_adic_generator(log_height).pow((u64)rev_reduced);
                auto it = reduced.find(log_height);
                if (it == reduced.end()) it = reduced.emplace(log_height, std::make_pair(Fp2::one(), Fp2::zero())).first;
                Fp2& alpha_pow = it->second.first;
                Fp2& ro = it->second.second;
                for (auto& pt : m.points) {
                    if (pt.second.size() != bo.opened_values[mi].size()) return false;
                    Fp2 zx = pt.first - x;
                    if (zx.is_zero()) return false;
                    Fp2 quotient = zx.inverse();
                    for (size_t c = 0; c < pt.second.size(); c++) {
                        ro += alpha_pow * (pt.second[c] - Fp2(bo.opened_values[mi][c])) * quotient;
                        alpha_pow *= alpha;
                    }
                }
            }
            auto small = reduced.find(log_blowup);
            if (small != reduced.end()) {
                if (!small->second.second.is_zero()) return false;
                reduced.erase(small);
            }
        }
        std::vector<std::pair<unsigned, Fp2>> ro_desc;
        for (auto it = reduced.rbegin(); it != reduced.rend(); ++it) ro_desc.push_back({it->first, it->second.second});

        // ---- verify_query ----
        if (qp.commit_phase_openings.size() != proof.commit_phase_commits.size()) return false;
        size_t ro_pos = 0;
        if (ro_desc.empty() || ro_desc[0].first != log_max_height) return false;
        Fp2 folded = ro_desc[ro_pos++].second;
        size_t domain_index = index;
        for (size_t k = 0; k < proof.commit_phase_commits.size(); k++) {
            unsigned log_folded_height = log_max_height - 1 - (unsigned)k;
            const CommitPhaseProofStep& step = qp.commit_phase_openings[k];
            if (step.log_arity != 1 || step.sibling_values.size() != 1) return false;
            size_t sibling = domain_index ^ 1;
            std::vector<Fp2> evals(2, folded);
            evals[sibling % 2] = step.sibling_values[0];
            domain_index >>= 1;
            std::vector<std::vector<Fp>> row(1);
            for (auto& e : evals) { row[0].push_back(e.c[0]); row[0].push_back(e.c[1]); }
            if (!verify_batch(proof.commit_phase_commits[k], {size_t(1) << log_folded_height}, domain_index, row, step.opening_proof))
                return false;
            folded = fold_row(domain_index, log_folded_height, betas[k], evals[0], evals[1]);
            if (ro_pos < ro_desc.size() && ro_desc[ro_pos].first == log_folded_height) folded += betas[k].square() * ro_desc[ro_pos++].second;
        }
        if (ro_pos != ro_desc.size()) return false;  // an input that was never rolled in
        (void)log_final_height;
        Fp x = two_adic_generator(log_max_height).pow((u64)reverse_bits_len(domain_index, log_max_height));
        Fp2 eval = Fp2::zero();
        for (size_t i = proof.final_poly.size(); i-- > 0;) eval = eval * x + proof.final_poly[i];
        if (eval != folded) return false;
    }
    return true;
}

// src/verifier.rs:536-695. Fills the quotient degrees per active circuit.
inline VerifyError verify_shape(const SystemShape& sys, const Proof& proof, std::vector<size_t>& quotient_degrees) {
    size_t num_circuits = sys.circuits.size();
    if (num_circuits == 0) return VerifyError::InvalidSystem;
    if (proof.active.size() != num_circuits) return VerifyError::InvalidProofShape;
    std::vector<size_t> active_indices;
    for (size_t i = 0; i < proof.active.size(); i++)
        if (proof.active[i]) active_indices.push_back(i);
    size_t num_active = active_indices.size();
    if (num_active == 0) return VerifyError::InvalidProofShape;
    if (proof.log_degrees.size() != num_active) return VerifyError::InvalidProofShape;
    size_t num_pre = sys.num_preprocessed;
    size_t got_pre = proof.preprocessed_opened_values ? proof.preprocessed_opened_values->size() : 0;
    if (got_pre != num_pre) return VerifyError::InvalidProofShape;
    for (size_t ci = 0; ci < num_circuits; ci++)
        if (sys.preprocessed_indices[ci] >= 0 && !proof.active[ci])
            if (!(*proof.preprocessed_opened_values)[sys.preprocessed_indices[ci]].empty()) return VerifyError::InvalidProofShape;
    if (proof.stage_1_opened_values.size() != num_active || proof.stage_2_opened_values.size() != num_active)
        return VerifyError::InvalidProofShape;
    for (size_t pos = 0; pos < num_active; pos++) {
        const Circuit& c = sys.circuits[active_indices[pos]];
        int slot = sys.preprocessed_indices[active_indices[pos]];
        if (proof.stage_1_opened_values[pos].size() != 2 || proof.stage_2_opened_values[pos].size() != 2) return VerifyError::InvalidProofShape;
        if (slot >= 0 && (*proof.preprocessed_opened_values)[slot].size() != 2) return VerifyError::InvalidProofShape;
        for (int j = 0; j < 2; j++) {
            if (slot >= 0 && (*proof.preprocessed_opened_values)[slot][j].size() != c.preprocessed_width) return VerifyError::InvalidProofShape;
            if (proof.stage_1_opened_values[pos][j].size() != c.main_width) return VerifyError::InvalidProofShape;
            if (proof.stage_2_opened_values[pos][j].size() != c.stage_2_width) return VerifyError::InvalidProofShape;
        }
    }
    quotient_degrees.clear();
    size_t max_log_degree = GL_TWO_ADICITY - sys.commitment.log_blowup;  // src/types.rs:131
    for (size_t pos = 0; pos < num_active; pos++) {
        size_t qd = sys.circuits[active_indices[pos]].quotient_degree();
        if ((size_t)proof.log_degrees[pos] + log2_strict(qd) > max_log_degree) return VerifyError::InvalidProofShape;
        quotient_degrees.push_back(qd);
    }
    if (proof.quotient_opened_values.size() != num_active) return VerifyError::InvalidProofShape;
    for (size_t pos = 0; pos < num_active; pos++) {
        if (proof.quotient_opened_values[pos].size() != 1) return VerifyError::InvalidProofShape;
        if (proof.quotient_opened_values[pos][0].size() != quotient_degrees[pos] * 2) return VerifyError::InvalidProofShape;
    }
    if (proof.intermediate_accumulators.size() != num_active) return VerifyError::InvalidProofShape;
    return VerifyError::Ok;
}

// src/verifier.rs:208-532. `preprocessed_commit` is the verifier key (System::new, src/system.rs:193-196).
inline VerifyError verify_multiple_claims(const SystemShape& sys, bool has_preprocessed_commit, const Digest& preprocessed_commit,
                                          const std::vector<std::vector<Fp>>& claims, const Proof& proof) {
    std::vector<size_t> quotient_degrees;
    if ((sys.num_preprocessed == 0) != !has_preprocessed_commit) return VerifyError::InvalidSystem;
    VerifyError se = verify_shape(sys, proof, quotient_degrees);
    if (se != VerifyError::Ok) return se;
    std::vector<size_t> active_indices;
    for (size_t i = 0; i < proof.active.size(); i++)
        if (proof.active[i]) active_indices.push_back(i);
    if (!(proof.intermediate_accumulators.back() == Fp2::zero())) return VerifyError::UnbalancedChannel;

    Challenger ch = Challenger::for_config(sys.commitment, sys.fri);
    sys.observe_shape(ch);
    for (bool a : proof.active) ch.observe(Fp(a ? 1 : 0));
    if (has_preprocessed_commit) ch.observe(preprocessed_commit);
    ch.observe(proof.stage_1_trace);
    for (u8 ld : proof.log_degrees) ch.observe(Fp((u64)ld));
    ch.observe_usize(claims.size());
    for (auto& claim : claims) {
        ch.observe_usize(claim.size());
        ch.observe_slice(claim.data(), claim.size());
    }
    Fp2 beta = ch.sample_ext();
    ch.observe(beta);
    Fp2 gamma = ch.sample_ext();
    ch.observe(gamma);
    ch.observe(proof.stage_2_trace);
    for (auto& a : proof.intermediate_accumulators) ch.observe(a);
    Fp2 acc = Fp2::zero();
    for (auto& claim : claims) acc += (beta + fingerprint(gamma, claim.data(), claim.size())).inverse();
    Fp2 alpha = ch.sample_ext();
    ch.observe(proof.quotient_chunks);
    Fp2 zeta = ch.sample_ext();

    std::vector<VerifierRound> rounds(3);
    rounds[0].commit = proof.stage_1_trace;
    rounds[1].commit = proof.stage_2_trace;
    rounds[2].commit = proof.quotient_chunks;
    for (size_t pos = 0; pos < active_indices.size(); pos++) {
        unsigned ld = proof.log_degrees[pos];
        Fp2 zeta_next = zeta * two_adic_generator(ld);
        rounds[0].mats.push_back(VerifierMat{ld, {{zeta, proof.stage_1_opened_values[pos][0]}, {zeta_next, proof.stage_1_opened_values[pos][1]}}});
        rounds[1].mats.push_back(VerifierMat{ld, {{zeta, proof.stage_2_opened_values[pos][0]}, {zeta_next, proof.stage_2_opened_values[pos][1]}}});
        rounds[2].mats.push_back(VerifierMat{ld, {{zeta, proof.quotient_opened_values[pos][0]}}});
    }
    if (has_preprocessed_commit) {
        VerifierRound r0;
        r0.commit = preprocessed_commit;
        std::vector<int> active_pos(proof.active.size(), -1);
        for (size_t pos = 0; pos < active_indices.size(); pos++) active_pos[active_indices[pos]] = (int)pos;
        for (size_t ci = 0; ci < sys.circuits.size(); ci++) {
            int slot = sys.preprocessed_indices[ci];
            if (slot < 0) continue;
            if (active_pos[ci] >= 0) {
                unsigned ld = proof.log_degrees[active_pos[ci]];
                Fp2 zeta_next = zeta * two_adic_generator(ld);
                const auto& pv = (*proof.preprocessed_opened_values)[slot];
                r0.mats.push_back(VerifierMat{ld, {{zeta, pv[0]}, {zeta_next, pv[1]}}});
            } else {
                r0.mats.push_back(VerifierMat{log2_strict(sys.circuits[ci].preprocessed_height), {}});
            }
        }
        rounds.push_back(std::move(r0));
    }
    if (!pcs_verify(rounds, proof.opening_proof, sys.commitment, sys.fri, ch)) return VerifyError::InvalidOpeningArgument;

    for (size_t pos = 0; pos < active_indices.size(); pos++) {
        size_t ci = active_indices[pos];
        const Circuit& circuit = sys.circuits[ci];
        unsigned ld = proof.log_degrees[pos];
        size_t degree = size_t(1) << ld;
        Fp2 next_acc = proof.intermediate_accumulators[pos];
        SelectorsAtPoint sels = selectors_at_point(ld, zeta);
        Fp inj_norm = (Fp((u64)degree) * two_adic_generator(ld)).inverse();
        std::vector<Fp2> publics;
        for (const Fp2& ef : {beta, gamma, acc, next_acc}) { publics.push_back(Fp2(ef.c[0])); publics.push_back(Fp2(ef.c[1])); }
        static const std::vector<Fp2> empty;
        int slot = sys.preprocessed_indices[ci];
        VarValues<Fp2> view{};
        view.preprocessed[0] = slot >= 0 ? (*proof.preprocessed_opened_values)[slot][0].data() : empty.data();
        view.preprocessed[1] = slot >= 0 ? (*proof.preprocessed_opened_values)[slot][1].data() : empty.data();
        view.main[0] = proof.stage_1_opened_values[pos][0].data();
        view.main[1] = proof.stage_1_opened_values[pos][1].data();
        view.stage2[0] = proof.stage_2_opened_values[pos][0].data();
        view.stage2[1] = proof.stage_2_opened_values[pos][1].data();
        view.publics = publics.data();
        view.is_first_row = sels.is_first_row;
        view.is_last_row = sels.is_last_row;
        view.is_transition = sels.is_transition;
        std::vector<Fp2> buf, cv;
        sweep_range(circuit.graph, view, buf, circuit.graph.nodes.size());
        for (u32 z : circuit.graph.zeros) cv.push_back(buf[z]);
        Fp2 delta_scaled[2] = {(publics[6] - publics[4]) * inj_norm, (publics[7] - publics[5]) * inj_norm};
        logup_constraint_values<Fp2>(circuit.graph.lookups, buf, view.stage2[0], view.stage2[1], publics.data(), delta_scaled,
                                     view.is_last_row, Fp(GL_EXT_W), cv);
        if (cv.size() != circuit.constraint_count) return VerifyError::InvalidProofShape;
        Fp2 composition = Fp2::zero();
        for (auto& v : cv) composition = composition * alpha + v;
        const std::vector<Fp2>& qrow = proof.quotient_opened_values[pos][0];
        Fp2 zeta_pow_n = zeta.exp_power_of_2(ld), zp = Fp2::one(), quotient = Fp2::zero();
        for (size_t k = 0; k + 1 < qrow.size(); k += 2) {
            // from_ext_basis: chunk[0] + chunk[1] * X
            Fp2 v = qrow[k] + qrow[k + 1] * Fp2(Fp::zero(), Fp::one());
            quotient += zp * v;
            zp *= zeta_pow_n;
        }
        if (composition * sels.inv_vanishing != quotient) return VerifyError::OodEvaluationMismatch;
        acc = next_acc;
    }
    return VerifyError::Ok;
}

}  // namespace orc
