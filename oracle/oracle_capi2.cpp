// ORACLE C entry points, part 2 (test infrastructure, not product code): systems, stage-2 traces, quotient
// values. Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's CPU legs ONLY.
#include "orc_system.hpp"
#include "../multi_stark_b200/host/system_from_graphs.hpp"
#include <cstring>
#include <memory>

using namespace orc;

static Matrix to_matrix2(const u64* in, u64 rows, u64 cols) {
    Matrix m(rows, cols);
    for (size_t i = 0; i < rows * cols; i++) m.values[i] = Fp(in[i]);
    return m;
}

extern "C" {

void* orc_system_create(const char* kind, u32 log_blowup, u32 log_final_poly_len, u32 max_log_arity, u32 num_queries,
                        u32 commit_pow_bits, u32 query_pow_bits) {
    try {
        CommitmentParameters cp;
        cp.log_blowup = log_blowup;
        FriParameters fp;
        fp.log_final_poly_len = log_final_poly_len;
        fp.max_log_arity = max_log_arity;
        fp.num_queries = num_queries;
        fp.commit_proof_of_work_bits = commit_pow_bits;
        fp.query_proof_of_work_bits = query_pow_bits;
        auto s = std::make_unique<OrcSystem>();
        s->shape = SystemShape::build(cp, fp, named_system_inputs(kind));
        return s.release();
    } catch (const std::exception&) {
        return nullptr;
    }
}
// the oracle's System from compiled circuits (same descriptors as msh_system_create_from_graphs)
void* orc_system_create_from_graphs(const msgpu_graph_desc* descs, u32 n, const u64* const* preprocessed, const u64* pre_heights,
                                    u32 log_blowup, u32 log_final_poly_len, u32 max_log_arity, u32 num_queries, u32 commit_pow_bits,
                                    u32 query_pow_bits) {
    try {
        CommitmentParameters cp;
        cp.log_blowup = log_blowup;
        FriParameters fp;
        fp.log_final_poly_len = log_final_poly_len;
        fp.max_log_arity = max_log_arity;
        fp.num_queries = num_queries;
        fp.commit_proof_of_work_bits = commit_pow_bits;
        fp.query_proof_of_work_bits = query_pow_bits;
        auto s = std::make_unique<OrcSystem>();
        s->shape = system_from_descs(cp, fp, descs, n, (const uint64_t* const*)preprocessed, (const uint64_t*)pre_heights);
        return s.release();
    } catch (const std::exception&) {
        return nullptr;
    }
}
void orc_system_free(void* s) { delete (OrcSystem*)s; }

// Node listing of a compiled circuit, for structural tests: op, a, b, imm per node (Var: a = source | offset << 2,
// b = column); zeros; lookups flattened.
u64 orc_graph_num_nodes(void* s, u32 ci) { return ((OrcSystem*)s)->shape.circuits[ci].graph.nodes.size(); }
void orc_graph_nodes(void* s, u32 ci, u8* op, u32* a, u32* b, u64* imm, u32* degrees) {
    const ConstraintGraph& g = ((OrcSystem*)s)->shape.circuits[ci].graph;
    for (size_t i = 0; i < g.nodes.size(); i++) {
        const Node& n = g.nodes[i];
        op[i] = (u8)n.op; a[i] = n.a; b[i] = n.b; imm[i] = n.c.v; degrees[i] = g.degrees[i];
        if (n.op == Op::Var) { a[i] = (u32)n.col.source | ((u32)n.col.offset << 2); b[i] = n.col.index; }
    }
}
u64 orc_graph_num_zeros(void* s, u32 ci) { return ((OrcSystem*)s)->shape.circuits[ci].graph.zeros.size(); }
void orc_graph_zeros(void* s, u32 ci, u32* out) {
    const ConstraintGraph& g = ((OrcSystem*)s)->shape.circuits[ci].graph;
    for (size_t i = 0; i < g.zeros.size(); i++) out[i] = g.zeros[i];
}

// compute_lookup_values + stage_2_traces for ONE circuit starting from a zero accumulator.
void orc_stage2_trace(void* s, u32 ci, const u64* main, u64 rows, const u64* beta2, const u64* gamma2, u64* out, u64* local_sum2) {
    const Circuit& c = ((OrcSystem*)s)->shape.circuits[ci];
    Matrix trace = to_matrix2(main, rows, c.main_width);
    LookupValues lv = compute_lookup_values(c, trace);
    std::vector<Matrix> traces;
    std::vector<Fp2> inter;
    stage_2_traces({&lv}, Fp2(Fp(beta2[0]), Fp(beta2[1])), Fp2(Fp(gamma2[0]), Fp(gamma2[1])), Fp2::zero(), traces, inter);
    for (size_t i = 0; i < traces[0].values.size(); i++) out[i] = traces[0].values[i].v;
    local_sum2[0] = inter[0].c[0].v;
    local_sum2[1] = inter[0].c[1].v;
}

// src/prover.rs:381-387
void orc_claims_accumulator(const u64* claims, u64 n, u64 len, const u64* beta2, const u64* gamma2, u64* out2) {
    Fp2 beta{Fp(beta2[0]), Fp(beta2[1])}, gamma{Fp(gamma2[0]), Fp(gamma2[1])}, acc = Fp2::zero();
    std::vector<Fp> tmp(len);
    for (u64 i = 0; i < n; i++) {
        for (u64 k = 0; k < len; k++) tmp[k] = Fp(claims[i * len + k]);
        acc += (beta + fingerprint(gamma, tmp.data(), len)).inverse();
    }
    out2[0] = acc.c[0].v;
    out2[1] = acc.c[1].v;
}

// quotient_values from the committed LDEs (stored bit-reversed, at least nq rows each). out: nq x 2, natural order.
void orc_quotient_values(void* s, u32 ci, u32 log_n, u32 log_q, const u64* pre_lde, const u64* s1_lde, const u64* s2_lde,
                         const u64* publics8, const u64* alpha2, u64* out) {
    const Circuit& c = ((OrcSystem*)s)->shape.circuits[ci];
    size_t nq = size_t(1) << (log_n + log_q);
    Matrix pre = pre_lde ? to_matrix2(pre_lde, nq, c.preprocessed_width) : Matrix();
    Matrix s1 = to_matrix2(s1_lde, nq, c.main_width), s2 = to_matrix2(s2_lde, nq, c.stage_2_width);
    DomainView vp{pre.values.data(), c.preprocessed_width, log_n + log_q};
    DomainView v1{s1.values.data(), c.main_width, log_n + log_q}, v2{s2.values.data(), c.stage_2_width, log_n + log_q};
    Fp pub[8];
    for (int i = 0; i < 8; i++) pub[i] = Fp(publics8[i]);
    auto q = quotient_values(c, pub, log_n, log_q, pre_lde ? &vp : nullptr, v1, v2, Fp2(Fp(alpha2[0]), Fp(alpha2[1])));
    for (size_t i = 0; i < nq; i++) { out[2 * i] = q[i].c[0].v; out[2 * i + 1] = q[i].c[1].v; }
}

// selectors_on_coset for the pinning test (src/lookup.rs:697-756). out: 4 vectors of n << rate_bits, natural order.
void orc_selectors_on_coset(u32 log_n, u32 rate_bits, u64 shift, u64* first, u64* last, u64* trans, u64* inv_van) {
    auto sel = selectors_on_coset(log_n, rate_bits, Fp(shift));
    for (size_t i = 0; i < sel.is_first_row.size(); i++) {
        first[i] = sel.is_first_row[i].v; last[i] = sel.is_last_row[i].v;
        trans[i] = sel.is_transition[i].v; inv_van[i] = sel.inv_vanishing[i].v;
    }
}

}  // extern "C"
