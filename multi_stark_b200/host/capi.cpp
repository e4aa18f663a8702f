// libmshost.so: C entry points of the host-side protocol layer (system assembly, workloads and -- in
// prover.hpp -- the prove() driver over the libmsgpu C ABI). Python (tests, bench.py) binds these with ctypes.
#include "system.hpp"
#include "program.hpp"
#include <cstring>
#include <memory>
#include <string>

using namespace msh;

struct msh_system {
    SystemShape shape;
    std::vector<std::unique_ptr<GraphDesc>> descs;
};

static thread_local std::string g_err;

extern "C" {

const char* msh_last_error() { return g_err.c_str(); }

msh_system* msh_system_create(const char* kind, uint32_t log_blowup, uint32_t log_final_poly_len, uint32_t max_log_arity,
                              uint32_t num_queries, uint32_t commit_pow_bits, uint32_t query_pow_bits) {
    try {
        CommitmentParameters cp;
        cp.log_blowup = log_blowup;
        FriParameters fp;
        fp.log_final_poly_len = log_final_poly_len;
        fp.max_log_arity = max_log_arity;
        fp.num_queries = num_queries;
        fp.commit_proof_of_work_bits = commit_pow_bits;
        fp.query_proof_of_work_bits = query_pow_bits;
        auto sys = std::make_unique<msh_system>();
        sys->shape = SystemShape::build(cp, fp, named_system_inputs(kind));
        for (auto& c : sys->shape.circuits) {
            auto d = std::make_unique<GraphDesc>();
            d->build(c.graph, c.preprocessed_width, c.main_width, c.stage_2_width);
            sys->descs.push_back(std::move(d));
        }
        return sys.release();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void msh_system_free(msh_system* s) { delete s; }
uint32_t msh_system_num_circuits(const msh_system* s) { return (uint32_t)s->shape.circuits.size(); }
// out[12]: main_width, pre_width, pre_height, num_lookups, stage2_width, constraint_count, max_constraint_degree,
//          quotient_degree, n_nodes, n_zeros, lookup_prefix_len, preprocessed_index (or ~0)
void msh_circuit_info(const msh_system* s, uint32_t i, uint64_t* out) {
    const Circuit& c = s->shape.circuits[i];
    out[0] = c.main_width; out[1] = c.preprocessed_width; out[2] = c.preprocessed_height; out[3] = c.num_lookups;
    out[4] = c.stage_2_width; out[5] = c.constraint_count; out[6] = c.max_constraint_degree; out[7] = c.quotient_degree();
    out[8] = c.graph.nodes.size(); out[9] = c.graph.zeros.size(); out[10] = c.graph.lookup_prefix_len;
    out[11] = s->shape.preprocessed_indices[i] < 0 ? ~0ull : (uint64_t)s->shape.preprocessed_indices[i];
}
const msgpu_graph_desc* msh_circuit_graph(const msh_system* s, uint32_t i) { return &s->descs[i]->desc; }
void msh_circuit_preprocessed(const msh_system* s, uint32_t i, uint64_t* out) {
    const Circuit& c = s->shape.circuits[i];
    for (size_t k = 0; k < c.preprocessed.values.size(); k++) out[k] = c.preprocessed.values[k].v;
}

// benches/multi_stark.rs:171-238. add_trace_out: next_pow2(num_adds) x 14; claims_out: num_adds x 4
void msh_u32add_workload(uint64_t num_adds, uint64_t* byte_trace_out, uint64_t* add_trace_out, uint64_t* claims_out) {
    auto w = circuits::u32_add_workload(num_adds);
    for (size_t i = 0; i < 256; i++) byte_trace_out[i] = w.byte_trace.values[i].v;
    for (size_t i = 0; i < w.add_trace.values.size(); i++) add_trace_out[i] = w.add_trace.values[i].v;
    if (claims_out)
        for (size_t i = 0; i < w.claims.size(); i++)
            for (int k = 0; k < 4; k++) claims_out[4 * i + k] = w.claims[i][k].v;
}
void msh_fib_trace(uint64_t rows, uint64_t* out) {
    Matrix m = circuits::fib_cubic_trace(rows);
    for (size_t i = 0; i < m.values.size(); i++) out[i] = m.values[i].v;
}

}  // extern "C"
