// ORACLE (test infrastructure, not product code): checks the product's bytecode lowering (multi_stark_b200/csrc/lowering.hpp:
// liveness-based slot reuse, leaves materialised at first use, OP_ROOT folds) on the CPU by interpreting the lowered program
// next to a direct node-by-node evaluation of the same ConstraintGraph (the reference's sweep, src/eval.rs:67-106) on random
// rows. The device interpreter (quotient.cu) executes exactly this bytecode; its arithmetic is covered by the GPU parity tests.
#include "orc_system.hpp"
#include "../multi_stark_b200/csrc/lowering.hpp"
#include "../multi_stark_b200/host/program.hpp"
#include <random>

using namespace msh;

namespace {
struct RowValues {
    std::vector<Fp> rows[3][2];  // [source][offset]
    Fp publics[8];
    Fp first, last, trans;
};

Fp leaf_value(uint32_t op, uint32_t a, uint32_t b, uint64_t imm, const RowValues& rv) {
    switch (op) {
        case msg::OP_CONST: return Fp(imm);
        case msg::OP_VAR: return rv.rows[a & 3][a >> 2][b];
        case msg::OP_PUBLIC: return rv.publics[a];
        case msg::OP_FIRST: return rv.first;
        case msg::OP_LAST: return rv.last;
        default: return rv.trans;
    }
}

// returns the number of mismatches between the lowered program and the graph on `trials` random rows
uint64_t check_program(const msgpu_graph_desc& g, uint32_t len, const std::vector<uint32_t>& pinned, const std::vector<uint32_t>& roots,
                       uint32_t trials, uint64_t seed, uint32_t& n_slots, uint32_t& n_instr) {
    msg::Lowered low = msg::lower(g, len, pinned, roots);
    n_slots = low.n_slots;
    n_instr = (uint32_t)low.code.size();
    std::mt19937_64 rng(seed);
    auto rnd = [&] { return Fp(rng() % GL_P); };
    uint64_t bad = 0;
    for (uint32_t t = 0; t < trials; t++) {
        RowValues rv;
        const uint32_t widths[3] = {g.pre_width, g.main_width, g.stage2_width};
        for (int s = 0; s < 3; s++)
            for (int o = 0; o < 2; o++) {
                rv.rows[s][o].resize(widths[s]);
                for (auto& v : rv.rows[s][o]) v = rnd();
            }
        for (auto& p : rv.publics) p = rnd();
        rv.first = rnd(); rv.last = rnd(); rv.trans = rnd();
        // direct evaluation: one value per node
        std::vector<Fp> buf(len);
        for (uint32_t i = 0; i < len; i++) {
            uint32_t op = g.op[i];
            if (op == msg::OP_ADD) buf[i] = buf[g.a[i]] + buf[g.b[i]];
            else if (op == msg::OP_SUB) buf[i] = buf[g.a[i]] - buf[g.b[i]];
            else if (op == msg::OP_MUL) buf[i] = buf[g.a[i]] * buf[g.b[i]];
            else if (op == msg::OP_NEG) buf[i] = Fp::zero() - buf[g.a[i]];
            else buf[i] = leaf_value(op, g.a[i], g.b[i], g.imm[i], rv);
        }
        // the lowered program
        std::vector<Fp> slots(low.n_slots);
        std::vector<Fp> root_val(roots.size());
        std::vector<char> root_seen(roots.size(), 0);
        for (const msg::Instr& in : low.code) {
            if (in.op == msg::OP_ROOT) {
                if (in.imm >= roots.size() || root_seen[in.imm]) { bad++; continue; }
                root_seen[in.imm] = 1;
                root_val[in.imm] = slots[in.a];
                continue;
            }
            Fp v;
            if (in.op == msg::OP_ADD) v = slots[in.a] + slots[in.b];
            else if (in.op == msg::OP_SUB) v = slots[in.a] - slots[in.b];
            else if (in.op == msg::OP_MUL) v = slots[in.a] * slots[in.b];
            else if (in.op == msg::OP_NEG) v = Fp::zero() - slots[in.a];
            else v = leaf_value(in.op, in.a, in.b, in.imm, rv);
            if (in.dst >= low.n_slots) { bad++; continue; }
            slots[in.dst] = v;
        }
        for (size_t j = 0; j < roots.size(); j++)
            if (!root_seen[j] || !(root_val[j] == buf[roots[j]])) bad++;
        for (uint32_t p : pinned)  // pinned nodes (lookup multiplicities and arguments) must survive to the end
            if (low.slot_of[p] >= low.n_slots || !(slots[low.slot_of[p]] == buf[p])) bad++;
    }
    return bad;
}
}  // namespace

extern "C" {

// out4 = slots and instructions of the full program, slots and instructions of the lookup prefix. Returns mismatches (0 = ok),
// or ~0 on an exception.
uint64_t orc_check_lowering(void* system, uint32_t circuit, uint32_t trials, uint64_t seed, uint32_t* out4) {
    try {
        OrcSystem& sys = *(OrcSystem*)system;
        const Circuit& c = sys.shape.circuits.at(circuit);
        GraphDesc d;
        d.build(c.graph, c.preprocessed_width, c.main_width, c.stage_2_width);
        const msgpu_graph_desc& g = d.desc;
        std::vector<uint32_t> pin_lk, roots(g.zeros, g.zeros + g.n_zeros);
        uint32_t n_args = g.n_lookups ? g.lookup_arg_off[g.n_lookups] : 0;
        for (uint32_t j = 0; j < g.n_lookups; j++) pin_lk.push_back(g.lookup_mult[j]);
        for (uint32_t k = 0; k < n_args; k++) pin_lk.push_back(g.lookup_args[k]);
        uint64_t bad = check_program(g, g.n_nodes, pin_lk, roots, trials, seed, out4[0], out4[1]);
        bad += check_program(g, g.lookup_prefix_len, pin_lk, {}, trials, seed + 1, out4[2], out4[3]);
        return bad;
    } catch (const std::exception&) {
        return ~0ull;
    }
}

}  // extern "C"
