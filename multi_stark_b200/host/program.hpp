// ConstraintGraph -> msgpu_graph_desc (the flat arrays the C ABI takes, include/msgpu.h).
#pragma once
#include "expr.hpp"
#include "../../include/msgpu.h"

namespace msh {

struct GraphDesc {
    std::vector<uint8_t> op;
    std::vector<uint32_t> a, b, zeros, lookup_mult, lookup_arg_off, lookup_args;
    std::vector<uint64_t> imm;
    msgpu_graph_desc desc{};

    GraphDesc() = default;
    GraphDesc(const GraphDesc&) = delete;
    GraphDesc& operator=(const GraphDesc&) = delete;

    void build(const ConstraintGraph& g, size_t pre_width, size_t main_width, size_t stage2_width) {
        size_t n = g.nodes.size();
        op.resize(n); a.assign(n, 0); b.assign(n, 0); imm.assign(n, 0);
        for (size_t i = 0; i < n; i++) {
            const Node& nd = g.nodes[i];
            op[i] = (uint8_t)nd.op;
            switch (nd.op) {
                case Op::Const: imm[i] = nd.c.v; break;
                case Op::Var: a[i] = (uint32_t)nd.col.source | ((uint32_t)nd.col.offset << 2); b[i] = nd.col.index; break;
                case Op::Public: a[i] = nd.a; break;
                case Op::Add: case Op::Sub: case Op::Mul: a[i] = nd.a; b[i] = nd.b; break;
                case Op::Neg: a[i] = nd.a; break;
                default: break;
            }
        }
        zeros.assign(g.zeros.begin(), g.zeros.end());
        lookup_mult.clear(); lookup_args.clear(); lookup_arg_off.assign(1, 0);
        for (auto& l : g.lookups) {
            lookup_mult.push_back(l.multiplicity);
            for (u32 x : l.args) lookup_args.push_back(x);
            lookup_arg_off.push_back((uint32_t)lookup_args.size());
        }
        desc.n_nodes = (uint32_t)n;
        desc.op = op.data(); desc.a = a.data(); desc.b = b.data(); desc.imm = imm.data();
        desc.n_zeros = (uint32_t)zeros.size(); desc.zeros = zeros.data();
        desc.n_lookups = (uint32_t)g.lookups.size();
        desc.lookup_mult = lookup_mult.data(); desc.lookup_arg_off = lookup_arg_off.data(); desc.lookup_args = lookup_args.data();
        desc.lookup_prefix_len = (uint32_t)g.lookup_prefix_len;
        desc.pre_width = (uint32_t)pre_width; desc.main_width = (uint32_t)main_width; desc.stage2_width = (uint32_t)stage2_width;
    }
};

}  // namespace msh
