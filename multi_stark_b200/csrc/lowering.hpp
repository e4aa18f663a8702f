// Lowering of a compiled ConstraintGraph (include/msgpu.h: msgpu_graph_desc, the reference's `Node` vector of
// src/graph.rs:35-76) to the bytecode the device interpreter runs (quotient.cu). Pure host C++ with no CUDA dependency, so
// that a CPU test can interpret the lowered program against the graph itself (tests/test_lowering.py).
#pragma once
#include "../../include/msgpu.h"
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace msg {

typedef unsigned long long u64;
typedef unsigned int u32;

struct LowerError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
#define LOWER_REQUIRE(cond, msg_)                         \
    do {                                                  \
        if (!(cond)) throw msg::LowerError(msg_);         \
    } while (0)

constexpr u64 kLowerP = 0xFFFFFFFF00000001ull;

enum : u32 { OP_CONST = 0, OP_VAR = 1, OP_PUBLIC = 2, OP_FIRST = 3, OP_LAST = 4, OP_TRANS = 5, OP_ADD = 6, OP_SUB = 7, OP_MUL = 8, OP_NEG = 9,
             OP_ROOT = 10 };  // OP_ROOT (lowering only): constraint `imm` has the value of slot `a`; folded at once, the slot is free again

struct alignas(16) Instr {  // 32 bytes: two 16-byte uniform loads per instruction
    u32 op, dst, a, b;
    u64 imm;
    u64 pad;
};

struct Lowered {
    std::vector<Instr> code;
    std::vector<u32> slot_of;  // node id -> slot (only meaningful for pinned nodes after the run)
    u32 n_slots = 0;
};

// Lower nodes [0, len) keeping `pinned` nodes alive to the end. Dead nodes are dropped. `roots[j]` = node of constraint j:
// an OP_ROOT instruction follows the node's own, so that the kernel folds the value into the alpha accumulators at once and
// the slot dies with the node's last real use (a wide AIR has hundreds of roots: kept alive to the end they made the
// interpreter's working set 2 KB per thread, k_quotient_eval<256>, issue rate 32 %).
inline Lowered lower(const msgpu_graph_desc& g, u32 len, const std::vector<u32>& pinned, const std::vector<u32>& roots = {}) {
    const u32 NONE = 0xffffffffu;
    std::vector<u32> last_use(len, NONE);
    std::vector<char> pin(len, 0), live(len, 0);
    for (u32 p : pinned) {
        LOWER_REQUIRE(p < len, "program: pinned node outside the evaluated range");
        pin[p] = 1;
        live[p] = 1;
    }
    std::vector<std::vector<u32>> root_ids(len);
    for (u32 j = 0; j < roots.size(); j++) {
        LOWER_REQUIRE(roots[j] < len, "program: constraint root outside the evaluated range");
        live[roots[j]] = 1;
        root_ids[roots[j]].push_back(j);
    }
    auto nchildren = [&](u32 i) -> int {
        uint8_t op = g.op[i];
        if (op == OP_ADD || op == OP_SUB || op == OP_MUL) return 2;
        if (op == OP_NEG) return 1;
        return 0;
    };
    for (u32 i = len; i-- > 0;) {  // liveness, children have smaller ids
        if (!live[i]) continue;
        int nc = nchildren(i);
        if (nc >= 1) {
            LOWER_REQUIRE(g.a[i] < i, "program: nodes are not topologically ordered");
            live[g.a[i]] = 1;
            if (last_use[g.a[i]] == NONE) last_use[g.a[i]] = i;
        }
        if (nc == 2) {
            LOWER_REQUIRE(g.b[i] < i, "program: nodes are not topologically ordered");
            live[g.b[i]] = 1;
            if (last_use[g.b[i]] == NONE) last_use[g.b[i]] = i;
        }
    }
    Lowered out;
    out.slot_of.assign(len, NONE);
    std::vector<u32> free_slots;
    // Leaves (constants, column reads, publics, selectors) are materialised at their FIRST USE, not at their node id: the
    // graph interns every column read up front, which would keep all of them alive at once.
    auto emit = [&](u32 i) {
        Instr in{};
        in.op = g.op[i];
        LOWER_REQUIRE(in.op <= OP_NEG, "program: bad opcode");
        int nc = nchildren(i);
        if (nc >= 1) in.a = out.slot_of[g.a[i]];
        if (nc == 2) in.b = out.slot_of[g.b[i]];
        if (in.op == OP_CONST) {
            LOWER_REQUIRE(g.imm[i] < kLowerP, "program: constant is not canonical");
            in.imm = g.imm[i];
        }
        if (in.op == OP_VAR) {
            u32 src = g.a[i] & 3, width = src == 0 ? g.pre_width : src == 1 ? g.main_width : g.stage2_width;
            LOWER_REQUIRE(src <= 2 && (g.a[i] >> 2) <= 1 && g.b[i] < width, "program: column reference out of range");
            in.a = g.a[i];
            in.b = g.b[i];
        }
        if (in.op == OP_PUBLIC) {
            LOWER_REQUIRE(g.a[i] < 8, "program: public index out of range");
            in.a = g.a[i];
        }
        // operands dying here release their slots first: the interpreter reads both operands before writing
        if (nc >= 1 && !pin[g.a[i]] && last_use[g.a[i]] == i) free_slots.push_back(out.slot_of[g.a[i]]);
        if (nc == 2 && g.b[i] != g.a[i] && !pin[g.b[i]] && last_use[g.b[i]] == i) free_slots.push_back(out.slot_of[g.b[i]]);
        u32 slot;
        if (!free_slots.empty()) {
            slot = free_slots.back();
            free_slots.pop_back();
        } else {
            slot = out.n_slots++;
        }
        out.slot_of[i] = slot;
        in.dst = slot;
        out.code.push_back(in);
        for (u32 j : root_ids[i]) {
            Instr r{};
            r.op = OP_ROOT;
            r.a = slot;
            r.imm = j;
            out.code.push_back(r);
        }
        // a root nobody reads later (and that no lookup needs) gives its slot back right after the fold
        if (!root_ids[i].empty() && !pin[i] && last_use[i] == NONE) free_slots.push_back(slot);
    };
    for (u32 i = 0; i < len; i++) {
        if (!live[i] || out.slot_of[i] != NONE) continue;
        const int nc = nchildren(i);
        if (nc == 0 && !pin[i] && root_ids[i].empty()) continue;  // lazy leaf
        if (nc >= 1 && out.slot_of[g.a[i]] == NONE) emit(g.a[i]);
        if (nc == 2 && out.slot_of[g.b[i]] == NONE) emit(g.b[i]);
        emit(i);
    }
    if (out.n_slots == 0) out.n_slots = 1;
    return out;
}

}  // namespace msg
