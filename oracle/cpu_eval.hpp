// ORACLE (test infrastructure, not product code): CPU restatement of the reference's constraint
// evaluation and lookup machinery:
//   sweep / sweep_lookup_prefix           src/eval.rs:51-106
//   logup_constraint_values               src/lookup.rs:152-256   (prover AND verifier share it)
//   compute_lookup_values                 src/system.rs:275-328
//   LookupValues::stage_2_traces          src/lookup.rs:472-555
//   selectors_on_coset / selectors_at_point   p3-commit TwoAdicMultiplicativeCoset (call src/prover.rs:775,
//                                         src/verifier.rs:455; pinned relation src/lookup.rs:697-756)
//   quotient_values                       src/prover.rs:756-962 (scalar instead of packed; exact arithmetic
//                                         makes the results identical)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this file.
#pragma once
#include "../multi_stark_b200/host/system.hpp"
#include "cpu_dft.hpp"

namespace orc {
using namespace msh;

// Working-type helpers: W is Fp (prover quotient domain, witness) or Fp2 (verifier at zeta).
inline Fp embed(Fp c, const Fp*) { return c; }
inline Fp2 embed(Fp c, const Fp2*) { return Fp2(c); }
inline Fp mul_w(Fp a, Fp w) { return a * w; }
inline Fp2 mul_w(const Fp2& a, Fp w) { return a * w; }

template <class W>
struct VarValues {
    const W* preprocessed[2];
    const W* main[2];
    const W* stage2[2];
    const W* publics;
    W is_first_row, is_last_row, is_transition;
    W var(const ColRef& col) const {
        const W* const* rows = col.source == Source::Preprocessed ? preprocessed : col.source == Source::Main ? main : stage2;
        return rows[(int)col.offset][col.index];
    }
};

// src/eval.rs:67-106
template <class W>
void sweep_range(const ConstraintGraph& g, const VarValues<W>& values, std::vector<W>& buf, size_t len) {
    buf.resize(len);
    for (size_t i = 0; i < len; i++) {
        const Node& n = g.nodes[i];
        switch (n.op) {
            case Op::Const: buf[i] = embed(n.c, (const W*)nullptr); break;
            case Op::Var: buf[i] = values.var(n.col); break;
            case Op::Public: buf[i] = values.publics[n.a]; break;
            case Op::IsFirstRow: buf[i] = values.is_first_row; break;
            case Op::IsLastRow: buf[i] = values.is_last_row; break;
            case Op::IsTransition: buf[i] = values.is_transition; break;
            case Op::Add: buf[i] = buf[n.a] + buf[n.b]; break;
            case Op::Sub: buf[i] = buf[n.a] - buf[n.b]; break;
            case Op::Mul: buf[i] = buf[n.a] * buf[n.b]; break;
            case Op::Neg: buf[i] = -buf[n.a]; break;
        }
    }
}

// src/lookup.rs:123-128
template <class A>
inline std::pair<A, A> mul2(std::pair<A, A> a, std::pair<A, A> b, Fp w) {
    A v0 = a.first * b.first, v1 = a.second * b.second;
    A cross = (a.first + a.second) * (b.first + b.second) - v0 - v1;
    return {v0 + mul_w(v1, w), cross};
}

// src/lookup.rs:152-208 (the D = 2 path; the reference's only production extension degree)
template <class A>
void logup_constraint_values(const std::vector<Lookup<u32>>& lookups, const std::vector<A>& node_vals, const A* stage2,
                             const A* stage2_next, const A* publics, const A* delta_scaled, A is_last_row, Fp w,
                             std::vector<A>& out) {
    std::pair<A, A> beta{publics[0], publics[1]}, gamma{publics[2], publics[3]};
    std::pair<A, A> inj{is_last_row * delta_scaled[0], is_last_row * delta_scaled[1]};
    if (lookups.empty()) {
        out.push_back(stage2_next[0] - stage2[0] + inj.first);
        out.push_back(stage2_next[1] - stage2[1] + inj.second);
        return;
    }
    size_t last = lookups.size() - 1;
    for (size_t j = 0; j < lookups.size(); j++) {
        std::pair<A, A> source{stage2[2 * j], stage2[2 * j + 1]};
        std::pair<A, A> target = j < last ? std::pair<A, A>{stage2[2 * j + 2], stage2[2 * j + 3]}
                                          : std::pair<A, A>{stage2_next[0] + inj.first, stage2_next[1] + inj.second};
        std::pair<A, A> f{A(), A()};
        for (size_t i = lookups[j].args.size(); i-- > 0;) {
            f = mul2(f, gamma, w);
            f.first = f.first + node_vals[lookups[j].args[i]];
        }
        auto c = mul2(std::pair<A, A>{f.first + beta.first, f.second + beta.second},
                      std::pair<A, A>{target.first - source.first, target.second - source.second}, w);
        out.push_back(c.first - node_vals[lookups[j].multiplicity]);
        out.push_back(c.second);
    }
}

// ---- lookup witness (src/lookup.rs:392-470, src/system.rs:275-328) ----------------------------
struct LookupValues {
    size_t height = 0, num_lookups = 0;
    std::vector<Fp> multiplicities;   // height * num_lookups
    std::vector<size_t> arg_offsets;  // num_lookups + 1
    std::vector<Fp> args;             // height * arg_offsets.back()
    const Fp* args_at(size_t row, size_t lookup) const { return args.data() + row * arg_offsets[num_lookups] + arg_offsets[lookup]; }
    size_t num_args(size_t lookup) const { return arg_offsets[lookup + 1] - arg_offsets[lookup]; }
};

inline LookupValues compute_lookup_values(const Circuit& c, const Matrix& trace) {
    LookupValues lv;
    lv.height = trace.height();
    lv.num_lookups = c.graph.lookups.size();
    lv.arg_offsets.push_back(0);
    for (auto& l : c.graph.lookups) lv.arg_offsets.push_back(lv.arg_offsets.back() + l.args.size());
    size_t aw = lv.arg_offsets.back();
    lv.multiplicities.assign(lv.height * lv.num_lookups, Fp::zero());
    lv.args.assign(lv.height * aw, Fp::zero());
    if (lv.height == 0 || lv.num_lookups == 0) return lv;
    size_t h = lv.height;
    long long hh = (long long)h;
#pragma omp parallel
    {
        std::vector<Fp> buf;
#pragma omp for schedule(static)
        for (long long rr = 0; rr < hh; rr++) {
            size_t r = (size_t)rr, rn = (r + 1) % h;
            VarValues<Fp> v{};
            v.main[0] = trace.row(r);
            v.main[1] = trace.row(rn);
            if (c.has_preprocessed) {
                v.preprocessed[0] = c.preprocessed.row(r);
                v.preprocessed[1] = c.preprocessed.row(rn);
            }
            v.is_first_row = r == 0 ? Fp::one() : Fp::zero();
            v.is_last_row = r == h - 1 ? Fp::one() : Fp::zero();
            v.is_transition = r == h - 1 ? Fp::zero() : Fp::one();
            sweep_range(c.graph, v, buf, c.graph.lookup_prefix_len);
            for (size_t s = 0; s < lv.num_lookups; s++) {
                auto& l = c.graph.lookups[s];
                lv.multiplicities[r * lv.num_lookups + s] = buf[l.multiplicity];
                for (size_t a = 0; a < l.args.size(); a++) lv.args[r * aw + lv.arg_offsets[s] + a] = buf[l.args[a]];
            }
        }
    }
    return lv;
}

// src/lookup.rs:375-384
inline Fp2 fingerprint(const Fp2& r, const Fp* coeffs, size_t n) {
    Fp2 acc = Fp2::zero();
    for (size_t i = n; i-- > 0;) acc = acc * r + coeffs[i];
    return acc;
}

// src/lookup.rs:472-555. Returns the stage-2 traces FLATTENED to base columns (height x max(L,1)*2) and
// the intermediate accumulators.
inline void stage_2_traces(const std::vector<const LookupValues*>& circuits, Fp2 beta, Fp2 gamma, Fp2 accumulator,
                           std::vector<Matrix>& traces, std::vector<Fp2>& intermediate) {
    traces.clear();
    intermediate.clear();
    for (const LookupValues* c : circuits) {
        size_t L = c->num_lookups, h = c->height;
        size_t width = std::max<size_t>(L, 1);
        Matrix t(h, 2 * width);
        if (L > 0) {
            size_t total = h * L;
            std::vector<Fp2> msgs(total), inv(total);
            long long tt = (long long)total;
#pragma omp parallel for schedule(static)
            for (long long i = 0; i < tt; i++) {
                size_t row = (size_t)i / L, lk = (size_t)i % L;
                msgs[i] = beta + fingerprint(gamma, c->args_at(row, lk), c->num_args(lk));
            }
            // chunked Montgomery inversion (p3 batch_multiplicative_inverse; any exact method is identical)
            const size_t CH = 4096;
            long long nch = (long long)((total + CH - 1) / CH);
#pragma omp parallel for schedule(static)
            for (long long ch = 0; ch < nch; ch++) {
                size_t lo = (size_t)ch * CH, n = std::min(CH, total - lo);
                batch_inverse(msgs.data() + lo, inv.data() + lo, n);
            }
            Fp2 local = Fp2::zero();
            for (size_t i = 0; i < total; i++) {
                t.values[2 * i] = local.c[0];
                t.values[2 * i + 1] = local.c[1];
                local += inv[i] * c->multiplicities[i];
            }
            accumulator += local;
        }
        intermediate.push_back(accumulator);
        traces.push_back(std::move(t));
    }
}

// ---- selectors (p3-commit) ------------------------------------------------------------------------
struct SelectorsOnCoset {
    std::vector<Fp> is_first_row, is_last_row, is_transition, inv_vanishing;
};
// trace domain H_n (shift 1), evaluated on the coset shift * H_{n << rate_bits}, natural order.
inline SelectorsOnCoset selectors_on_coset(unsigned log_n, unsigned rate_bits, Fp shift) {
    size_t big = size_t(1) << (log_n + rate_bits);
    SelectorsOnCoset s;
    Fp s_pow_n = shift.exp_power_of_2(log_n);
    size_t q = size_t(1) << rate_bits;
    std::vector<Fp> evals(q), evals_inv(q);
    Fp wq = two_adic_generator(rate_bits), acc = Fp::one();
    for (size_t i = 0; i < q; i++) { evals[i] = s_pow_n * acc - Fp::one(); acc *= wq; }
    batch_inverse(evals.data(), evals_inv.data(), q);
    std::vector<Fp> xs(big);
    Fp g = two_adic_generator(log_n + rate_bits);
    acc = shift;
    for (size_t i = 0; i < big; i++) { xs[i] = acc; acc *= g; }
    Fp subgroup_last = two_adic_generator(log_n).inverse();
    auto single = [&](Fp point) {
        std::vector<Fp> den(big), inv(big), out(big);
        for (size_t i = 0; i < big; i++) den[i] = xs[i] - point;
        batch_inverse(den.data(), inv.data(), big);
        for (size_t i = 0; i < big; i++) out[i] = evals[i % q] * inv[i];
        return out;
    };
    s.is_first_row = single(Fp::one());
    s.is_last_row = single(subgroup_last);
    s.is_transition.resize(big);
    s.inv_vanishing.resize(big);
    for (size_t i = 0; i < big; i++) { s.is_transition[i] = xs[i] - subgroup_last; s.inv_vanishing[i] = evals_inv[i % q]; }
    return s;
}
struct SelectorsAtPoint {
    Fp2 is_first_row, is_last_row, is_transition, inv_vanishing;
};
inline SelectorsAtPoint selectors_at_point(unsigned log_n, Fp2 point) {
    SelectorsAtPoint s;
    Fp2 z_h = point.exp_power_of_2(log_n) - Fp::one();
    Fp ginv = two_adic_generator(log_n).inverse();
    s.is_first_row = z_h * (point - Fp::one()).inverse();
    s.is_last_row = z_h * (point - ginv).inverse();
    s.is_transition = point - ginv;
    s.inv_vanishing = z_h.inverse();
    return s;
}

// View of a committed LDE on the quotient domain: `get_evaluations_on_domain` = the first nq stored rows,
// bit-reversed. row(i) = natural index i on the coset GENERATOR * H_{nq}.
struct DomainView {
    const Fp* data = nullptr;
    size_t width = 0;
    unsigned log_nq = 0;
    const Fp* row(size_t i) const { return data + reverse_bits_len(i, log_nq) * width; }
};

// src/prover.rs:756-962. Returns nq extension values (natural order on the quotient domain).
inline std::vector<Fp2> quotient_values(const Circuit& c, const Fp publics[8], unsigned log_n, unsigned log_q,
                                        const DomainView* pre, const DomainView& s1, const DomainView& s2, Fp2 alpha) {
    size_t nq = size_t(1) << (log_n + log_q), n = size_t(1) << log_n;
    SelectorsOnCoset sels = selectors_on_coset(log_n, log_q, Fp(GL_GENERATOR));
    Fp inj_norm = (Fp((u64)n) * two_adic_generator(log_n)).inverse();
    size_t next_step = size_t(1) << log_q;
    size_t k = c.constraint_count;
    std::vector<Fp2> alpha_powers(k);  // reversed: constraint i weighted by alpha^{k-1-i}
    Fp2 acc = Fp2::one();
    for (size_t i = 0; i < k; i++) { alpha_powers[k - 1 - i] = acc; acc *= alpha; }
    Fp delta_scaled[2] = {(publics[6] - publics[4]) * inj_norm, (publics[7] - publics[5]) * inj_norm};
    std::vector<Fp2> out(nq);
    long long nn = (long long)nq;
#pragma omp parallel
    {
        std::vector<Fp> buf, cv;
#pragma omp for schedule(static)
        for (long long ii = 0; ii < nn; ii++) {
            size_t i = (size_t)ii, inext = (i + next_step) & (nq - 1);
            VarValues<Fp> v{};
            if (pre) { v.preprocessed[0] = pre->row(i); v.preprocessed[1] = pre->row(inext); }
            v.main[0] = s1.row(i); v.main[1] = s1.row(inext);
            v.stage2[0] = s2.row(i); v.stage2[1] = s2.row(inext);
            v.publics = publics;
            v.is_first_row = sels.is_first_row[i];
            v.is_last_row = sels.is_last_row[i];
            v.is_transition = sels.is_transition[i];
            sweep_range(c.graph, v, buf, c.graph.nodes.size());
            cv.clear();
            for (u32 z : c.graph.zeros) cv.push_back(buf[z]);
            logup_constraint_values<Fp>(c.graph.lookups, buf, v.stage2[0], v.stage2[1], publics, delta_scaled, v.is_last_row,
                                        Fp(GL_EXT_W), cv);
            Fp a0 = Fp::zero(), a1 = Fp::zero();
            for (size_t j = 0; j < k; j++) { a0 += cv[j] * alpha_powers[j].c[0]; a1 += cv[j] * alpha_powers[j].c[1]; }
            out[i] = Fp2(a0 * sels.inv_vanishing[i], a1 * sels.inv_vanishing[i]);
        }
    }
    return out;
}

}  // namespace orc
