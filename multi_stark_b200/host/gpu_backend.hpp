// ProverBackend over the libmsgpu C ABI (include/msgpu.h): what a Rust `GpuFriPcs` + the prover.rs hook
// would call. All matrices stay resident in HBM between the LDE, Merkle, quotient and opening
// stages; the host sees roots, opened values, the final polynomial and the query openings only.
#pragma once
#include "prover.hpp"
#include <cstdlib>
#include <cstring>
#include "program.hpp"
#include "../../include/msgpu.h"

namespace msh {

struct GpuError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void gpu_check(int code) {
    if (code != 0) throw GpuError(std::string("libmsgpu: ") + msgpu_last_error());
}

struct GpuPcsHandle : PcsHandle {
    msgpu_pdata* pd = nullptr;
    std::vector<std::pair<size_t, size_t>> shapes;
    bool owns = true;
    explicit GpuPcsHandle(msgpu_pdata* p, bool owns_ = true) : pd(p), owns(owns_) {
        for (uint64_t i = 0; i < msgpu_pdata_num_matrices(pd); i++) {
            uint64_t r = 0, c = 0;
            gpu_check(msgpu_pdata_matrix(pd, i, nullptr, &r, &c));
            shapes.push_back({(size_t)r, (size_t)c});
        }
    }
    ~GpuPcsHandle() override {
        if (owns) msgpu_pdata_free(pd);
    }
    size_t num_matrices() const override { return shapes.size(); }
    size_t matrix_height(size_t i) const override { return shapes[i].first; }
    size_t matrix_width(size_t i) const override { return shapes[i].second; }
};

inline std::vector<BatchOpening> gpu_open_batch(msgpu_ctx* ctx, const msgpu_pdata* pd, const std::vector<size_t>& indices) {
    size_t nm = msgpu_pdata_num_matrices(pd), tw = 0, maxh = 0;
    std::vector<size_t> widths;
    for (size_t i = 0; i < nm; i++) {
        uint64_t r = 0, c = 0;
        gpu_check(msgpu_pdata_matrix(pd, i, nullptr, &r, &c));
        widths.push_back((size_t)c);
        tw += (size_t)c;
        maxh = std::max(maxh, (size_t)r);
    }
    size_t depth = log2_strict(maxh), n = indices.size();
    std::vector<uint64_t> idx(indices.begin(), indices.end()), opened(std::max<size_t>(n * tw, 1));
    std::vector<uint8_t> proofs(std::max<size_t>(n * depth * 32, 1));
    gpu_check(msgpu_open_batch(ctx, pd, idx.data(), n, opened.data(), proofs.data()));
    std::vector<BatchOpening> out(n);
    for (size_t q = 0; q < n; q++) {
        size_t o = q * tw;
        for (size_t m = 0; m < nm; m++) {
            std::vector<Fp> row(widths[m]);
            for (size_t c = 0; c < widths[m]; c++) row[c].v = opened[o++];
            out[q].opened_values.push_back(std::move(row));
        }
        out[q].opening_proof.resize(depth);
        for (size_t l = 0; l < depth; l++) memcpy(out[q].opening_proof[l].data(), proofs.data() + (q * depth + l) * 32, 32);
    }
    return out;
}

class GpuOpenDevice : public OpenDevice {
  public:
    GpuOpenDevice(msgpu_ctx* ctx, const std::vector<OpenRound>& rounds, uint32_t log_blowup) : ctx_(ctx) {
        std::vector<uint64_t> npts, pts;
        for (auto& r : rounds) {
            auto* h = dynamic_cast<GpuPcsHandle*>(r.data);
            if (!h) throw GpuError("open: prover data does not belong to the GPU backend");
            pds_.push_back(h->pd);
            shapes_.push_back(h->shapes);
            if (r.points.size() != h->shapes.size()) throw GpuError("open: one point list per committed matrix expected");
            for (auto& mp : r.points) {
                npts.push_back(mp.size());
                for (auto& z : mp) { pts.push_back(z.c[0].v); pts.push_back(z.c[1].v); }
            }
            points_.push_back(r.points);
        }
        if (pts.empty()) pts.push_back(0);
        gpu_check(msgpu_open_begin(ctx_, pds_.size(), pds_.data(), npts.data(), pts.data(), log_blowup, &op_, &n_values_));
    }
    ~GpuOpenDevice() override { msgpu_open_free(op_); }

    std::vector<OpenedValuesForRound> evaluate() override {
        std::vector<uint64_t> flat(std::max<size_t>(2 * n_values_, 1));
        gpu_check(msgpu_open_values(op_, flat.data()));
        std::vector<OpenedValuesForRound> out(pds_.size());
        size_t o = 0;
        for (size_t r = 0; r < pds_.size(); r++) {
            out[r].resize(shapes_[r].size());
            for (size_t m = 0; m < shapes_[r].size(); m++) {
                out[r][m].resize(points_[r][m].size());
                for (auto& pv : out[r][m]) {
                    pv.resize(shapes_[r][m].second);
                    for (auto& v : pv) { v.c[0].v = flat[o++]; v.c[1].v = flat[o++]; }
                }
            }
        }
        return out;
    }
    void reduce(Fp2 alpha, unsigned& log_max_height) override {
        uint64_t a[2] = {alpha.c[0].v, alpha.c[1].v}, n = 0;
        uint32_t lm = 0;
        gpu_check(msgpu_open_reduce(op_, a, &n, &lm));
        log_max_height = lm;
    }
    size_t current_len() override {
        uint64_t len = 0;
        gpu_check(msgpu_fri_current_len(op_, &len));
        return (size_t)len;
    }
    Digest commit_round() override {
        Digest d;
        gpu_check(msgpu_fri_commit_round(op_, d.data()));
        return d;
    }
    void fold(Fp2 beta) override {
        uint64_t b[2] = {beta.c[0].v, beta.c[1].v};
        gpu_check(msgpu_fri_fold(op_, b));
    }
    // p3-fri `commit_phase` with the transcript on the device (no root read-back per round) when there is no proof of work
    // to grind; the host challenger replays the rounds afterwards and every beta is compared.
    void commit_phase(Challenger& ch, size_t stop_len, size_t pow_bits, FriProof& proof) override {
        const std::vector<u8>& buf = ch.input_buffer();
        static const bool host_loop = getenv("MSGPU_FRI_HOST_LOOP") != nullptr;
        if (pow_bits != 0 || buf.empty() || buf.size() > 960 || host_loop || current_len() <= stop_len) {
            OpenDevice::commit_phase(ch, stop_len, pow_bits, proof);
            return;
        }
        uint8_t roots[64 * 32];
        uint64_t bt[64 * 2], n = 0;
        gpu_check(msgpu_fri_commit_phase(op_, buf.data(), buf.size(), stop_len, 64, roots, bt, &n));
        for (uint64_t k = 0; k < n; k++) {
            Digest commit;
            memcpy(commit.data(), roots + 32 * k, 32);
            ch.observe(commit);
            proof.commit_phase_commits.push_back(commit);
            proof.commit_pow_witnesses.push_back(ch.grind(0));
            Fp2 beta = ch.sample_ext();
            if (beta.c[0].v != bt[2 * k] || beta.c[1].v != bt[2 * k + 1])
                throw GpuError("fri: the device transcript and the host challenger disagree on a folding challenge");
            betas.push_back(beta);
        }
    }
    std::vector<Fp2> read_current() override {
        size_t len = current_len();
        std::vector<uint64_t> flat(2 * len);
        gpu_check(msgpu_fri_read_current(op_, flat.data()));
        std::vector<Fp2> out(len);
        for (size_t i = 0; i < len; i++) { out[i].c[0].v = flat[2 * i]; out[i].c[1].v = flat[2 * i + 1]; }
        return out;
    }
    std::vector<BatchOpening> open_round(size_t r, const std::vector<size_t>& indices) override {
        return gpu_open_batch(ctx_, pds_[r], indices);
    }
    std::vector<BatchOpening> open_layer(size_t k, const std::vector<size_t>& pair_indices) override {
        const msgpu_pdata* pd = msgpu_fri_layer_pdata(op_, k);
        if (!pd) throw GpuError("open: no such commit-phase layer");
        return gpu_open_batch(ctx_, pd, pair_indices);
    }

    void open_queries(const std::vector<size_t>& indices, const std::vector<unsigned>& round_shifts, size_t n_layers,
                      std::vector<std::vector<BatchOpening>>& rounds_out, std::vector<std::vector<BatchOpening>>& layers_out) override {
        std::vector<const msgpu_pdata*> trees(pds_.begin(), pds_.end());
        std::vector<uint32_t> shifts(round_shifts.begin(), round_shifts.end());
        for (size_t k = 0; k < n_layers; k++) {
            const msgpu_pdata* pd = msgpu_fri_layer_pdata(op_, k);
            if (!pd) throw GpuError("open: no such commit-phase layer");
            trees.push_back(pd);
            shifts.push_back((uint32_t)(k + 1));
        }
        size_t n = indices.size(), open_total = 0, proof_total = 0;
        std::vector<std::vector<size_t>> widths(trees.size());
        std::vector<size_t> depth(trees.size()), tw(trees.size());
        for (size_t t = 0; t < trees.size(); t++) {
            size_t nm = msgpu_pdata_num_matrices(trees[t]), maxh = 0;
            for (size_t i = 0; i < nm; i++) {
                uint64_t r = 0, c = 0;
                gpu_check(msgpu_pdata_matrix(trees[t], i, nullptr, &r, &c));
                widths[t].push_back((size_t)c);
                tw[t] += (size_t)c;
                maxh = std::max(maxh, (size_t)r);
            }
            depth[t] = log2_strict(maxh);
            open_total += n * tw[t];
            proof_total += n * depth[t] * 32;
        }
        std::vector<uint64_t> idx(indices.begin(), indices.end()), opened(std::max<size_t>(open_total, 1));
        std::vector<uint8_t> proofs(std::max<size_t>(proof_total, 1));
        gpu_check(msgpu_open_batch_multi(ctx_, trees.data(), shifts.data(), trees.size(), idx.data(), n, opened.data(), proofs.data()));
        rounds_out.assign(round_shifts.size(), {});
        layers_out.assign(n_layers, {});
        size_t oo = 0, po = 0;
        for (size_t t = 0; t < trees.size(); t++) {
            std::vector<BatchOpening>& out = t < round_shifts.size() ? rounds_out[t] : layers_out[t - round_shifts.size()];
            out.resize(n);
            for (size_t q = 0; q < n; q++) {
                for (size_t m = 0; m < widths[t].size(); m++) {
                    std::vector<Fp> row(widths[t][m]);
                    for (auto& v : row) v.v = opened[oo++];
                    out[q].opened_values.push_back(std::move(row));
                }
                out[q].opening_proof.resize(depth[t]);
                for (size_t l = 0; l < depth[t]; l++, po += 32) memcpy(out[q].opening_proof[l].data(), proofs.data() + po, 32);
            }
        }
    }

  private:
    msgpu_ctx* ctx_;
    msgpu_open* op_ = nullptr;
    uint64_t n_values_ = 0;
    std::vector<const msgpu_pdata*> pds_;
    std::vector<std::vector<std::pair<size_t, size_t>>> shapes_;
    std::vector<std::vector<std::vector<Fp2>>> points_;
};

class GpuBackend : public ProverBackend {
  public:
    GpuBackend(msgpu_ctx* ctx, const SystemShape& shape) : ctx_(ctx), shape_(shape) {
        for (auto& c : shape_.circuits) {
            GraphDesc d;
            d.build(c.graph, c.preprocessed_width, c.main_width, c.stage_2_width);
            msgpu_program* p = nullptr;
            gpu_check(msgpu_program_create(ctx_, &d.desc, &p));
            programs_.push_back(p);
            // the preprocessed trace in natural order stays on the device for the stage-2 construction
            uint64_t* dp = nullptr;
            if (c.has_preprocessed && !c.preprocessed.values.empty())
                dp = upload((const uint64_t*)c.preprocessed.values.data(), c.preprocessed.values.size());
            pre_dev_.push_back(dp);
        }
    }
    ~GpuBackend() override {
        end_proof();
        if (injected_) msgpu_pdata_free(injected_);
        for (auto* p : programs_) msgpu_program_free(p);
        for (auto* d : pre_dev_)
            if (d) msgpu_free(ctx_, d);
    }

    PcsHandlePtr commit(const std::vector<const Matrix*>& evals, Digest& root) override {
        std::vector<const uint64_t*> ptrs;
        std::vector<uint64_t> hs, ws;
        for (auto* m : evals) { ptrs.push_back((const uint64_t*)m->values.data()); hs.push_back(m->height()); ws.push_back(m->width); }
        msgpu_pdata* pd = nullptr;
        gpu_check(msgpu_commit(ctx_, ptrs.data(), hs.data(), ws.data(), evals.size(), (uint32_t)shape_.log_blowup(), &pd, root.data()));
        return std::make_shared<GpuPcsHandle>(pd);
    }

    // The next proof's stage-1 commitment was made elsewhere (one wide matrix committed by column blocks over several GPUs and
    // assembled with msgpu_pdata_from_parts): adopt it instead of uploading and committing the traces. Only for systems whose
    // stage-2 construction does not read the main traces (no lookups): they are not on this device.
    void inject_stage1(msgpu_pdata* pd) {
        if (injected_) msgpu_pdata_free(injected_);
        injected_ = pd;
    }

    PcsHandlePtr commit_stage1(const std::vector<size_t>& circuits, const std::vector<MatrixView>& traces, Digest& root) override {
        end_proof();
        active_ = circuits;
        if (injected_) {
            msgpu_pdata* pd = injected_;
            injected_ = nullptr;
            auto h = std::make_shared<GpuPcsHandle>(pd);  // owns it from here
            if (h->num_matrices() != traces.size()) throw GpuError("injected stage-1 commitment: matrix count does not match the active circuits");
            for (size_t i = 0; i < traces.size(); i++) {
                if (h->matrix_height(i) != (traces[i].height() << shape_.log_blowup()) || h->matrix_width(i) != traces[i].width)
                    throw GpuError("injected stage-1 commitment: matrix shape does not match the trace");
                if (shape_.circuits[circuits[i]].num_lookups != 0)
                    throw GpuError("injected stage-1 commitment: a circuit with lookups needs its main trace on the device");
                trace_rows_.push_back(traces[i].height());
            }
            trace_dev_.assign(traces.size(), nullptr);
            announced_ = ClaimsView();
            gpu_check(msgpu_pdata_root(pd, root.data()));
            return h;
        }
        std::vector<const uint64_t*> ptrs;
        std::vector<uint64_t> hs, ws;
        for (auto& m : traces) {
            ptrs.push_back((const uint64_t*)m.data);
            hs.push_back(m.height());
            ws.push_back(m.width);
            trace_rows_.push_back(m.height());
        }
        // uploads on the copy stream (trace i + 1 crosses PCIe while trace i is extended), the claims queued right behind them
        msgpu_upload* up = nullptr;
        gpu_check(msgpu_upload_begin(ctx_, ptrs.data(), hs.data(), ws.data(), traces.size(), &up));
        try {
            prefetch_announced_claims();
        } catch (...) {
            msgpu_upload_free(up);
            throw;
        }
        trace_dev_.assign(traces.size(), nullptr);
        msgpu_pdata* pd = nullptr;
        int rc = msgpu_commit_upload(up, (uint32_t)shape_.log_blowup(), 1, trace_dev_.data(), 0, &pd, root.data());
        if (rc != 0) {
            trace_dev_.clear();  // the library released the buffers
            gpu_check(rc);
        }
        return std::make_shared<GpuPcsHandle>(pd);
    }

    void announce_claims(const ClaimsView& claims) override { announced_ = claims; }

    bool observe_claims(Challenger& ch, const ClaimsView& claims) override {
        Digest d;
        if (!claims_transcript_digest(ch.input_buffer(), claims, d)) return false;
        ch.set_flushed(d);
        return true;
    }

    Fp2 claims_accumulator(const ClaimsView& claims, Fp2 beta, Fp2 gamma) override {
        Fp2 acc = Fp2::zero();
        uint64_t b[2] = {beta.c[0].v, beta.c[1].v}, g[2] = {gamma.c[0].v, gamma.c[1].v};
        if (claims_dev_) {  // already resident (observe_claims)
            uint64_t out[2];
            gpu_check(msgpu_claims_accumulate(claims_dev_, b, g, out));
            drop_claims();
            return Fp2(Fp(out[0]), Fp(out[1]));
        }
        // the C ABI takes claims of one length per call; group consecutive claims of equal length
        size_t i = 0;
        while (i < claims.size()) {
            size_t len = claims.len(i), j = i;
            while (j < claims.size() && claims.len(j) == len) j++;
            uint64_t out[2];
            gpu_check(msgpu_claims_accumulator(ctx_, (const uint64_t*)claims.at(i), j - i, len, b, g, out));
            acc += Fp2(Fp(out[0]), Fp(out[1]));
            i = j;
        }
        return acc;
    }

    Challenger::BigHash big_hash() override {
        msgpu_ctx* ctx = ctx_;
        return [ctx](const u8* data, size_t n) {
            Digest d;
            gpu_check(msgpu_blake3_hash(ctx, data, n, d.data()));
            return d;
        };
    }

    PcsHandlePtr commit_stage2(Fp2 beta, Fp2 gamma, Fp2 acc, std::vector<Fp2>& intermediate, Digest& root) override {
        uint64_t b[2] = {beta.c[0].v, beta.c[1].v}, g[2] = {gamma.c[0].v, gamma.c[1].v};
        std::vector<uint64_t*> s2;
        std::vector<uint64_t> hs, ws;
        intermediate.clear();
        struct FreeAll {  // the stage-2 traces are scratch: released on every path, including a throwing msgpu_stage2_trace
            msgpu_ctx* ctx;
            std::vector<uint64_t*>& v;
            ~FreeAll() { for (auto* d : v) msgpu_free(ctx, d); }
        } guard{ctx_, s2};
        for (size_t p = 0; p < active_.size(); p++) {
            const Circuit& c = shape_.circuits[active_[p]];
            uint64_t rows = trace_rows_[p];
            void* out = nullptr;
            gpu_check(msgpu_malloc(ctx_, std::max<uint64_t>(rows * c.stage_2_width * 8, 8), &out));
            s2.push_back((uint64_t*)out);
            uint64_t local[2] = {0, 0};
            gpu_check(msgpu_stage2_trace(ctx_, programs_[active_[p]], pre_dev_[active_[p]], trace_dev_[p], rows, b, g, (uint64_t*)out, local));
            acc += Fp2(Fp(local[0]), Fp(local[1]));
            intermediate.push_back(acc);
            hs.push_back(rows);
            ws.push_back(c.stage_2_width);
        }
        msgpu_pdata* pd = nullptr;
        gpu_check(msgpu_commit_dev(ctx_, (const uint64_t* const*)s2.data(), hs.data(), ws.data(), s2.size(), (uint32_t)shape_.log_blowup(), &pd,
                                   root.data()));
        // the natural-order traces are not needed any more
        for (auto* d : trace_dev_) msgpu_free(ctx_, d);
        trace_dev_.clear();
        return std::make_shared<GpuPcsHandle>(pd);
    }

    PcsHandlePtr commit_quotient(const std::vector<QuotientJob>& jobs, PcsHandle* pre, PcsHandle* s1, PcsHandle* s2, Fp2 alpha,
                                 Digest& root) override {
        auto* h1 = dynamic_cast<GpuPcsHandle*>(s1);
        auto* h2 = dynamic_cast<GpuPcsHandle*>(s2);
        auto* hp = pre ? dynamic_cast<GpuPcsHandle*>(pre) : nullptr;
        uint64_t a[2] = {alpha.c[0].v, alpha.c[1].v};
        std::vector<uint64_t*> ldes;
        std::vector<uint64_t> hs, ws;
        uint32_t lb = (uint32_t)shape_.log_blowup();
        try {
            for (auto& j : jobs) {
                uint64_t pub[8];
                for (int k = 0; k < 8; k++) pub[k] = j.publics[k].v;
                uint64_t* lde = nullptr;
                gpu_check(msgpu_quotient(ctx_, programs_[j.circuit], j.preprocessed_idx >= 0 && hp ? hp->pd : nullptr,
                                         j.preprocessed_idx >= 0 ? (uint64_t)j.preprocessed_idx : 0, h1->pd, j.pos, h2->pd, j.pos,
                                         j.log_degree, j.log_quotient_degree, lb, pub, a, &lde, nullptr));
                ldes.push_back(lde);
                hs.push_back((uint64_t)1 << (j.log_degree + lb));
                ws.push_back((uint64_t)2 << j.log_quotient_degree);
            }
            msgpu_pdata* pd = nullptr;
            gpu_check(msgpu_commit_ldes_dev(ctx_, ldes.data(), hs.data(), ws.data(), ldes.size(), 1, &pd, root.data()));
            return std::make_shared<GpuPcsHandle>(pd);
        } catch (...) {
            for (auto* d : ldes) msgpu_free(ctx_, d);
            throw;
        }
    }

    std::unique_ptr<OpenDevice> open_begin(const std::vector<OpenRound>& rounds) override {
        return std::make_unique<GpuOpenDevice>(ctx_, rounds, (uint32_t)shape_.log_blowup());
    }

    void end_proof() override {
        drop_claims();
        for (auto* d : trace_dev_) msgpu_free(ctx_, d);
        trace_dev_.clear();
        trace_rows_.clear();
        active_.clear();
    }

  protected:  // the sharded backend (dist_backend.hpp) reuses the per-circuit state
    // true if the claim set is large and uniform (same rule on every rank: it depends on the shape only)
    static bool claims_on_device(const ClaimsView& claims) {
        size_t len = claims.uniform_len();
        return len != 0 && claims.size() * len >= 4096;
    }
    // BLAKE3(prefix || length-prefixed claims) on the device (false = small or ragged set: the host loop is cheaper; the
    // ABI's precondition is then checked here)
    bool claims_transcript_digest(const std::vector<u8>& prefix, const ClaimsView& claims, Digest& d) {
        if (!claims_on_device(claims)) {
            size_t total = claims.total();
            for (size_t k = 0; k < total; k++)
                if (claims.values[k].v >= GL_P) throw GpuError("claim value is not canonical");
            return false;
        }
        size_t len = claims.uniform_len();
        if (claims_dev_ && prefetched_from_ == (const void*)claims.at(0) && prefetched_n_ == claims.size() && prefetched_len_ == len) {
            gpu_check(msgpu_claims_digest(claims_dev_, prefix.data(), prefix.size(), d.data()));
        } else {
            drop_claims();
            gpu_check(msgpu_claims_upload(ctx_, (const uint64_t*)claims.at(0), claims.size(), len, prefix.data(), prefix.size(), &claims_dev_,
                                          d.data()));
        }
        prefetched_from_ = nullptr;
        return true;
    }
    // H2D copy + check that every value is canonical (the ABI's precondition, verified on the device)
    uint64_t* upload(const uint64_t* src, size_t n) {
        void* d = nullptr;
        gpu_check(msgpu_malloc(ctx_, std::max<size_t>(n * 8, 8), &d));
        if (n) {
            int rc = msgpu_upload_canonical(ctx_, d, src, n);
            if (rc != 0) {
                msgpu_free(ctx_, d);
                gpu_check(rc);
            }
        }
        return (uint64_t*)d;
    }
    void drop_claims() {
        if (claims_dev_) msgpu_claims_free(claims_dev_);
        claims_dev_ = nullptr;
        prefetched_from_ = nullptr;
    }
    // same size rule as observe_claims: small or ragged claim sets stay on the host
    void prefetch_announced_claims() {
        ClaimsView cl = announced_;
        announced_ = ClaimsView();
        size_t len = cl.uniform_len();
        if (len == 0 || cl.size() * len < 4096) return;
        drop_claims();
        gpu_check(msgpu_claims_prefetch(ctx_, (const uint64_t*)cl.at(0), cl.size(), len, &claims_dev_));
        prefetched_from_ = (const void*)cl.at(0);
        prefetched_n_ = cl.size();
        prefetched_len_ = len;
    }
    msgpu_pdata* injected_ = nullptr;
    ClaimsView announced_;
    const void* prefetched_from_ = nullptr;
    size_t prefetched_n_ = 0, prefetched_len_ = 0;
    msgpu_ctx* ctx_;
    msgpu_claims* claims_dev_ = nullptr;
    const SystemShape& shape_;
    std::vector<msgpu_program*> programs_;
    std::vector<uint64_t*> pre_dev_;
    std::vector<size_t> active_;
    std::vector<uint64_t*> trace_dev_;
    std::vector<uint64_t> trace_rows_;
};

}  // namespace msh
