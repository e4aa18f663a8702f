// ORACLE C entry points, part 3 (test infrastructure, not product code): the whole prover on the CPU
// (src/prover.rs:289-603 over oracle/cpu_fri.hpp) and the restated verifier (oracle/cpu_verify.hpp).
// Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's CPU legs ONLY.
#include "orc_system.hpp"
#include <cstdlib>
#include <cstring>

using namespace orc;

static std::vector<std::vector<Fp>> read_claims(const u64* claims, const u64* offsets, u64 n) {
    std::vector<std::vector<Fp>> out(n);
    for (u64 i = 0; i < n; i++)
        for (u64 k = offsets[i]; k < offsets[i + 1]; k++) out[i].push_back(Fp(claims[k]));
    return out;
}

static thread_local std::string g_orc_err;
static thread_local TranscriptTrace g_orc_trace;  // of the last orc_prove on this thread

extern "C" {

const char* orc_last_error() { return g_orc_err.c_str(); }

// traces[i]: heights[i] x main_width of circuit i (heights[i] = 0: inactive). claims: flat values, offsets[n_claims + 1].
// stage_ms[6]: stage1_commit, claims, stage2_commit, quotient, fri_open, total. Returns 0 or -1 (orc_last_error()).
int orc_prove(void* s, const u64* const* traces, const u64* heights, const u64* claims, const u64* offsets, u64 n_claims,
              u8** proof_out, u64* proof_len, double* stage_ms) {
    try {
        OrcSystem& sys = *(OrcSystem*)s;
        std::vector<Matrix> mats;
        for (size_t i = 0; i < sys.shape.circuits.size(); i++) {
            size_t w = sys.shape.circuits[i].main_width;
            Matrix m(heights[i], w);
            for (size_t k = 0; k < heights[i] * w; k++) m.values[k] = Fp(traces[i][k]);
            mats.push_back(std::move(m));
        }
        std::vector<const Matrix*> ptrs;
        for (auto& m : mats) ptrs.push_back(&m);
        ProveTimings tm;
        Proof proof = sys.get_prover().prove(read_claims(claims, offsets, n_claims), ptrs, &tm);
        g_orc_trace = tm.trace;
        std::vector<u8> bytes = proof_to_bytes(proof);
        *proof_out = (u8*)malloc(bytes.size());
        memcpy(*proof_out, bytes.data(), bytes.size());
        *proof_len = bytes.size();
        if (stage_ms) {
            const char* names[5] = {"stark/stage1_commit", "stark/claims", "stark/stage2_commit", "stark/quotient", "stark/fri_open"};
            double total = 0;
            for (int i = 0; i < 5; i++) { stage_ms[i] = tm.ms[names[i]]; total += stage_ms[i]; }
            stage_ms[5] = total;
        }
        return 0;
    } catch (const std::exception& e) {
        g_orc_err = e.what();
        return -1;
    }
}
void orc_bytes_free(u8* p) { free(p); }
// the challenges / query indices of the last orc_prove on this thread (same layout as msh_last_transcript)
u64 orc_last_transcript(u64* challenges2, u64 cap, u64* indices, u64 cap_idx, u64* n_indices) {
    for (size_t i = 0; i < g_orc_trace.challenges.size() && i < cap; i++) {
        challenges2[2 * i] = g_orc_trace.challenges[i].c[0].v;
        challenges2[2 * i + 1] = g_orc_trace.challenges[i].c[1].v;
    }
    for (size_t i = 0; i < g_orc_trace.query_indices.size() && i < cap_idx; i++) indices[i] = g_orc_trace.query_indices[i];
    if (n_indices) *n_indices = g_orc_trace.query_indices.size();
    return g_orc_trace.challenges.size();
}

// the preprocessed commitment (verifier key); returns 0 if the system has no preprocessed trace
int orc_preprocessed_commit(void* s, u8* out32) {
    OrcSystem& sys = *(OrcSystem*)s;
    const ProverKey& k = sys.get_prover().key();
    if (!k.has_preprocessed) return 0;
    memcpy(out32, k.preprocessed_commit.data(), 32);
    return 1;
}

// Returns the VerifyError code (0 = accepted), or -1 when the bytes do not deserialize (Proof::from_bytes Err).
int orc_verify(void* s, const u64* claims, const u64* offsets, u64 n_claims, const u8* proof_bytes, u64 len) {
    try {
        OrcSystem& sys = *(OrcSystem*)s;
        Proof proof;
        if (!proof_from_bytes(proof_bytes, len, proof)) return -1;
        const ProverKey& k = sys.get_prover().key();
        return (int)verify_multiple_claims(sys.shape, k.has_preprocessed, k.preprocessed_commit, read_claims(claims, offsets, n_claims), proof);
    } catch (const std::exception& e) {
        g_orc_err = e.what();
        return -2;
    }
}

// The two transcript scripts of the reference's `gen_challenger_refs` (src/types.rs:284-319): a challenger over an EMPTY seed.
// out8 = SAMPLE_BITS, APCS (c0, c1), AFRI (c0, c1), BETA (c0, c1), SAMPLE_BITS2
void orc_gen_challenger_refs(u64* out8) {
    Challenger a{std::vector<u8>()};
    a.observe(Fp(0x0102030405060708ull));
    out8[0] = a.sample_bits(20);
    Challenger ch{std::vector<u8>()};
    ch.observe(Fp(0x0102030405060708ull));
    ch.observe(Fp(0x1122334455667788ull));
    Fp2 apcs = ch.sample_ext(), afri = ch.sample_ext();
    out8[1] = apcs.c[0].v; out8[2] = apcs.c[1].v;
    out8[3] = afri.c[0].v; out8[4] = afri.c[1].v;
    ch.observe(Fp(0x00000000deadbeefull));
    Fp2 beta = ch.sample_ext();
    out8[5] = beta.c[0].v; out8[6] = beta.c[1].v;
    ch.observe(Fp(0x0a0b0c0d01020304ull));
    ch.observe(Fp(0x0000000000000002ull));
    out8[7] = ch.sample_bits(20);
}

}  // extern "C"

// ---- examples/pcs_example.rs:28-122 on the CPU: commit -> observe -> sample zeta -> open at [zeta; num_open] -> bytes,
// and the matching verification (transcript replay + TwoAdicFriPcs::verify). One round, any number of matrices.
extern "C" {

static void pcs_params(u32 log_blowup, u32 log_final_poly_len, u32 num_queries, u32 commit_pow, u32 query_pow,
                       CommitmentParameters& cp, FriParameters& fp) {
    cp.log_blowup = log_blowup;
    fp.log_final_poly_len = log_final_poly_len;
    fp.max_log_arity = 1;
    fp.num_queries = num_queries;
    fp.commit_proof_of_work_bits = commit_pow;
    fp.query_proof_of_work_bits = query_pow;
}

int orc_pcs_example_prove(const u64* const* mats, const u64* heights, const u64* widths, u64 n_mats, u32 log_blowup,
                          u32 log_final_poly_len, u32 num_queries, u32 commit_pow, u32 query_pow, u32 num_open, u8* root32,
                          u64* zeta2, u8** out, u64* out_len) {
    try {
        CommitmentParameters cp;
        FriParameters fp;
        pcs_params(log_blowup, log_final_poly_len, num_queries, commit_pow, query_pow, cp, fp);
        std::vector<Matrix> ms;
        for (u64 i = 0; i < n_mats; i++) {
            Matrix m(heights[i], widths[i]);
            for (size_t k = 0; k < m.values.size(); k++) m.values[k] = Fp(mats[i][k]);
            ms.push_back(std::move(m));
        }
        std::vector<const Matrix*> ptrs;
        for (auto& m : ms) ptrs.push_back(&m);
        Digest root;
        auto pd = cpu_commit(ptrs, log_blowup, root);
        memcpy(root32, root.data(), 32);
        Challenger ch = Challenger::for_config(cp, fp);
        ch.observe(root);
        Fp2 zeta = ch.sample_ext();
        zeta2[0] = zeta.c[0].v;
        zeta2[1] = zeta.c[1].v;
        std::vector<OpenRound> rounds(1);
        rounds[0].data = pd.get();
        for (u64 i = 0; i < n_mats; i++) rounds[0].points.push_back(std::vector<Fp2>(num_open, zeta));
        CpuOpenDevice dev(rounds, log_blowup);
        std::vector<OpenedValuesForRound> opened;
        FriProof proof;
        pcs_open(dev, rounds, cp, fp, ch, opened, proof);
        std::vector<u8> bytes = pcs_open_to_bytes(opened, proof);
        *out = (u8*)malloc(bytes.size());
        memcpy(*out, bytes.data(), bytes.size());
        *out_len = bytes.size();
        return 0;
    } catch (const std::exception& e) {
        g_orc_err = e.what();
        return -1;
    }
}

// 1 = accepted, 0 = rejected, -1 = the bytes do not parse
int orc_pcs_example_verify(const u8* root32, const u64* heights, const u64* widths, u64 n_mats, u32 log_blowup,
                           u32 log_final_poly_len, u32 num_queries, u32 commit_pow, u32 query_pow, u32 num_open, const u8* bytes,
                           u64 len) {
    try {
        CommitmentParameters cp;
        FriParameters fp;
        pcs_params(log_blowup, log_final_poly_len, num_queries, commit_pow, query_pow, cp, fp);
        std::vector<OpenedValuesForRound> opened;
        FriProof proof;
        if (!pcs_open_from_bytes(bytes, len, opened, proof)) return -1;
        if (opened.size() != 1 || opened[0].size() != n_mats) return 0;
        Digest root;
        memcpy(root.data(), root32, 32);
        Challenger ch = Challenger::for_config(cp, fp);
        ch.observe(root);
        Fp2 zeta = ch.sample_ext();
        VerifierRound vr;
        vr.commit = root;
        for (u64 i = 0; i < n_mats; i++) {
            if (opened[0][i].size() != num_open) return 0;
            VerifierMat vm;
            vm.log_degree = log2_strict(heights[i]);
            for (u32 k = 0; k < num_open; k++) {
                if (opened[0][i][k].size() != widths[i]) return 0;
                vm.points.push_back({zeta, opened[0][i][k]});
            }
            vr.mats.push_back(std::move(vm));
        }
        return pcs_verify({vr}, proof, cp, fp, ch) ? 1 : 0;
    } catch (const std::exception& e) {
        g_orc_err = e.what();
        return -2;
    }
}

}  // extern "C"
