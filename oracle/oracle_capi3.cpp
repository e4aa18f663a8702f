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

}  // extern "C"
