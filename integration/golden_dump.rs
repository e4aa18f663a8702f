//! Golden-vector dump for the REAL reference (argumentcomputer/multi-stark + Plonky3 at the pinned rev).
//!
//! Not compiled in this repository (no cargo in the build image). A maintainer drops this file into the reference as
//! `tests/golden_dump.rs`, runs `cargo test --release --features parallel golden_dump -- --nocapture`, and diffs the printed
//! JSON lines against `tests/golden/proofs.json` of the B200 repository: the per-case `stage_1_commit`, `stage_2_commit`,
//! `quotient_commit`, `proof_bytes` and `proof_sha256` must be identical. Together with `cargo test gen_pcs_refs
//! gen_challenger_refs -- --nocapture` (labels of `tests/golden/pcs_refs.json`) this pins the MMCS tree shape, the transcript
//! byte order, the FRI rules and the proof wire format that the CPU oracle restates from published Plonky3 semantics.
//!
//! The workload is the benchmark's (`benches/multi_stark.rs:73-238`): U32-add circuit + preprocessed byte table, xorshift32
//! streams from 0xdeadbeef / 0xcafebabe, claims `[1, x, y, z]`. `u32_add_system` / `u32_add_witness` below stand for the
//! bench's own `build_system` / `build_witness` helpers, which live in the bench binary and have to be made reachable from
//! the test (e.g. moved into `src/test_circuits/`).
#![cfg(test)]
use multi_stark::types::{CommitmentParameters, FriParameters};
use sha2::{Digest, Sha256};

struct Case {
    log_adds: u32,
    log_blowup: usize,
    num_queries: usize,
    log_final_poly_len: usize,
    commit_pow_bits: usize,
    query_pow_bits: usize,
}

const CASES: &[Case] = &[
    Case { log_adds: 4, log_blowup: 1, num_queries: 10, log_final_poly_len: 0, commit_pow_bits: 0, query_pow_bits: 0 },
    Case { log_adds: 8, log_blowup: 1, num_queries: 100, log_final_poly_len: 0, commit_pow_bits: 0, query_pow_bits: 0 },
    Case { log_adds: 12, log_blowup: 1, num_queries: 100, log_final_poly_len: 0, commit_pow_bits: 0, query_pow_bits: 0 },
    Case { log_adds: 10, log_blowup: 2, num_queries: 30, log_final_poly_len: 0, commit_pow_bits: 10, query_pow_bits: 10 },
    Case { log_adds: 9, log_blowup: 3, num_queries: 20, log_final_poly_len: 2, commit_pow_bits: 0, query_pow_bits: 0 },
];

#[test]
fn golden_dump() {
    for c in CASES {
        let commitment_parameters = CommitmentParameters { log_blowup: c.log_blowup, cap_height: 0 };
        let fri_parameters = FriParameters {
            log_final_poly_len: c.log_final_poly_len,
            max_log_arity: 1,
            num_queries: c.num_queries,
            commit_proof_of_work_bits: c.commit_pow_bits,
            query_proof_of_work_bits: c.query_pow_bits,
        };
        // benches/multi_stark.rs:260-267 (system = [ByteTable, U32Add]) and :171-238 (witness + claims)
        let (system, key) = u32_add_system(commitment_parameters, fri_parameters);
        let (witness, claims) = u32_add_witness(&system, 1usize << c.log_adds);
        let claim_refs: Vec<&[_]> = claims.iter().map(|c| c.as_slice()).collect();
        let proof = system.prove_multiple_claims(&key, &claim_refs, witness);
        system.verify_multiple_claims(&claim_refs, &proof).expect("the reference verifier accepts its own proof");
        let bytes = proof.to_bytes().expect("serialisable");
        let n_act = u64::from_le_bytes(bytes[..8].try_into().unwrap()) as usize;
        let o = 8 + n_act;
        println!(
            "{{\"log_adds\": {}, \"log_blowup\": {}, \"num_queries\": {}, \"proof_bytes\": {}, \"proof_sha256\": \"{}\", \
             \"stage_1_commit\": \"{}\", \"stage_2_commit\": \"{}\", \"quotient_commit\": \"{}\"}}",
            c.log_adds,
            c.log_blowup,
            c.num_queries,
            bytes.len(),
            hex(&Sha256::digest(&bytes)),
            hex(&bytes[o..o + 32]),
            hex(&bytes[o + 32..o + 64]),
            hex(&bytes[o + 64..o + 96]),
        );
    }
}

fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{:02x}", x)).collect()
}
