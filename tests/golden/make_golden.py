"""Writes the golden fixtures of tests/golden/ (run from the repository root: python tests/golden/make_golden.py).

The reference is Rust with an un-vendored Plonky3 dependency and cannot run in this image, so these vectors come from the
ORACLE (oracle/, the CPU restatement; leaf / compress values independently from the `blake3` Python package). They serve two
purposes: (1) regression pins -- the oracle (CPU tests) and the CUDA path (GPU tests) must keep reproducing them bit for bit;
(2) a diff target for a machine with cargo: `pcs_refs.json` uses the labels that the reference's own generators print
(`cargo test gen_pcs_refs gen_challenger_refs -- --nocapture`, src/types.rs:246-319), and `integration/golden_dump.rs` prints
the commitments / proof digests of `proofs.json` from the real prover. Until such a diff has been made the MMCS, transcript and
FRI entries are "predicted, unpinned against p3" (DESIGN.md section 5)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import blake3  # noqa: E402
import numpy as np  # noqa: E402

from tests import _oracle as orc  # noqa: E402
import multi_stark_b200.system as mss  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def limbs(d):
    return [int.from_bytes(d[8 * i:8 * i + 8], "little") for i in range(4)]


def pcs_refs(L):
    out = {"_source": "oracle + python blake3; labels of src/types.rs:246-319 (gen_pcs_refs, gen_challenger_refs)"}
    for n in (3, 17, 22, 20):
        row = b"".join(int(v).to_bytes(8, "little") for v in range(1, n + 1))
        out["LEAF%d" % n] = limbs(blake3.blake3(row).digest())
    dig = lambda xs: b"".join(int(x).to_bytes(8, "little") for x in xs)  # noqa: E731
    out["COMPRESS"] = limbs(blake3.blake3(dig([1, 2, 3, 4]) + dig([5, 6, 7, 8])).digest())
    m0 = np.zeros((8, 2), dtype=np.uint64); m0[5] = [11, 12]
    m1 = np.zeros((4, 3), dtype=np.uint64); m1[2] = [107, 108, 109]
    m2 = np.zeros((2, 1), dtype=np.uint64); m2[1] = [202]
    t = orc.MmcsTree(L, [m0, m1, m2])
    opened, proof = t.open(5)
    out["OPENED"] = [[11, 12], [107, 108, 109], [202]]
    assert [int(x) for x in opened] == [11, 12, 107, 108, 109, 202]
    for i, s in enumerate(proof):
        out["SIB%d" % i] = limbs(bytes(s))
    out["COMMIT"] = limbs(bytes(t.root))
    ch = np.zeros(8, dtype=np.uint64)
    L.orc_gen_challenger_refs(ch)
    out["SAMPLE_BITS"] = int(ch[0])
    out["APCS"] = [int(ch[1]), int(ch[2])]
    out["AFRI"] = [int(ch[3]), int(ch[4])]
    out["BETA"] = [int(ch[5]), int(ch[6])]
    out["SAMPLE_BITS2"] = int(ch[7])
    return out


PROOF_CASES = [
    dict(kind="u32_add", log_adds=4, params=dict(log_blowup=1, num_queries=10)),
    dict(kind="u32_add", log_adds=8, params=dict(log_blowup=1, num_queries=100)),
    dict(kind="u32_add", log_adds=12, params=dict(log_blowup=1, num_queries=100)),   # BASELINE configs[0]
    dict(kind="u32_add", log_adds=10, params=dict(log_blowup=2, num_queries=30, commit_pow_bits=10, query_pow_bits=10)),  # the bench's own parameters
    dict(kind="u32_add", log_adds=9, params=dict(log_blowup=3, num_queries=20, log_final_poly_len=2)),
]


def proofs(L):
    out = {"_source": "oracle prover (oracle/, restatement of src/prover.rs:289-603); sha256 of Proof::to_bytes",
           "cases": []}
    for case in PROOF_CASES:
        S = orc.OracleSystem(L, case["kind"], **case["params"])
        byte, add, claims = mss.u32_add_workload(1 << case["log_adds"])
        proof, _ = S.prove([byte, add], list(claims))
        assert S.verify(list(claims), proof) == "Ok"
        rec = dict(case)
        rec["proof_bytes"] = len(proof)
        rec["proof_sha256"] = hashlib.sha256(proof).hexdigest()
        rec["preprocessed_commit"] = S.preprocessed_commit().hex()
        # the three per-proof commitments sit right after the activation vector: u64 len + len bytes, then 3 x 32 bytes
        n_act = int.from_bytes(proof[:8], "little")
        o = 8 + n_act
        rec["stage_1_commit"] = proof[o:o + 32].hex()
        rec["stage_2_commit"] = proof[o + 32:o + 64].hex()
        rec["quotient_commit"] = proof[o + 64:o + 96].hex()
        out["cases"].append(rec)
        S.close()
    return out


def commits(L):
    """Pcs::commit roots of seeded matrices (the headline path), log_blowup 1..2"""
    out = {"_source": "oracle Pcs::commit (coset LDE + MMCS), numpy default_rng(seed) matrices mod p", "cases": []}
    for seed, shapes, lb in [(1, [(256, 1), (4096, 14)], 1), (2, [(256, 2), (4096, 26)], 1), (3, [(1024, 5), (64, 3), (1024, 2)], 2),
                             (4, [(128, 200)], 1)]:
        rng = np.random.default_rng(seed)
        mats = [orc.rand_matrix(rng, h, w) for h, w in shapes]
        root, h = orc.pcs_commit(L, mats, lb)
        L.orc_mmcs_free(h)
        out["cases"].append({"seed": seed, "shapes": shapes, "log_blowup": lb, "root": bytes(root).hex()})
    return out


if __name__ == "__main__":
    L = orc.lib()
    for name, fn in (("pcs_refs.json", pcs_refs), ("proofs.json", proofs), ("commits.json", commits)):
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(fn(L), f, indent=1)
        print("wrote", name)
