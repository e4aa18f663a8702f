"""CPU oracle end to end: the restated prover (src/prover.rs:289-603) against the restated verifier
(src/verifier.rs:208-695), mirroring the reference's accept / reject tests (src/verifier.rs:783-921,
src/lookup.rs:1043-1130, src/system.rs:424). Runs without a GPU."""
import copy

import numpy as np
import pytest

from tests import _oracle as orc
from tests import _proof

P = orc.P


def workload(kind, log_rows):
    import multi_stark_b200.system as mss
    if kind == "fib":
        return [mss.fib_trace(1 << log_rows)], []
    byte, add, claims = mss.u32_add_workload(1 << log_rows)
    if kind == "u32_add":
        return [byte, add], list(claims)
    return [mss.fib_trace(1 << max(log_rows - 1, 1)), byte, add], list(claims)


@pytest.fixture(scope="module")
def proved(oracle):
    """One small U32-add proof shared by the negative tests."""
    S = orc.OracleSystem(oracle, "u32_add", log_blowup=1, num_queries=30)
    traces, claims = workload("u32_add", 5)
    proof, _ = S.prove(traces, claims)
    yield S, traces, claims, proof
    S.close()


@pytest.mark.parametrize("kind,log_rows,lb,fp_len,pow_bits", [
    ("u32_add", 4, 1, 0, 0), ("u32_add", 7, 2, 0, 0), ("fib", 5, 1, 0, 0), ("fib", 3, 2, 1, 0), ("mixed", 6, 2, 0, 0),
    ("u32_add", 5, 1, 0, 3), ("mixed", 5, 3, 2, 0)])
def test_prove_verify_accepts(oracle, kind, log_rows, lb, fp_len, pow_bits):
    if kind in ("fib", "mixed") and lb < 1:
        pytest.skip("fib_cubic needs quotient degree 2")
    S = orc.OracleSystem(oracle, kind, log_blowup=lb, log_final_poly_len=fp_len, num_queries=25, commit_pow_bits=pow_bits,
                         query_pow_bits=pow_bits)
    traces, claims = workload(kind, log_rows)
    proof, _ = S.prove(traces, claims)
    assert S.verify(claims, proof) == "Ok"
    # Proof::to_bytes / from_bytes round trip through the independent Python restatement of the wire format
    assert _proof.serialize(_proof.parse(proof)) == proof
    # determinism
    proof2, _ = S.prove(traces, claims)
    assert proof2 == proof
    S.close()


def test_sparse_inactive_circuit_accepted(oracle):
    # src/lookup.rs:1057: a circuit with an empty trace is deactivated
    S = orc.OracleSystem(oracle, "mixed", log_blowup=2, num_queries=20)
    traces, claims = workload("mixed", 5)
    traces[0] = np.zeros((0, 3), dtype=np.uint64)
    proof, _ = S.prove(traces, claims)
    pr = _proof.parse(proof)
    assert pr["active"] == [0, 1, 1]
    assert S.verify(claims, proof) == "Ok"
    # flipping the bitmap is rejected (src/lookup.rs:1079)
    pr["active"] = [1, 1, 1]
    assert S.verify(claims, _proof.serialize(pr)) != "Ok"
    S.close()


def test_sparse_needed_circuit_rejected(oracle):
    # src/lookup.rs:1100: deactivating the byte table leaves the U32 circuit's sends unmatched
    S = orc.OracleSystem(oracle, "u32_add", log_blowup=1, num_queries=20)
    traces, claims = workload("u32_add", 4)
    traces[0] = np.zeros((0, 1), dtype=np.uint64)
    proof, _ = S.prove(traces, claims)
    assert S.verify(claims, proof) == "UnbalancedChannel"
    S.close()


def test_wrong_claim_rejected(proved):
    S, _, claims, proof = proved
    bad = [c.copy() for c in claims]
    bad[3][3] = (int(bad[3][3]) + 1) % P
    assert S.verify(bad, proof) != "Ok"
    assert S.verify(claims[:-1], proof) != "Ok"
    # splitting a claim changes the transcript (src/lookup.rs:1118)
    split = [claims[0][:2], claims[0][2:]] + [c for c in claims[1:]]
    assert S.verify(split, proof) != "Ok"


def test_tampered_stage_1_values_rejected(proved):
    S, _, claims, proof = proved
    pr = _proof.parse(proof)
    v = pr["stage_1_opened_values"][1][0][2]
    pr["stage_1_opened_values"][1][0][2] = ((v[0] + 1) % P, v[1])
    assert S.verify(claims, _proof.serialize(pr)) != "Ok"


def test_tampered_accumulator_rejected(proved):
    S, _, claims, proof = proved
    pr = _proof.parse(proof)
    pr2 = copy.deepcopy(pr)
    pr2["intermediate_accumulators"][-1] = (1, 0)
    assert S.verify(claims, _proof.serialize(pr2)) == "UnbalancedChannel"
    pr3 = copy.deepcopy(pr)
    a = pr3["intermediate_accumulators"][0]
    pr3["intermediate_accumulators"][0] = ((a[0] + 1) % P, a[1])
    assert S.verify(claims, _proof.serialize(pr3)) != "Ok"


def test_log_degrees_rejected(proved):
    S, _, claims, proof = proved
    pr = _proof.parse(proof)
    t = copy.deepcopy(pr)
    t["log_degrees"] = t["log_degrees"][:-1]
    assert S.verify(claims, _proof.serialize(t)) == "InvalidProofShape"
    o = copy.deepcopy(pr)
    o["log_degrees"][0] = 200
    assert S.verify(claims, _proof.serialize(o)) == "InvalidProofShape"


def test_truncated_and_tampered_proof_rejected(proved):
    S, _, claims, proof = proved
    assert S.verify(claims, proof[:-8]) == "Deserialize"
    assert S.verify(claims, proof[:len(proof) // 2]) == "Deserialize"
    pr = _proof.parse(proof)
    # a sibling digest of a query, a FRI sibling value, the final polynomial
    for mutate in (lambda p: p["opening_proof"]["query_proofs"][0]["input_proof"][0]["opening_proof"].__setitem__(0, b"\x00" * 32),
                   lambda p: p["opening_proof"]["query_proofs"][1]["commit_phase_openings"][0]["sibling_values"].__setitem__(0, (5, 6)),
                   lambda p: p["opening_proof"]["final_poly"].__setitem__(0, (1, 2)),
                   lambda p: p["opening_proof"]["commit_phase_commits"].__setitem__(1, b"\x11" * 32)):
        t = copy.deepcopy(pr)
        mutate(t)
        assert S.verify(claims, _proof.serialize(t)) == "InvalidOpeningArgument"


def test_quotient_value_tamper_is_ood_mismatch(oracle):
    # the OOD check itself (src/verifier.rs:523-527): a proof for a trace that violates a constraint
    S = orc.OracleSystem(oracle, "fib", log_blowup=2, num_queries=10)
    traces, claims = workload("fib", 4)
    traces[0] = traces[0].copy()
    traces[0][5, 2] = (int(traces[0][5, 2]) + 1) % P  # c != a*b*b on row 5
    try:
        proof, _ = S.prove(traces, claims)
    except RuntimeError:
        return  # the prover noticed the quotient is not a polynomial (final FRI polynomial degree check)
    assert S.verify(claims, proof) in ("OodEvaluationMismatch", "InvalidOpeningArgument")
    S.close()


def test_multi_circuit_system_balances_and_verifies(oracle):
    """BASELINE configs[3] shape: K U32-add circuits of different heights sharing the byte table; the byte multiplicities of
    all circuits are summed (multi_workload), so the lookup accumulator closes and the restated verifier accepts."""
    import multi_stark_b200.system as mss
    traces, claims = mss.multi_workload([6, 4, 5])
    assert [t.shape for t in traces] == [(256, 1), (64, 14), (16, 14), (32, 14)] and claims.shape == (64 + 16 + 32, 4)
    assert int(traces[0].sum()) == 12 * claims.shape[0]      # 12 byte lookups per addition
    S = orc.OracleSystem(oracle, "multi:3", log_blowup=1, num_queries=8)
    proof, _ = S.prove(traces, list(claims))
    assert S.verify(list(claims), proof) == "Ok"
    # dropping one circuit's claims unbalances the accumulator
    assert S.verify(list(claims[:64]), proof) != "Ok"
    S.close()


def test_commitment_wire_switch_is_one_setting():
    """the serde shape of p3's `Commitment` at the pinned rev cannot be checked here. Every commitment in
    Proof::to_bytes goes through write_commitment / read_commitment (host/proof.hpp); MSH_COMMITMENT_WIRE=cap switches all of them
    to the Merkle-cap form (u64 length 1 + digest) and nothing else in the bytes moves."""
    import os
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); from tests import _oracle as orc; import multi_stark_b200.system as mss; "
            "L = orc.lib(); S = orc.OracleSystem(L, 'u32_add', log_blowup=1, num_queries=8); "
            "b, a, c = mss.u32_add_workload(16); p, _ = S.prove([b, a], list(c)); "
            "assert S.verify(list(c), p) == 'Ok'; sys.stdout.write(p.hex())") % orc.ROOT
    raw = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, MSH_COMMITMENT_WIRE="raw"))
    cap = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, MSH_COMMITMENT_WIRE="cap"))
    assert raw.returncode == 0 and cap.returncode == 0, raw.stderr + cap.stderr
    praw, pcap = bytes.fromhex(raw.stdout), bytes.fromhex(cap.stdout)
    a = _proof.parse(praw)
    _proof.CAP_WIRE = True
    try:
        b = _proof.parse(pcap)
        assert _proof.serialize(b) == pcap
    finally:
        _proof.CAP_WIRE = False
    assert a == b
    assert len(pcap) == len(praw) + 8 * (3 + len(a["opening_proof"]["commit_phase_commits"]))
