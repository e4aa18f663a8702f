"""The independent Python restatement (tests/_pyverifier.py: challenger, graph compiler, MMCS, FRI and STARK verifier written
from the reference text, sharing nothing with multi_stark_b200/host/ or oracle/) against the C++ layers:

  * the challenger scripts of src/types.rs:285-319 and random scripts: PyChallenger == msh_challenger_* byte for byte;
  * compile(): node vector, roots, lookups and prefix of the benchmark circuits == the product's graph compiler;
  * proofs of the C++ prover: every challenge the host transcript produced (beta, gamma, alpha, zeta, alpha_pcs, FRI betas,
    query indices) == the ones the Python transcript re-derives from the proof's commitments, and the Python verifier
    accepts them (and rejects tampered ones).
CPU only; tests/test_gpu_prove.py runs the same verifier on device proofs."""
import ctypes as C

import numpy as np
import pytest

from tests import _oracle as orc
from tests import _proof
from tests import _pyverifier as pv


def params(**kw):
    d = dict(log_blowup=1, log_final_poly_len=0, max_log_arity=1, num_queries=100, commit_pow_bits=0, query_pow_bits=0)
    d.update(kw)
    return d


def test_pychallenger_matches_host_challenger():
    import multi_stark_b200.pcs as mpcs
    rng = np.random.default_rng(5)
    for case in range(6):
        kw = dict(log_blowup=int(rng.integers(1, 4)), log_final_poly_len=int(rng.integers(0, 3)), num_queries=int(rng.integers(1, 120)),
                  commit_pow_bits=int(rng.integers(0, 3)), query_pow_bits=int(rng.integers(0, 3)))
        host = mpcs.Challenger(**kw)
        mine = pv.PyChallenger.for_config(kw["log_blowup"], 0, kw["log_final_poly_len"], 1, kw["num_queries"], kw["commit_pow_bits"],
                                          kw["query_pow_bits"])
        for step in range(40):
            op = int(rng.integers(0, 3))
            if op == 0:
                d = bytes(rng.integers(0, 256, size=32, dtype=np.uint8))
                host.observe(d)
                mine.observe(d)
            elif op == 1:
                vals = rng.integers(0, pv.P, size=int(rng.integers(1, 9)), dtype=np.uint64)
                host.observe_values(vals)
                for v in vals:
                    mine.observe(int(v))
            else:  # several samples in a row cross the 32-byte output buffer (4 u64 per flush)
                for _ in range(int(rng.integers(1, 4))):
                    e = mine.sample_ext()
                    assert host.sample_algebra_element() == (e.a, e.b)
        host.close()


def test_reference_generator_scripts_against_golden():
    """The scripts of the reference's own generators (gen_pcs_refs / gen_challenger_refs, src/types.rs:246-319), replayed by
    the pure-Python restatement, reproduce the records in tests/golden/pcs_refs.json (which the C++ oracle wrote): leaf and
    compression digests, the 3-matrix MMCS opened at 5 (verify_batch against COMMIT), and the challenger continuation."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pcs_refs.json")))

    def limbs(d):
        return [int.from_bytes(d[8 * i:8 * i + 8], "little") for i in range(4)]

    def dig(xs):
        return b"".join(int(x).to_bytes(8, "little") for x in xs)
    for n in (3, 17, 22, 20):
        assert limbs(pv.hash_rows([list(range(1, n + 1))])) == g["LEAF%d" % n]
    assert limbs(pv.compress(dig([1, 2, 3, 4]), dig([5, 6, 7, 8]))) == g["COMPRESS"]
    proof = [dig(g["SIB%d" % i]) for i in range(3)]
    assert pv.verify_batch(dig(g["COMMIT"]), [8, 4, 2], 5, g["OPENED"], proof)
    assert not pv.verify_batch(dig(g["COMMIT"]), [8, 4, 2], 4, g["OPENED"], proof)
    ch = pv.PyChallenger(b"")
    ch.observe(0x0102030405060708)
    assert ch.sample_bits(20) == g["SAMPLE_BITS"]
    ch = pv.PyChallenger(b"")
    ch.observe(0x0102030405060708)
    ch.observe(0x1122334455667788)
    apcs, afri = ch.sample_ext(), ch.sample_ext()
    assert [apcs.a, apcs.b] == g["APCS"] and [afri.a, afri.b] == g["AFRI"]
    ch.observe(0x00000000deadbeef)
    beta = ch.sample_ext()
    assert [beta.a, beta.b] == g["BETA"]
    ch.observe(0x0a0b0c0d01020304)
    ch.observe(2)
    assert ch.sample_bits(20) == g["SAMPLE_BITS2"]


def _graph_from_desc(system, i):
    """Read msgpu_graph_desc (include/msgpu.h) of circuit i through ctypes into the tuple form of _pyverifier."""
    class Desc(C.Structure):
        _fields_ = [("n_nodes", C.c_uint32), ("op", C.POINTER(C.c_uint8)), ("a", C.POINTER(C.c_uint32)), ("b", C.POINTER(C.c_uint32)),
                    ("imm", C.POINTER(C.c_uint64)), ("n_zeros", C.c_uint32), ("zeros", C.POINTER(C.c_uint32)),
                    ("n_lookups", C.c_uint32), ("lookup_mult", C.POINTER(C.c_uint32)), ("lookup_arg_off", C.POINTER(C.c_uint32)),
                    ("lookup_args", C.POINTER(C.c_uint32)), ("lookup_prefix_len", C.c_uint32), ("pre_width", C.c_uint32),
                    ("main_width", C.c_uint32), ("stage2_width", C.c_uint32)]
    d = C.cast(system.graph_desc(i), C.POINTER(Desc)).contents
    nodes = []
    for k in range(d.n_nodes):
        op, a, b, imm = d.op[k], d.a[k], d.b[k], d.imm[k]
        if op == 0:
            nodes.append(("const", imm))
        elif op == 1:
            nodes.append(("var", a & 3, a >> 2, b))
        elif op == 2:
            nodes.append(("public", a))
        elif op in (3, 4, 5):
            nodes.append(({3: "first", 4: "last", 5: "trans"}[op],))
        elif op in (6, 7, 8):
            nodes.append(({6: "add", 7: "sub", 8: "mul"}[op], a, b))
        else:
            nodes.append(("neg", a))
    zeros = [d.zeros[k] for k in range(d.n_zeros)]
    lookups = []
    for j in range(d.n_lookups):
        lookups.append((d.lookup_mult[j], [d.lookup_args[k] for k in range(d.lookup_arg_off[j], d.lookup_arg_off[j + 1])]))
    return nodes, zeros, lookups, d.lookup_prefix_len


@pytest.mark.parametrize("kind", ["u32_add", "multi:3"])
def test_python_compile_matches_product_graph(kind):
    """compile() interning order (src/graph.rs:120-188): the node vector IS the wire between host and device."""
    import multi_stark_b200.system as mss
    S = mss.System(kind)
    mine = pv.named_system(kind)
    assert S.num_circuits == len(mine)
    for i, c in enumerate(mine):
        nodes, zeros, lookups, prefix = _graph_from_desc(S, i)
        assert c.graph.nodes == nodes
        assert c.graph.zeros == zeros
        assert [(m, list(a)) for m, a in c.graph.lookups] == lookups
        assert c.graph.lookup_prefix_len == prefix
        info = S.circuits[i]
        assert (c.constraint_count, c.max_constraint_degree, c.quotient_degree(), c.stage2_width) == \
               (info["constraint_count"], info["max_constraint_degree"], info["quotient_degree"], info["stage2_width"])
    # the U32-add DAG has 48 nodes (14 Var, 14 Add, 11 Mul, 6 Const, 2 Sub, 1 Neg), lookup prefix 37, 2 roots
    g = mine[1].graph
    kinds = [n[0] for n in g.nodes]
    assert (len(g.nodes), g.lookup_prefix_len, len(g.zeros)) == (48, 37, 2)
    assert [kinds.count(k) for k in ("var", "add", "mul", "const", "sub", "neg")] == [14, 14, 11, 6, 2, 1]
    S.close()


def _workload(log_rows):
    import multi_stark_b200.system as mss
    byte, add, claims = mss.u32_add_workload(1 << log_rows)
    return [byte, add], [list(map(int, c)) for c in claims]


CASES = [dict(log_rows=4), dict(log_rows=6, log_blowup=2), dict(log_rows=5, commit_pow_bits=3, query_pow_bits=2),
         dict(log_rows=5, log_blowup=2, log_final_poly_len=2, num_queries=40), dict(log_rows=9, num_queries=30)]


@pytest.mark.parametrize("case", CASES)
def test_python_verifier_accepts_oracle_proofs_and_rederives_challenges(oracle, case):
    case = dict(case)
    log_rows = case.pop("log_rows")
    prm = params(num_queries=25)
    prm.update(case)
    S = orc.OracleSystem(oracle, "u32_add", **prm)
    traces, claims = _workload(log_rows)
    proof, _ = S.prove(traces, claims)
    want_ch, want_idx = S.last_transcript()
    got = {}
    assert pv.verify(pv.named_system("u32_add"), prm, S.preprocessed_commit(), claims, _proof.parse(proof), got) == "Ok"
    mine = [got["beta"], got["gamma"], got["alpha"], got["zeta"], got["alpha_pcs"]] + got["fri_betas"]
    assert [(e.a, e.b) for e in mine] == want_ch
    assert got["query_indices"] == want_idx
    S.close()


def test_python_verifier_rejects_tampering(oracle):
    prm = params(num_queries=20)
    S = orc.OracleSystem(oracle, "u32_add", **prm)
    traces, claims = _workload(5)
    proof, _ = S.prove(traces, claims)
    circuits, pre = pv.named_system("u32_add"), S.preprocessed_commit()
    pr = _proof.parse(proof)
    assert pv.verify(circuits, prm, pre, claims, pr) == "Ok"
    bad = [list(c) for c in claims]
    bad[3][1] ^= 1
    assert pv.verify(circuits, prm, pre, bad, pr) != "Ok"                         # src/verifier.rs:852
    t = _proof.parse(proof)
    t["stage_1_opened_values"][1][0][2] = ((t["stage_1_opened_values"][1][0][2][0] + 1) % pv.P, 0)
    assert pv.verify(circuits, prm, pre, claims, t) == "InvalidOpeningArgument"
    t = _proof.parse(proof)
    t["quotient_chunks"] = bytes(32)
    assert pv.verify(circuits, prm, pre, claims, t) != "Ok"
    t = _proof.parse(proof)
    t["opening_proof"]["final_poly"][0] = (1, 2)
    assert pv.verify(circuits, prm, pre, claims, t) == "InvalidOpeningArgument"
    t = _proof.parse(proof)
    t["intermediate_accumulators"][-1] = (1, 0)
    assert pv.verify(circuits, prm, pre, claims, t) == "UnbalancedChannel"
    t = _proof.parse(proof)
    q = t["opening_proof"]["query_proofs"][0]["commit_phase_openings"][0]
    q["sibling_values"][0] = ((q["sibling_values"][0][0] + 1) % pv.P, q["sibling_values"][0][1])
    assert pv.verify(circuits, prm, pre, claims, t) == "InvalidOpeningArgument"
    # a different parameter set seeds a different transcript (src/types.rs:118-130)
    prm2 = dict(prm, num_queries=prm["num_queries"])
    prm2["log_final_poly_len"] = 1
    assert pv.verify(circuits, prm2, pre, claims, pr) != "Ok"
    S.close()


def test_python_verifier_multi_circuit(oracle):
    import multi_stark_b200.system as mss
    prm = params(num_queries=15, log_blowup=1)
    S = orc.OracleSystem(oracle, "multi:3", **prm)
    traces, claims = mss.multi_workload([6, 5, 4])
    claims = [list(map(int, c)) for c in claims]
    proof, _ = S.prove(traces, claims)
    assert S.verify(claims, proof) == "Ok"
    assert pv.verify(pv.named_system("multi:3"), prm, S.preprocessed_commit(), claims, _proof.parse(proof)) == "Ok"
    S.close()
