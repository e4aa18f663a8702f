"""`msh_system_create_from_graphs` (the generic System::new for circuits the caller compiled itself, src/system.rs:115-203):
a system assembled from msgpu_graph_desc descriptors -- here produced by the independent Python compiler of
tests/_pyverifier.py -- is the same system as the named one: same shape record, same proof bytes. CPU part (host layer + the
oracle's CPU backend); tests/test_gpu_prove.py repeats it through msh_prove on the device."""
import numpy as np
import pytest

from tests import _oracle as orc
from tests import _proof
from tests import _pyverifier as pv


def u32_graphs():
    cs = pv.named_system("u32_add")
    pre = [np.arange(256, dtype=np.uint64).reshape(256, 1), None]
    return [pv.graph_dict(c) for c in cs], pre


def test_system_from_descriptors_has_the_named_shape():
    import multi_stark_b200.system as mss
    graphs, pre = u32_graphs()
    A = mss.System("u32_add")
    B = mss.System.from_graphs(graphs, pre)
    assert A.circuits == B.circuits
    assert np.array_equal(A.preprocessed(0), B.preprocessed(0))
    A.close()
    B.close()


def test_descriptor_validation():
    import multi_stark_b200.system as mss
    graphs, pre = u32_graphs()
    bad = dict(graphs[1])
    bad["nodes"] = list(bad["nodes"])
    bad["nodes"][10] = ("add", 40, 3)        # child after parent
    with pytest.raises(ValueError, match="children must precede"):
        mss.System.from_graphs([graphs[0], bad], pre)
    bad = dict(graphs[1], zeros=list(reversed(graphs[1]["zeros"])))
    with pytest.raises(ValueError, match="sorted"):
        mss.System.from_graphs([graphs[0], bad], pre)
    with pytest.raises(ValueError, match="preprocessed"):
        mss.System.from_graphs(graphs, [None, None])
    bad = dict(graphs[1])
    bad["nodes"] = [("const", pv.P)] + list(bad["nodes"][1:])
    with pytest.raises(ValueError, match="canonical"):
        mss.System.from_graphs([graphs[0], bad], pre)


def test_oracle_proof_from_descriptors_equals_named(oracle):
    import multi_stark_b200.system as mss
    graphs, pre = u32_graphs()
    byte, add, claims = mss.u32_add_workload(1 << 6)
    claims = [list(map(int, c)) for c in claims]
    kw = dict(log_blowup=1, num_queries=20)
    A = orc.OracleSystem(oracle, "u32_add", **kw)
    B = orc.OracleSystem(oracle, None, graphs=graphs, preprocessed=pre, **kw)
    pa, _ = A.prove([byte, add], claims)
    pb, _ = B.prove([byte, add], claims)
    assert pa == pb
    assert B.verify(claims, pb) == "Ok"
    A.close()
    B.close()


def test_custom_circuit_from_descriptors_proves_and_verifies(oracle):
    """A circuit that exists only as a descriptor: columns (a, b, c) with c = a * b + 3 on every row and the transition
    a' = a + 1, compiled by the Python compiler; proved by the oracle's CPU prover from the descriptor, accepted by the restated
    verifier (and the node vector round-trips through msh_circuit_graph)."""
    E = pv.Expr
    m = E.main
    trans = E("trans")
    cons = [m(2) - (m(0) * m(1) + E.const(3)), trans * (E.main_next(0) - m(0) - E.const(1))]
    c = pv.Circuit(3, [], cons)
    assert c.max_constraint_degree == 2 and c.quotient_degree() == 1
    g = pv.graph_dict(c)
    n = 32
    rng = np.random.default_rng(3)
    a = np.arange(5, 5 + n, dtype=np.uint64)
    b = rng.integers(0, pv.P, size=n, dtype=np.uint64)
    cc = np.array([(int(x) * int(y) + 3) % pv.P for x, y in zip(a, b)], dtype=np.uint64)
    trace = np.stack([a, b, cc], axis=1)
    S = orc.OracleSystem(oracle, None, graphs=[g], log_blowup=1, num_queries=15)
    proof, _ = S.prove([trace], [])
    assert S.verify([], proof) == "Ok"
    prm = dict(log_blowup=1, log_final_poly_len=0, max_log_arity=1, num_queries=15, commit_pow_bits=0, query_pow_bits=0)
    assert pv.verify([c], prm, None, [], _proof.parse(proof)) == "Ok"
    bad = trace.copy()
    bad[7, 2] = (int(bad[7, 2]) + 1) % pv.P
    # an invalid witness still yields proof bytes (the quotient of n*q evaluations is always interpolated); both verifiers
    # reject them at the out-of-domain check (src/verifier.rs:519-523)
    bad_proof, _ = S.prove([bad], [])
    assert S.verify([], bad_proof) == "OodEvaluationMismatch"
    assert pv.verify([c], prm, None, [], _proof.parse(bad_proof)) == "OodEvaluationMismatch"
    S.close()
