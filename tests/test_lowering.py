"""CPU check of the bytecode lowering the device interpreter runs (multi_stark_b200/csrc/lowering.hpp): for every circuit of the
named systems the lowered program (slot reuse by liveness, column reads materialised at first use, constraint roots folded by
OP_ROOT) is interpreted on random rows next to a node-by-node evaluation of the ConstraintGraph itself -- the reference's
`sweep_range` (src/eval.rs:67-106) -- and must give the same value for every constraint root and every lookup operand."""
import numpy as np
import pytest

from tests import _oracle as orc


@pytest.mark.parametrize("kind,n_circuits", [("u32_add", 2), ("mixed", 3), ("fib", 1), ("wide:256", 1), ("wide:6", 1), ("multi:3", 4)])
def test_lowered_program_equals_graph(oracle, kind, n_circuits):
    S = orc.OracleSystem(oracle, kind, log_blowup=2, num_queries=4)
    for ci in range(n_circuits):
        out = np.zeros(4, dtype=np.uint32)
        bad = int(oracle.orc_check_lowering(S.h, ci, 16, 1234 + ci, out))
        assert bad == 0, "circuit %d of %s: %d mismatches between the lowered program and the graph" % (ci, kind, bad)
        assert out[0] >= 1 and out[1] >= 1
    S.close()


def test_wide_air_runs_in_a_handful_of_slots(oracle):
    """256 columns, 128 degree-3 constraints: with roots folded as they are computed and column reads materialised at first
    use the working set is independent of the width (it was 256 slots = 2 KB of local memory per thread before)."""
    S = orc.OracleSystem(oracle, "wide:256", log_blowup=2, num_queries=4)
    out = np.zeros(4, dtype=np.uint32)
    assert int(oracle.orc_check_lowering(S.h, 0, 2, 7, out)) == 0
    assert out[0] <= 4, "full program needs %d slots" % out[0]
    assert out[1] == 640 + 128, "128 OP_ROOT instructions follow the 640 graph nodes"
    S.close()


def test_u32_add_slot_counts(oracle):
    S = orc.OracleSystem(oracle, "u32_add", log_blowup=1, num_queries=4)
    out = np.zeros(4, dtype=np.uint32)
    assert int(oracle.orc_check_lowering(S.h, 1, 4, 3, out)) == 0
    assert out[0] <= 32 and out[2] <= 32, "the U32-add circuit must fit the 32-slot kernels (shared-memory slots in k_lookup_messages)"
    S.close()


@pytest.mark.parametrize("k,lo,hi", [(300, 257, 1024), (1200, 1025, 4096)])
def test_large_dag_lowering(oracle, k, lo, hi):
    """The synthetic large DAG (tests/_bigdag.py, thousands of nodes, ~k live values) lowers correctly and really needs the
    1024- / 4096-slot instantiations of the device interpreter (tests/test_gpu_quotient.py runs them)."""
    from tests import _bigdag, _pyverifier as pv
    g = pv.graph_dict(_bigdag.big_dag_circuit(k=k))
    S = orc.OracleSystem(oracle, None, graphs=[g], log_blowup=1, num_queries=4)
    out = np.zeros(4, dtype=np.uint32)
    assert int(oracle.orc_check_lowering(S.h, 0, 4, 99, out)) == 0
    assert lo <= out[0] <= hi, "full program needs %d slots" % out[0]
    assert out[2] <= 32, "the lookup prefix stays in the 32-slot kernel"
    S.close()
