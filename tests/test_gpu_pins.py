"""The reference's own relational pins for selectors and logUp, mirrored on the CUDA path:

  * selector_normalization_constants (src/lookup.rs:697-756): the tables `k_selectors` builds for the quotient kernel
    (p3's UNNORMALISED Lagrange selectors on the quotient coset) against the textbook Lagrange basis products, with the
    constants 1/(n g) and 1/n the logUp boundary injection relies on;
  * direct_logup_matches_synthesized_reference (src/lookup.rs:763-867): `k_quotient_eval` -- bytecode sweep, direct logUp
    values, reversed-alpha fold, division by Z_H -- against a SCHOOLBOOK evaluation of the constraints `synthesize_lookups`
    (src/lookup.rs:280-330) specifies, computed in pure Python from the frontend expression trees (no compiled graph, no
    Karatsuba, big ints), on the reference's assorted lookup shapes (multi-argument with a product, single argument pull, empty
    arguments) at every point of the quotient coset.
Both go through the C ABI (msgpu_selectors_on_coset, msgpu_program_create from a caller-made descriptor, msgpu_quotient)."""
import ctypes as C

import numpy as np
import pytest

from tests import _oracle as orc
from tests import _pyverifier as pv

pytestmark = pytest.mark.gpu
P = pv.P


@pytest.fixture(scope="module")
def gpu():
    import multi_stark_b200 as ms
    ctx = ms.GpuContext(0)
    yield ms, ctx
    ctx.close()


def device_selectors(ctx, log_n, log_q):
    nq = 1 << (log_n + log_q)
    f, l, z = (np.zeros(nq, dtype=np.uint64) for _ in range(3))
    from multi_stark_b200._ffi import check
    check(ctx.L.msgpu_selectors_on_coset(ctx.h, log_n, log_q, f.ctypes.data_as(C.c_void_p), l.ctypes.data_as(C.c_void_p),
                                         z.ctypes.data_as(C.c_void_p)))
    return f, l, z


@pytest.mark.parametrize("log_n", [2, 3, 5, 8])
@pytest.mark.parametrize("log_q", [0, 1, 2])
def test_selector_normalization_constants(gpu, log_n, log_q):
    ms, ctx = gpu
    n, nq = 1 << log_n, 1 << (log_n + log_q)
    g = pv.two_adic_generator(log_n)
    first, last, inv_zh = device_selectors(ctx, log_n, log_q)
    gi = [pow(g, i, P) for i in range(n)]
    # denominators of the textbook Lagrange bases of the last and the first row
    den_last = 1
    for i in range(n - 1):
        den_last = den_last * (gi[n - 1] - gi[i]) % P
    den_first = 1
    for i in range(1, n):
        den_first = den_first * (1 - gi[i]) % P
    inv_den_last, inv_den_first = pv.inv(den_last), pv.inv(den_first)
    norm_last, norm_first = pv.inv(n * g), pv.inv(n)
    w = pv.two_adic_generator(log_n + log_q)
    for i in range(nq):
        x = pv.GENERATOR * pow(w, i, P) % P
        ref_last = inv_den_last
        for k in range(n - 1):
            ref_last = ref_last * (x - gi[k]) % P
        ref_first = inv_den_first
        for k in range(1, n):
            ref_first = ref_first * (x - gi[k]) % P
        assert int(last[i]) * norm_last % P == ref_last, "last-row normalisation, log_n=%d point %d" % (log_n, i)
        assert int(first[i]) * norm_first % P == ref_first, "first-row normalisation, log_n=%d point %d" % (log_n, i)
        assert int(inv_zh[i]) * (pow(x, n, P) - 1) % P == 1


# ---- schoolbook evaluation of the synthesized logUp constraints -------------------------------------------------------
def eval_expr(e, row, nxt):
    """src/eval.rs eval_expr over base values: the frontend tree, not the compiled graph."""
    k = e.k
    if k == "const":
        return e.x
    if k == "var":
        src, off, idx = e.x
        assert src == pv.MAIN
        return int((nxt if off else row)[idx])
    if k == "neg":
        return -eval_expr(e.x, row, nxt) % P
    a, b = eval_expr(e.x[0], row, nxt), eval_expr(e.x[1], row, nxt)
    return {"add": a + b, "sub": a - b, "mul": a * b}[k] % P


def ext_mul_schoolbook(a, b):
    return ((a[0] * b[0] + pv.W * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def synthesized_logup(lookups, row, nxt, s2, s2n, publics, is_last_norm):
    """src/lookup.rs:280-330 evaluated coordinate by coordinate: publics = (beta, gamma, acc_initial, acc_final) as 8 base
    values, is_last_norm = the NORMALISED last-row selector (value 1 on the last row)."""
    beta, gamma = (publics[0], publics[1]), (publics[2], publics[3])
    inj = tuple(is_last_norm * (publics[6 + k] - publics[4 + k]) % P for k in range(2))
    out = []
    last = len(lookups) - 1
    for j, (mult, args) in enumerate(lookups):
        source = (int(s2[2 * j]), int(s2[2 * j + 1]))
        target = (int(s2[2 * j + 2]), int(s2[2 * j + 3])) if j < last else ((int(s2n[0]) + inj[0]) % P, (int(s2n[1]) + inj[1]) % P)
        rev = list(reversed(args))
        f = (eval_expr(rev[0], row, nxt), 0) if rev else (0, 0)
        for a in rev[1:]:
            f = ext_mul_schoolbook(f, gamma)
            f = ((f[0] + eval_expr(a, row, nxt)) % P, f[1])
        msg = ((beta[0] + f[0]) % P, (beta[1] + f[1]) % P)
        c = ext_mul_schoolbook(msg, ((target[0] - source[0]) % P, (target[1] - source[1]) % P))
        out += [(c[0] - eval_expr(mult, row, nxt)) % P, c[1]]
    return out


@pytest.mark.parametrize("log_n,lb", [(3, 1), (5, 2), (7, 1)])
def test_direct_logup_matches_synthesized_reference(gpu, log_n, lb):
    ms, ctx = gpu
    E = pv.Expr
    m = E.main
    # the lookup shapes of src/lookup.rs:776-791, plus one user constraint so that roots and logUp values share the fold
    lookups = [pv.push(m(0), [E.const(7), m(1), m(2) * m(3)]), pv.pull(m(4), [m(5)]), (E.const(1), [])]
    user = [m(0) * m(1) - m(2)]
    circ = pv.Circuit(6, lookups, user)
    assert circ.max_constraint_degree == 3 and circ.quotient_degree() == 2
    system = ms.System.from_graphs([pv.graph_dict(circ)], log_blowup=lb)
    assert system.circuits[0]["quotient_degree"] == 2 and system.circuits[0]["stage2_width"] == 6
    rng = np.random.default_rng(17 + log_n)
    n, log_q = 1 << log_n, 1
    nq = n << log_q
    main = orc.rand_matrix(rng, n, 6)
    s2 = orc.rand_matrix(rng, n, 6)
    pcs = ms.GpuPcs(ctx, lb)
    _, pd1 = pcs.commit([main])
    _, pd2 = pcs.commit([s2])
    alpha = rng.integers(0, P, size=2, dtype=np.uint64)
    publics = rng.integers(0, P, size=8, dtype=np.uint64)
    prog = ms.Program(ctx, system, 0)
    lde, rows, cols, got = prog.quotient(None, 0, pd1, 0, pd2, 0, log_n, log_q, lb, publics, alpha, want_values=True)
    ctx.free(lde)
    # stored row s of the committed LDE holds the evaluation at GENERATOR * w^{rev(s)}: the quotient coset is the first nq rows
    m_lde, s_lde = pd1.read_rows(0, 0, nq), pd2.read_rows(0, 0, nq)
    log_nq = log_n + log_q
    nat = [pv.rev_bits(i, log_nq) for i in range(nq)]
    g = pv.two_adic_generator(log_n)
    ginv = pv.inv(g)
    w = pv.two_adic_generator(log_nq)
    norm_last = pv.inv(n * g)
    a = pv.E(int(alpha[0]), int(alpha[1]))
    pub = [int(v) for v in publics]
    for i in range(nq):
        x = pv.GENERATOR * pow(w, i, P) % P
        zh = (pow(x, n, P) - 1) % P
        is_last_norm = zh * pv.inv(x - ginv) % P * norm_last % P
        cur, nxt = nat[i], nat[(i + (1 << log_q)) % nq]   # next trace row = q points further on the quotient coset
        values = [eval_expr(c, m_lde[cur], m_lde[nxt]) for c in user]
        values += synthesized_logup(lookups, m_lde[cur], m_lde[nxt], s_lde[cur], s_lde[nxt], pub, is_last_norm)
        assert len(values) == circ.constraint_count
        comp = pv.E(0)
        for v in values:                      # sum_j v_j alpha^{k-1-j}
            comp = comp * a + v
        want = comp * pv.inv(zh)
        assert (int(got[i][0]), int(got[i][1])) == (want.a, want.b), "quotient value at point %d" % i
    prog.free()
    pd1.free()
    pd2.free()
    system.close()
