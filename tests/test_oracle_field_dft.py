"""Pins the CPU oracle's field and DFT/LDE layer (oracle/cpu_dft.hpp).

Mirrors the reference's relational pinning tests:
  * lde_from_coefficients_matches_commit_transform   (src/prover.rs:975-999)
  * shifted_quotient_slices_matches_naive_composition (src/prover.rs:1006-1041)
and checks the DFT convention / committed-LDE layout of SURVEY Appendix A.2-A.3 against naive
big-int evaluation (tests/_naive.py)."""
import numpy as np
import pytest

from tests import _naive as nv
from tests import _oracle as orc

P = nv.P


def test_field_constants(oracle):
    # p - 1 = 2^32 * 3 * 5 * 17 * 257 * 65537 ; 7 generates the multiplicative group.
    assert (P - 1) == 2**32 * 3 * 5 * 17 * 257 * 65537
    for q in (2, 3, 5, 17, 257, 65537):
        assert pow(7, (P - 1) // q, P) != 1
    # p3's two-adic generator of order 2^32 is 7^((p-1)/2^32).
    assert pow(7, (P - 1) >> 32, P) == nv.ROOT32
    for bits in range(0, 33):
        g = oracle.orc_two_adic_generator(bits)
        assert g == nv.two_adic_generator(bits)
        assert pow(g, 1 << bits, P) == 1
        if bits:
            assert pow(g, 1 << (bits - 1), P) == P - 1
    # 7 is a quadratic non-residue => X^2 - 7 irreducible (ExtVal, src/types.rs:26).
    assert pow(7, (P - 1) // 2, P) == P - 1


def test_field_mul_inv(oracle):
    rng = np.random.default_rng(1)
    edge = [0, 1, 2, P - 1, P - 2, 2**32, 2**32 - 1, 2**63, 0xFFFFFFFF00000000]
    vals = edge + [int(x) for x in rng.integers(0, P, size=200, dtype=np.uint64)]
    for a in vals:
        for b in vals[:40]:
            assert oracle.orc_fp_mul(a, b) == a * b % P
        if a:
            assert oracle.orc_fp_inv(a) == pow(a, P - 2, P)


def test_ext_mul_inv(oracle):
    rng = np.random.default_rng(2)
    for _ in range(200):
        a = rng.integers(0, P, size=2, dtype=np.uint64)
        b = rng.integers(0, P, size=2, dtype=np.uint64)
        out = np.zeros(2, dtype=np.uint64)
        oracle.orc_fp2_mul(a, b, out)
        assert tuple(int(x) for x in out) == nv.e_mul((int(a[0]), int(a[1])), (int(b[0]), int(b[1])))
        oracle.orc_fp2_inv(a, out)
        assert nv.e_mul((int(a[0]), int(a[1])), (int(out[0]), int(out[1]))) == (1, 0)


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 7])
@pytest.mark.parametrize("w", [1, 3])
def test_dft_matches_naive(oracle, log_n, w):
    rng = np.random.default_rng(log_n * 10 + w)
    n = 1 << log_n
    m = orc.rand_matrix(rng, n, w)
    out = np.empty_like(m)
    oracle.orc_dft_batch(m, n, w, out)
    for c in range(w):
        assert [int(x) for x in out[:, c]] == nv.dft([int(x) for x in m[:, c]])
    # raw bit-reversed storage: natural k lives at row rev(k)
    br = orc.dft_bitrev(oracle, m)
    for k in range(n):
        assert (br[nv.rev(k, log_n)] == out[k]).all()
    inv = np.empty_like(m)
    oracle.orc_idft_batch(out, n, w, inv)
    assert (inv == m).all()


@pytest.mark.parametrize("log_n,added", [(0, 1), (1, 1), (2, 2), (3, 1), (3, 2), (4, 3), (6, 1)])
def test_committed_lde_layout(oracle, log_n, added):
    """stored[i] = P(7 * w_{nB}^{rev(i)}) -- SURVEY A.3(1); first n*q rows bit-reversed = natural
    order on the coset 7*H_{nq} -- A.3(3) (get_evaluations_on_domain, src/prover.rs:454-468)."""
    rng = np.random.default_rng(100 + log_n + added)
    n = 1 << log_n
    m = orc.rand_matrix(rng, n, 2)
    lde = orc.coset_lde(oracle, m, added)
    for c in range(2):
        assert [int(x) for x in lde[:, c]] == nv.coset_lde_bitrev([int(x) for x in m[:, c]], added)
    coeffs = nv.idft([int(x) for x in m[:, 0]])
    for log_q in range(0, added + 1):
        nq = n << log_q
        lq = log_n + log_q
        wq = nv.two_adic_generator(lq)
        for i in range(nq):
            assert int(lde[nv.rev(i, lq), 0]) == nv.poly_eval(coeffs, 7 * pow(wq, i, P) % P)


def test_lde_from_coefficients_matches_commit_transform(oracle):
    """Reference pin src/prover.rs:975-999 (same sizes; numpy PCG instead of SmallRng)."""
    rng = np.random.default_rng(0)
    for log_h in [0, 1, 2, 5, 8]:
        for log_blowup in [1, 2, 3]:
            for width in [1, 2, 7]:
                h = 1 << log_h
                coeffs = orc.rand_matrix(rng, h, width)
                evals = np.empty_like(coeffs)
                oracle.orc_coset_dft_batch(coeffs, h, width, 1, evals)
                expected = orc.coset_lde(oracle, evals, log_blowup, 7)
                # lde_from_coefficients: scale rows by GENERATOR^j, zero-pad, one DFT
                scaled = np.array([[int(coeffs[j, c]) * pow(7, j, P) % P for c in range(width)] for j in range(h)],
                                  dtype=np.uint64)
                got = np.empty((h << log_blowup, width), dtype=np.uint64)
                oracle.orc_lde_from_shifted_coefficients(scaled, h, width, log_blowup, got)
                assert (got == expected).all(), (log_h, log_blowup, width)


def test_shifted_quotient_slices_matches_naive_composition(oracle):
    """Reference pin src/prover.rs:1006-1041."""
    rng = np.random.default_rng(1)
    for log_n in [0, 1, 2, 5, 7]:
        for q in [1, 2, 4]:
            for d in [1, 2]:
                n = 1 << log_n
                big = n * q
                evals = orc.rand_matrix(rng, big, d)
                coeffs = np.empty_like(evals)
                oracle.orc_coset_idft_batch(evals, big, d, 7, coeffs)
                width = q * d
                expected = np.zeros((n, width), dtype=np.uint64)
                for row in range(n):
                    s = pow(7, row, P)
                    for k in range(q):
                        for c in range(d):
                            expected[row, k * d + c] = int(coeffs[k * n + row, c]) * s % P
                got = np.empty((n, width), dtype=np.uint64)
                oracle.orc_shifted_quotient_slices(evals, big, d, q, got)
                assert (got == expected).all(), (log_n, q, d)


def test_large_dft_roundtrip(oracle):
    """Size-independent property at a size that exercises the chunked/threaded path."""
    rng = np.random.default_rng(7)
    n, w = 1 << 14, 5
    m = orc.rand_matrix(rng, n, w)
    f = np.empty_like(m)
    oracle.orc_dft_batch(m, n, w, f)
    back = np.empty_like(m)
    oracle.orc_idft_batch(f, n, w, back)
    assert (back == m).all()
    # Parseval-style spot check: dft_0 = sum of column
    assert int(f[0, 0]) == sum(int(x) for x in m[:, 0]) % P
