"""The vectorised paths of the CPU oracle (oracle/cpu_simd.*: AVX-512 / AVX2 Goldilocks row kernels, 16-lane BLAKE3; and the
two-pass cache-blocked DFT of oracle/cpu_dft.hpp) against independent references: big-int sums for sampled DFT outputs, the
official `blake3` package for leaves and nodes, and the scalar clone of the same library (ORC_NO_AVX512 subprocess)."""
import os
import subprocess
import sys

import blake3
import numpy as np
import pytest

from tests import _naive as nv
from tests import _oracle as orc

P = orc.P


def test_simd_level_reported(oracle):
    assert oracle.orc_simd_level() in (1, 3, 4)


@pytest.mark.parametrize("log_n,w", [(14, 9), (15, 5), (13, 26)])
def test_two_pass_dft_matches_bigint_sums(oracle, log_n, w):
    """n * w * 8 bytes > 512 KB takes the gather / transform / twiddle / scatter split: spot-check outputs against the definition
    dft(f)_k = sum_j f_j w^{jk} (row rev(k) of the bit-reversed storage), all columns incl. the masked vector tail."""
    rng = np.random.default_rng(log_n)
    n = 1 << log_n
    assert n * w * 8 > (1 << 19)
    m = orc.rand_matrix(rng, n, w)
    m[0, :] = P - 1
    out = orc.dft_bitrev(oracle, m)
    g = nv.two_adic_generator(log_n)
    cols = [[int(v) for v in m[:, c]] for c in range(w)]
    for k in [0, 1, 2, n // 2, n - 1, 12345 % n, int(rng.integers(0, n))]:
        wk = pow(g, k, P)
        row = out[nv.rev(k, log_n)]
        for c in range(w):
            acc, x = 0, 1
            for v in cols[c]:
                acc += v * x
                x = x * wk % P
            assert acc % P == int(row[c]), (k, c)


@pytest.mark.parametrize("widths", [[1], [2], [14], [26], [1, 14], [2, 26], [7, 8, 9], [128], [129], [300]])
def test_leaf_hashing_lanes_match_official_blake3(oracle, widths):
    """rows of 8..1024 bytes go through the 16-lane compression (one row per lane, ragged last group); wider rows through the
    scalar multi-chunk path."""
    rng = np.random.default_rng(sum(widths))
    h = 64 if sum(widths) < 200 else 32
    mats = [orc.rand_matrix(rng, h, w) for w in widths]
    tree = orc.MmcsTree(oracle, mats)
    leaves = tree.layers()[0]
    for r in range(h):
        want = blake3.blake3(b"".join(m[r].tobytes() for m in mats)).digest()
        assert bytes(leaves[r]) == want
    nodes = tree.layers()[1]
    for i in range(h // 2):
        assert bytes(nodes[i]) == blake3.blake3(bytes(leaves[2 * i]) + bytes(leaves[2 * i + 1])).digest()


def test_injected_matrices_lanes(oracle):
    rng = np.random.default_rng(9)
    mats = [orc.rand_matrix(rng, 64, 3), orc.rand_matrix(rng, 32, 5), orc.rand_matrix(rng, 8, 2)]
    tree = orc.MmcsTree(oracle, mats)
    layers = tree.layers()
    prev = layers[0]
    want = [blake3.blake3(blake3.blake3(bytes(prev[2 * i]) + bytes(prev[2 * i + 1])).digest()
                          + blake3.blake3(mats[1][i].tobytes()).digest()).digest() for i in range(32)]
    assert [bytes(x) for x in layers[1]] == want


def test_vector_and_scalar_clones_agree():
    """The same commitment with the AVX-512 row kernels disabled (the auto-vectorised / scalar clones) gives the same root."""
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from tests import _oracle as orc; L = orc.lib(); "
            "rng = np.random.default_rng(3); m = [orc.rand_matrix(rng, 1 << 12, 14), orc.rand_matrix(rng, 256, 1)]; "
            "r, h = orc.pcs_commit(L, m, 2); print(r.hex())") % orc.ROOT
    a = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, ORC_NO_AVX512="1"))
    b = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=os.environ)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    assert a.stdout.strip() == b.stdout.strip() and len(a.stdout.strip()) == 64
