"""GPU parity of the DFT / coset-LDE slot (libmsgpu through the C ABI) against the CPU oracle.
Bit-exact: integer field arithmetic, canonical outputs. Mirrors the reference's pinning test
`lde_from_coefficients_matches_commit_transform` (src/prover.rs:975-999)."""
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


def edge_matrix(rng, rows, cols):
    m = orc.rand_matrix(rng, rows, cols)
    flat = m.reshape(-1)
    if flat.size >= 4:  # field edge values
        flat[0] = P - 1
        flat[1] = 0
        flat[2] = 1
        flat[3] = P - 2**32
    return m


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 7, 8, 10, 11, 12, 13, 15])
@pytest.mark.parametrize("w", [1, 2, 7, 14, 26])
def test_dft_bitrev_matches_oracle(gpu, oracle, log_n, w):
    ms, ctx = gpu
    rng = np.random.default_rng(1000 * log_n + w)
    m = edge_matrix(rng, 1 << log_n, w)
    got = ms.GpuDft(ctx).dft_batch_bitrev(m)
    assert np.array_equal(got, orc.dft_bitrev(oracle, m))


@pytest.mark.parametrize("log_n,w", [(0, 3), (3, 5), (9, 4), (12, 3), (16, 2)])
def test_dft_natural_and_idft(gpu, oracle, log_n, w):
    ms, ctx = gpu
    rng = np.random.default_rng(77 + log_n)
    m = edge_matrix(rng, 1 << log_n, w)
    dft = ms.GpuDft(ctx)
    want = np.empty_like(m)
    oracle.orc_dft_batch(m, m.shape[0], m.shape[1], want)
    got = dft.dft_batch(m)
    assert np.array_equal(got, want)
    assert np.array_equal(dft.idft_batch(got), m)


@pytest.mark.parametrize("log_h", [0, 1, 2, 5, 8, 10, 11, 13])
@pytest.mark.parametrize("lb", [1, 2, 3])
@pytest.mark.parametrize("w", [1, 2, 7])
def test_coset_lde_matches_oracle(gpu, oracle, log_h, lb, w):
    """The sizes of the reference's pinning test (h in 2^{0,1,2,5,8}, lb in {1,2,3}, w in {1,2,7}) + larger."""
    ms, ctx = gpu
    rng = np.random.default_rng(0 + 100 * log_h + 10 * lb + w)
    m = edge_matrix(rng, 1 << log_h, w)
    got = ms.GpuDft(ctx).coset_lde_batch_bitrev(m, lb, 7)
    assert np.array_equal(got, orc.coset_lde(oracle, m, lb, 7))


@pytest.mark.parametrize("log_h,lb,w", [(16, 1, 14), (14, 2, 26), (17, 1, 1), (12, 1, 40), (20, 1, 2), (21, 1, 1)])
def test_coset_lde_large(gpu, oracle, log_h, lb, w):
    ms, ctx = gpu
    rng = np.random.default_rng(5)
    m = edge_matrix(rng, 1 << log_h, w)
    got = ms.GpuDft(ctx).coset_lde_batch_bitrev(m, lb, 7)
    assert np.array_equal(got, orc.coset_lde(oracle, m, lb, 7))


def test_coset_lde_other_shift(gpu, oracle):
    ms, ctx = gpu
    rng = np.random.default_rng(6)
    m = edge_matrix(rng, 256, 3)
    for shift in (1, 49, P - 5):
        assert np.array_equal(ms.GpuDft(ctx).coset_lde_batch_bitrev(m, 2, shift), orc.coset_lde(oracle, m, 2, shift))


@pytest.mark.parametrize("log_h", [0, 1, 2, 5, 8, 12])
@pytest.mark.parametrize("lb", [1, 2, 3])
@pytest.mark.parametrize("w", [1, 2, 7])
def test_lde_from_shifted_coefficients(gpu, oracle, log_h, lb, w):
    """src/prover.rs:709-717 / :975-999."""
    ms, ctx = gpu
    rng = np.random.default_rng(0)
    m = edge_matrix(rng, 1 << log_h, w)
    want = np.empty((m.shape[0] << lb, w), dtype=np.uint64)
    oracle.orc_lde_from_shifted_coefficients(m, m.shape[0], w, lb, want)
    assert np.array_equal(ms.GpuDft(ctx).lde_from_shifted_coefficients(m, lb), want)


def test_lde_linearity_full_size(gpu):
    """Size-independent property at the bench size (2^20 x 14, blowup 2): LDE(a) + LDE(b) == LDE(a + b),
    and the first-coset rows restrict to ... the LDE of a constant column is constant."""
    ms, ctx = gpu
    rng = np.random.default_rng(9)
    n, w = 1 << 20, 14
    a = orc.rand_matrix(rng, n, w)
    b = orc.rand_matrix(rng, n, w)
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    dft = ms.GpuDft(ctx)
    la, lb_, ls = (dft.coset_lde_batch_bitrev(x, 1, 7) for x in (a, b, s))
    tot = la.astype(object) + lb_.astype(object)
    assert np.array_equal((tot % P).astype(np.uint64), ls)
    c = np.full((n, 1), 12345, dtype=np.uint64)
    assert np.all(dft.coset_lde_batch_bitrev(c, 1, 7) == 12345)


def test_rejects_bad_shapes(gpu):
    ms, ctx = gpu
    with pytest.raises(ms.MsgpuError):
        ms.GpuDft(ctx).dft_batch_bitrev(np.zeros((3, 2), dtype=np.uint64))
    with pytest.raises(ms.MsgpuError):
        ms.GpuDft(ctx).coset_lde_batch_bitrev(np.zeros((4, 2), dtype=np.uint64), 1, 0)
