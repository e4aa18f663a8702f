"""Standalone PCS (examples/pcs_example.rs:28-122): commit -> observe -> sample zeta -> open at [zeta; 2] -> verify.
The device's opened values + FRI proof must equal the CPU oracle's byte for byte, and the restated verifier must accept."""
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


def example_matrix(log_n, w):
    """entry (i, j) = i * w + j (examples/pcs_example.rs:58-60)"""
    n = 1 << log_n
    return (np.arange(n * w, dtype=np.uint64) % np.uint64(P)).reshape(n, w)


def gpu_pcs_example(ms, ctx, mats, num_open=2, **params):
    pcs = ms.GpuPcs(ctx, params.get("log_blowup", 1))
    root, pd = pcs.commit(mats)
    ch = ms.Challenger(**params)
    ch.observe(bytes(root))
    zeta = ch.sample_algebra_element()
    data, ms5 = ms.pcs_open(ctx, [(pd, [[zeta] * num_open for _ in mats])], ch)
    pd.free()
    ch.close()
    return bytes(root), zeta, data


@pytest.mark.parametrize("log_n,w,lb,fpl", [(5, 1, 1, 0), (0, 3, 1, 0), (1, 2, 2, 0), (10, 16, 1, 0), (12, 64, 2, 1), (9, 7, 3, 2),
                                            (14, 1, 1, 0), (13, 200, 1, 0)])
def test_pcs_example_matches_oracle(gpu, oracle, log_n, w, lb, fpl):
    ms, ctx = gpu
    params = dict(log_blowup=lb, log_final_poly_len=fpl, num_queries=25)
    m = example_matrix(log_n, w)
    if log_n <= fpl:
        pytest.skip("polynomial shorter than the final polynomial")
    root, zeta, data = gpu_pcs_example(ms, ctx, [m], **params)
    want_root, want_zeta, want = orc.pcs_example_prove(oracle, [m], log_blowup=lb, log_final_poly_len=fpl, num_queries=25)
    assert root == want_root and zeta == want_zeta
    assert data == want
    assert orc.pcs_example_verify(oracle, root, [m.shape], data, log_blowup=lb, log_final_poly_len=fpl, num_queries=25) == 1


def test_pcs_mixed_heights_and_random_values(gpu, oracle):
    ms, ctx = gpu
    rng = np.random.default_rng(5)
    mats = [orc.rand_matrix(rng, 1 << 9, 5), orc.rand_matrix(rng, 1 << 12, 3), orc.rand_matrix(rng, 1 << 9, 1),
            orc.rand_matrix(rng, 1 << 4, 2)]
    root, zeta, data = gpu_pcs_example(ms, ctx, mats, log_blowup=2, num_queries=30)
    want_root, want_zeta, want = orc.pcs_example_prove(oracle, mats, log_blowup=2, num_queries=30)
    assert (root, zeta) == (want_root, want_zeta)
    assert data == want
    shapes = [m.shape for m in mats]
    assert orc.pcs_example_verify(oracle, root, shapes, data, log_blowup=2, num_queries=30) == 1
    bad = bytearray(data)
    bad[40] ^= 2  # an opened value
    assert orc.pcs_example_verify(oracle, root, shapes, bytes(bad), log_blowup=2, num_queries=30) == 0


@pytest.mark.parametrize("log_n,w,lb", [(20, 4, 1), (22, 1, 2), (18, 64, 1)])
def test_pcs_large_sizes_verify(gpu, oracle, log_n, w, lb):
    """Sizes the oracle cannot prove in a unit test: the device opening is checked by the restated verifier."""
    ms, ctx = gpu
    rng = np.random.default_rng(log_n)
    m = orc.rand_matrix(rng, 1 << log_n, w)
    root, zeta, data = gpu_pcs_example(ms, ctx, [m], log_blowup=lb, num_queries=40)
    assert orc.pcs_example_verify(oracle, root, [m.shape], data, log_blowup=lb, num_queries=40) == 1
