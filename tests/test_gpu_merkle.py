"""GPU parity of the BLAKE3 Merkle MMCS and of Pcs::commit against the CPU oracle and the official
`blake3` package. Cases follow the reference's `gen_pcs_refs` inputs (src/types.rs:246-282) and its
BLAKE3 known-answer vector (src/test_circuits/blake3.rs:2646-2746)."""
import ctypes as C
import struct

import blake3 as pyb3
import numpy as np
import pytest

from tests import _oracle as orc
from tests.test_oracle_hash import KAT_MSG, KAT_OUT, KAT_STATE, leaf, b3

pytestmark = pytest.mark.gpu
P = orc.P


@pytest.fixture(scope="module")
def gpu():
    import multi_stark_b200 as ms
    ctx = ms.GpuContext(0)
    yield ms, ctx
    ctx.close()


def test_compression_known_answer(gpu):
    ms, ctx = gpu
    st = np.array(KAT_STATE, dtype=np.uint32)
    msg = np.array(KAT_MSG, dtype=np.uint32)
    out = np.zeros(16, dtype=np.uint32)
    from multi_stark_b200._ffi import check
    check(ctx.L.msgpu_blake3_compress_raw(ctx.h, st.ctypes.data_as(C.c_void_p), msg.ctypes.data_as(C.c_void_p),
                                          out.ctypes.data_as(C.c_void_p)))
    assert [int(x) for x in out] == KAT_OUT


@pytest.mark.parametrize("n", [1, 3, 7, 8, 9, 16, 17, 20, 22, 127, 128, 129, 130, 256, 257, 384, 385, 1000, 2625, 7000])
def test_leaf_hash_widths(gpu, n):
    """Row of n elements 1..=n (gen_pcs_refs uses n = 3, 17, 22, 20); wide rows cross BLAKE3 chunk
    boundaries (128 columns = 1024 bytes); 2625 is the widest circuit of the reference's tests; 7000
    exercises the unstaged fallback."""
    ms, ctx = gpu
    row = np.arange(1, n + 1, dtype=np.uint64).reshape(1, n)
    root, pd = ms.GpuMmcs(ctx).commit([row])
    assert bytes(root) == leaf(range(1, n + 1))
    pd.free()


def test_gen_pcs_refs_merkle(gpu):
    ms, ctx = gpu
    m0 = np.zeros((8, 2), dtype=np.uint64); m0[5] = [11, 12]
    m1 = np.zeros((4, 3), dtype=np.uint64); m1[2] = [107, 108, 109]
    m2 = np.zeros((2, 1), dtype=np.uint64); m2[1] = [202]
    from tests.test_oracle_hash import py_mmcs_commit
    layers = py_mmcs_commit([m0, m1, m2])
    root, pd = ms.GpuMmcs(ctx).commit([m0, m1, m2])
    assert bytes(root) == layers[-1][0]
    got = pd.layers()
    assert len(got) == 4
    for g, e in zip(got, layers):
        assert [bytes(x) for x in g] == e
    opened, proofs = pd.open_batch([5])
    assert [int(x) for x in opened[0]] == [11, 12, 107, 108, 109, 202]
    assert [bytes(p) for p in proofs[0]] == [layers[0][4], layers[1][3], layers[2][0]]


@pytest.mark.parametrize("shapes", [[(16, 1)], [(16, 3), (16, 2)], [(4, 5), (32, 1), (32, 9), (8, 2)], [(2, 140)],
                                    [(64, 14), (2, 1)], [(8, 300), (8, 1), (4, 129)], [(1, 5)], [(1, 1), (1, 2)],
                                    [(4096, 14), (256, 1)], [(1 << 14, 26), (1 << 14, 2), (1 << 9, 2)],
                                    [(512, 256)], [(256, 2625)],
                                    # the streamed wide-row kernel: segment / chunk boundaries, matrices straddling segments,
                                    # ragged last block, partial last CTA
                                    [(64, 129)], [(256, 200), (256, 57)], [(128, 300), (128, 3), (128, 64), (32, 131)],
                                    [(32, 513)], [(1024, 97), (1024, 33), (64, 1000)]])
def test_mmcs_random_vs_oracle(gpu, oracle, shapes):
    ms, ctx = gpu
    rng = np.random.default_rng(len(shapes) * 31 + shapes[0][0])
    mats = [orc.rand_matrix(rng, h, w) for h, w in shapes]
    t = orc.MmcsTree(oracle, mats)
    root, pd = ms.GpuMmcs(ctx).commit(mats)
    assert bytes(root) == bytes(t.root)
    for g, e in zip(pd.layers(), t.layers()):
        assert np.array_equal(g, e)
    max_h = max(h for h, _ in shapes)
    idx = sorted({0, 1 % max_h, max_h // 2, max_h - 1, (max_h * 3) // 7})
    opened, proofs = pd.open_batch(idx)
    for k, i in enumerate(idx):
        eo, ep = t.open(i)
        assert np.array_equal(opened[k], eo)
        assert np.array_equal(proofs[k], ep)
        assert t.verify(i, opened[k], proofs[k])
    pd.free()


@pytest.mark.parametrize("shapes,lb", [([(256, 1), (4096, 14)], 1), ([(256, 2), (4096, 26)], 1), ([(64, 3)], 2),
                                       ([(1, 2), (2, 2), (1 << 13, 5)], 3), ([(1 << 16, 14)], 1)])
def test_pcs_commit_matches_oracle(gpu, oracle, shapes, lb):
    """Pcs::commit = coset LDE (shift 7, bit-reversed) + MMCS; first case = BASELINE config 1 stage-1 shapes."""
    ms, ctx = gpu
    rng = np.random.default_rng(11)
    mats = [orc.rand_matrix(rng, h, w) for h, w in shapes]
    ldes = [orc.coset_lde(oracle, m, lb, 7) for m in mats]
    t = orc.MmcsTree(oracle, ldes)
    pcs = ms.GpuPcs(ctx, lb)
    root, pd = pcs.commit(mats)
    assert bytes(root) == bytes(t.root)
    for i, l in enumerate(ldes):
        assert np.array_equal(pd.read_rows(i), l)
    # device-input variant and commit_ldes on the device-resident LDEs give the same root
    dptrs = [(ctx.upload(m), m.shape[0], m.shape[1]) for m in mats]
    root2, pd2 = pcs.commit_dev(dptrs)
    assert bytes(root2) == bytes(root)
    views = [pd.matrix_info(i) for i in range(pd.num_matrices)]
    root3, pd3 = pcs.commit_ldes(views)
    assert bytes(root3) == bytes(root)
    # get_evaluations_on_domain: first n*q stored rows, bit-reversed view = coset 7*H_{nq} in natural order
    ptr, nq, cols = pcs.get_evaluations_on_domain(pd, len(mats) - 1, (shapes[-1][0]).bit_length() - 1)
    assert (nq, cols) == shapes[-1]
    for p, _, _ in dptrs:
        ctx.free(p)
    pd3.free(); pd2.free(); pd.free()


def _split_columns(m, widths):
    out, c = [], 0
    for w in widths:
        out.append(np.ascontiguousarray(m[:, c:c + w]))
        c += w
    assert c == m.shape[1]
    return out


@pytest.mark.parametrize("shapes", [
    # (rows, block widths or None): matrices of one commitment, commit order
    [(4096, [7, 7])],                                         # one matrix from two blocks: staged leaf kernel
    [(256, None), (4096, [2, 2, 2, 2, 2, 2, 1, 1])],          # BASELINE cfg1 stage-1 shapes split over 8 ranks + a whole byte table
    [(2048, [4, 4, 3, 3, 3, 3, 3, 3]), (2048, [1, 1, 0, 0])],  # two assembled matrices in one class, zero-width blocks
    [(1024, [32, 32, 32, 32, 32, 32, 32, 32])],               # 256 columns: the streamed leaf kernel
    [(512, [100, 100, 56]), (512, None), (1024, [3, 2])],     # streamed, mixed with a plain matrix and a taller class
    [(64, [1, 1]), (64, [1, 1]), (64, [2, 1]), (64, [1, 2]), (64, [3, 3])],  # five assembled matrices: one falls back to its own gather
    [(8, [6000, 3000])],                                      # rows too wide to stage: gather pass + direct kernel
])
def test_commit_from_column_blocks(gpu, shapes):
    """msgpu_commit_ldes_blocks_dev: matrices that still exist as column blocks (the row shard of a column-sharded LDE, read from
    the peers' windows in the multi-GPU prover) are assembled by the leaf-hash pass. Same root, same digest layers and the same
    row-major matrices as committing the assembled matrices."""
    ms, ctx = gpu
    rng = np.random.default_rng(5)
    pcs = ms.GpuPcs(ctx, 1)
    whole, args, frees, dsts = [], [], [], []
    for rows, widths in shapes:
        w = sum(widths) if widths else 5
        m = orc.rand_matrix(rng, rows, w)
        whole.append(m)
        if widths is None:
            p = ctx.upload(m)
            frees.append(p)
            args.append((p, rows, w, None))
            dsts.append(None)
            continue
        blocks = []
        for b in _split_columns(m, widths):
            bp = ctx.upload(b) if b.shape[1] else ctx.malloc(8)
            frees.append(bp)
            blocks.append((bp, b.shape[1]))
        dst = ctx.malloc(rows * w * 8)
        frees.append(dst)
        dsts.append(dst)
        args.append((dst, rows, w, blocks))
    root, pd = pcs.commit_ldes_blocks(args)
    ref_ptrs = [(ctx.upload(m), m.shape[0], m.shape[1]) for m in whole]
    root_ref, pd_ref = pcs.commit_ldes(ref_ptrs)
    assert bytes(root) == bytes(root_ref)
    for i, m in enumerate(whole):
        assert np.array_equal(pd.read_rows(i), m), "matrix %d was not assembled correctly" % i
    for a, b in zip(pd.layers(), pd_ref.layers()):
        assert np.array_equal(a, b)
    max_h = max(m.shape[0] for m in whole)
    idx = [0, max_h - 1, 3 % max_h]
    got, want = pd.open_batch(idx), pd_ref.open_batch(idx)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    pd.free(); pd_ref.free()
    for p, _, _ in ref_ptrs:
        ctx.free(p)
    for p in frees:
        ctx.free(p)


def test_commit_full_size_root_of_roots(gpu):
    """Bench-size commit (2^20 x 14, blowup 2): leaf digests recomputed on the host with the official
    blake3 package for a sample of rows, and every sampled opening verifies against the root."""
    ms, ctx = gpu
    rng = np.random.default_rng(3)
    m = orc.rand_matrix(rng, 1 << 20, 14)
    pcs = ms.GpuPcs(ctx, 1)
    root, pd = pcs.commit([m])
    idx = [0, 1, 12345, (1 << 21) - 1, 1 << 20, 77777]
    opened, proofs = pd.open_batch(idx)
    for k, i in enumerate(idx):
        d = leaf(opened[k])
        j = i
        for sib in proofs[k]:
            d = b3(d + bytes(sib)) if j % 2 == 0 else b3(bytes(sib) + d)
            j >>= 1
        assert d == bytes(root)
    pd.free()


@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 1023, 1024, 1025, 2047, 2048, 2049, 3072, 4097, 5 * 1024, 7 * 1024 + 3,
                               (1 << 16) + 1, (1 << 20) + 777, 3 * (1 << 20)])
def test_blake3_long_hash_matches_official(gpu, n):
    """msgpu_blake3_hash (the transcript's long flushes) against the official BLAKE3 implementation, across chunk
    and subtree boundaries (odd chunk counts exercise the left-heavy tree)."""
    ms, ctx = gpu
    rng = np.random.default_rng(n)
    data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
    assert ctx.blake3_hash(data) == pyb3.blake3(data).digest()


def test_upload_canonical_rejects_out_of_range(gpu):
    ms, ctx = gpu
    good = np.array([0, 1, ms.P - 1], dtype=np.uint64)
    ctx.free(ctx.upload_canonical(good))
    for bad in (ms.P, 2**64 - 1):
        a = np.zeros(5000, dtype=np.uint64)
        a[4321] = bad
        with pytest.raises(ms.MsgpuError):
            ctx.upload_canonical(a)


@pytest.mark.parametrize("log_n,w,lb", [(22, 14, 1), (22, 1, 2), (23, 4, 1), (24, 1, 1), (24, 4, 1), (24, 14, 1), (18, 256, 2)])
def test_commit_root_bit_exact_at_full_sizes(gpu, oracle, log_n, w, lb):
    """Bit-exact Pcs::commit roots against the CPU oracle at the sizes of BASELINE configs[4] (2^22 .. 2^24 rows: the three- and
    four-pass NTT plans 8/7/7, 8/8/7, 8/8/8 and the fused middle pass at tb = 7, 8) and of configs[2] (256 columns: two BLAKE3
    chunks per leaf). The root binds every LDE row, so root equality is LDE equality; a sample of rows is compared as well."""
    ms, ctx = gpu
    rng = np.random.default_rng(log_n * 100 + w)
    m = orc.rand_matrix(rng, 1 << log_n, w)
    m[0, :] = P - 1
    want_root, h = orc.pcs_commit(oracle, [m], lb)
    pcs = ms.GpuPcs(ctx, lb)
    root, pd = pcs.commit([m])
    assert bytes(root) == want_root
    rows = (1 << log_n) << lb
    for r0 in (0, rows // 2 - 3, rows - 8):
        got = pd.read_rows(0, r0, 8)
        # the oracle handle keeps its LDE: compare through its opening API
        for k in range(8):
            opened, _ = _orc_open(oracle, h, r0 + k, w, rows)
            assert np.array_equal(got[k], opened)
    oracle.orc_mmcs_free(h)
    pd.free()


def _orc_open(L, h, index, width, rows):
    opened = np.zeros(width, dtype=np.uint64)
    depth = rows.bit_length() - 1
    proof = np.zeros((max(depth, 1), 32), dtype=np.uint8)
    L.orc_mmcs_open(h, index, opened, proof)
    return opened, proof[:depth]
