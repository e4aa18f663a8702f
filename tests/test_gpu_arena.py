"""Device memory arena (Ctx::alloc/free behind msgpu_malloc / msgpu_free, csrc/capi.cu) and the sharded-commit entry points'
argument checks. The arena hands freed blocks out again at once (single-stream ordering), coalesces neighbours and never lets
a foreign or stale pointer corrupt its bookkeeping."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import multi_stark_b200 as ms
    ctx = ms.GpuContext(0)
    yield ms, ctx
    ctx.close()


def test_reuse_and_coalescing(gpu):
    ms, ctx = gpu
    a = ctx.malloc(1 << 20)
    b = ctx.malloc(1 << 20)
    c = ctx.malloc(1 << 20)
    assert len({a, b, c}) == 3 and all(p % 512 == 0 for p in (a, b, c))
    ctx.free(b)
    assert ctx.malloc(1 << 20) == b          # exact fit is taken again
    ctx.free(a)
    ctx.free(b)                                # a and b are neighbours: they merge
    ab = ctx.malloc(2 << 20)
    assert ab == a
    ctx.free(ab)
    ctx.free(c)


def test_data_survives_neighbouring_traffic(gpu):
    ms, ctx = gpu
    rng = np.random.default_rng(3)
    keep = rng.integers(0, 2**63, size=1 << 16, dtype=np.uint64)
    p = ctx.upload(keep)
    for k in range(20):
        q = [ctx.malloc(int(s)) for s in rng.integers(1, 1 << 22, size=8)]
        for x in q[::2]:
            ctx.free(x)
        junk = ctx.upload(rng.integers(0, 2**63, size=1 << 12, dtype=np.uint64))
        ctx.free(junk)
        for x in q[1::2]:
            ctx.free(x)
    assert np.array_equal(ctx.download(p, keep.shape), keep)
    ctx.free(p)


def test_double_free_and_foreign_pointer_are_errors(gpu):
    ms, ctx = gpu
    p = ctx.malloc(4096)
    ctx.free(p)
    with pytest.raises(ms.MsgpuError):
        ctx.free(p)
    with pytest.raises(ms.MsgpuError):
        ctx.free(p + 512)
    assert ctx.malloc(4096)  # the arena still works


def test_tree_from_digests_rejects_bad_classes(gpu):
    ms, ctx = gpu
    L = ctx.L
    d = ctx.malloc(64 * 32)
    out, root = C.c_void_p(), np.zeros(32, dtype=np.uint8)
    hs = (C.c_uint64 * 2)(64, 64)          # two classes of one height cannot both be injected
    ps = (C.c_void_p * 2)(d, d)
    assert L.msgpu_tree_from_digests(ctx.h, 2, hs, ps, C.byref(out), root.ctypes.data_as(C.c_void_p)) != 0
    hs = (C.c_uint64 * 1)(48)              # not a power of two
    assert L.msgpu_tree_from_digests(ctx.h, 1, hs, ps, C.byref(out), root.ctypes.data_as(C.c_void_p)) != 0
    assert L.msgpu_tree_from_digests(ctx.h, 0, hs, ps, C.byref(out), root.ctypes.data_as(C.c_void_p)) != 0
    ctx.free(d)


def test_local_commit_plus_tree_equals_commit(gpu):
    """msgpu_commit_local_dev + msgpu_tree_from_digests on ONE rank must reproduce msgpu_commit_dev exactly."""
    ms, ctx = gpu
    L = ctx.L
    rng = np.random.default_rng(11)
    shapes = [(256, 3), (1024, 5), (256, 2), (64, 9)]
    mats = [rng.integers(0, ms.P, size=s, dtype=np.uint64) for s in shapes]
    dev = [ctx.upload(m) for m in mats]
    n = len(mats)
    ptrs = (C.c_void_p * n)(*dev)
    hs = (C.c_uint64 * n)(*[s[0] for s in shapes])
    ws = (C.c_uint64 * n)(*[s[1] for s in shapes])
    root, pd = ms.GpuPcs(ctx, 1).commit_dev([(d, s[0], s[1]) for d, s in zip(dev, shapes)])
    loc = C.c_void_p()
    assert L.msgpu_commit_local_dev(ctx.h, ptrs, hs, ws, n, 1, 0, C.byref(loc)) == 0
    k = int(L.msgpu_pdata_num_classes(loc))
    assert k == 3
    ch, cp = (C.c_uint64 * k)(), (C.c_void_p * k)()
    for i in range(k):
        h, p = C.c_uint64(), C.c_void_p()
        assert L.msgpu_pdata_class_digests(loc, i, C.byref(h), C.byref(p)) == 0
        ch[i], cp[i] = h.value, p.value
    assert list(ch) == [2048, 512, 128]
    tree, r2 = C.c_void_p(), np.zeros(32, dtype=np.uint8)
    assert L.msgpu_tree_from_digests(ctx.h, k, ch, cp, C.byref(tree), r2.ctypes.data_as(C.c_void_p)) == 0
    assert bytes(r2) == bytes(root)
    assert int(L.msgpu_pdata_max_height(tree)) == 2048
    # rows from the local part, paths from the tree part == open_batch of the ordinary commitment
    idx = np.array([0, 5, 2047, 1024], dtype=np.uint64)
    want_rows, want_paths = pd.open_batch(idx)
    trees = (C.c_void_p * 2)(loc, tree)
    shifts = (C.c_uint32 * 2)(0, 0)
    rows = np.zeros((2, 0), dtype=np.uint64)
    opened = np.zeros(len(idx) * 19, dtype=np.uint64)
    proofs = np.zeros(len(idx) * 11 * 32, dtype=np.uint8)
    assert L.msgpu_open_batch_multi(ctx.h, trees, shifts, 2, idx.ctypes.data_as(C.c_void_p), len(idx),
                                    opened.ctypes.data_as(C.c_void_p), proofs.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(opened.reshape(len(idx), 19), want_rows)
    assert np.array_equal(proofs.reshape(len(idx), 11, 32), want_paths)
    L.msgpu_pdata_free(loc)
    L.msgpu_pdata_free(tree)
    pd.free()
    for d in dev:
        ctx.free(d)


def _upload_begin(ms, ctx, mats):
    n = len(mats)
    ptrs = (C.c_void_p * n)(*[m.ctypes.data for m in mats])
    hs = (C.c_uint64 * n)(*[m.shape[0] for m in mats])
    ws = (C.c_uint64 * n)(*[m.shape[1] for m in mats])
    up = C.c_void_p()
    rc = ctx.L.msgpu_upload_begin(ctx.h, ptrs, hs, ws, n, C.byref(up))
    return rc, up


def test_pipelined_commit_equals_commit(gpu):
    """msgpu_upload_begin + msgpu_commit_upload (uploads on the copy stream) == msgpu_commit, and the kept device inputs are the
    natural-order matrices"""
    ms, ctx = gpu
    rng = np.random.default_rng(21)
    mats = [ctx.pinned_copy(rng.integers(0, ms.P, size=s, dtype=np.uint64)) for s in [(256, 1), (4096, 14), (1024, 3), (4096, 2)]]
    want_root, want_pd = ms.GpuPcs(ctx, 1).commit(mats)
    rc, up = _upload_begin(ms, ctx, mats)
    assert rc == 0
    kept = (C.c_void_p * len(mats))()
    pd, root = C.c_void_p(), np.zeros(32, dtype=np.uint8)
    assert ctx.L.msgpu_commit_upload(up, 1, 1, kept, 0, C.byref(pd), root.ctypes.data_as(C.c_void_p)) == 0
    assert bytes(root) == bytes(want_root)
    for k, m in enumerate(mats):
        assert np.array_equal(ctx.download(kept[k], m.shape), m)
        ctx.free(kept[k])
    ctx.L.msgpu_pdata_free(pd)
    want_pd.free()


def test_pipelined_commit_rejects_non_canonical_values(gpu):
    ms, ctx = gpu
    bad = np.zeros((64, 2), dtype=np.uint64)
    bad[17, 1] = ms.P            # = 0 mod p, but not the canonical representative
    rc, up = _upload_begin(ms, ctx, [np.ones((64, 3), dtype=np.uint64), bad])
    assert rc == 0
    pd, root = C.c_void_p(), np.zeros(32, dtype=np.uint8)
    assert ctx.L.msgpu_commit_upload(up, 1, 1, None, 0, C.byref(pd), root.ctypes.data_as(C.c_void_p)) != 0
    assert b"canonical" in ctx.L.msgpu_last_error()
    # the handle was consumed and its buffers released: the arena serves the same sizes again
    rc, up = _upload_begin(ms, ctx, [np.ones((64, 3), dtype=np.uint64)])
    assert rc == 0
    ctx.L.msgpu_upload_free(up)   # abandoning an upload is allowed


def test_upload_begin_argument_checks(gpu):
    ms, ctx = gpu
    up = C.c_void_p()
    assert ctx.L.msgpu_upload_begin(ctx.h, None, None, None, 0, C.byref(up)) != 0
    m = np.ones((48, 2), dtype=np.uint64)  # height not a power of two
    rc, up = _upload_begin(ms, ctx, [m])
    assert rc != 0
