"""Pins the oracle's BLAKE3 and MMCS layer (oracle/cpu_mmcs.hpp, host/blake3_host.hpp).

Absolute pins: the reference's two known-answer vectors (src/test_circuits/blake3.rs:2615-2746) and
the official `blake3` Python package (same algorithm as the `blake3` crate the reference links,
Cargo.lock:71-72). The `gen_pcs_refs` inputs (src/types.rs:246-282) are reproduced with expected
values computed independently here in Python."""
import struct

import blake3 as pyb3
import numpy as np
import pytest

from tests import _oracle as orc

P = orc.P


def b3(data: bytes) -> bytes:
    return pyb3.blake3(data).digest()


def limbs(d: bytes):
    return list(struct.unpack("<4Q", d))


def leaf(vals) -> bytes:
    return b3(b"".join(struct.pack("<Q", int(v)) for v in vals))


def test_g_function_test_vector():
    """src/test_circuits/blake3.rs:2615-2644 (pure arithmetic restatement)."""
    M = 0xFFFFFFFF
    rotr = lambda x, n: ((x >> n) | (x << (32 - n))) & M
    a, b, c, d, mx, my = 0x11111111, 0x22222222, 0x33333333, 0x44444444, 0x55555555, 0x66666666
    a = (a + b + mx) & M; d = rotr(d ^ a, 16); c = (c + d) & M; b = rotr(b ^ c, 12)
    a = (a + b + my) & M; d = rotr(d ^ a, 8); c = (c + d) & M; b = rotr(b ^ c, 7)
    assert (a, b, c, d) == (0xCCCCCCCB, 0x45B64444, 0x06FFFFFF, 0x07000000)


KAT_STATE = [i * 0x1111 for i in range(16)]
KAT_MSG = [i * 0x11110000 for i in range(16)]
KAT_OUT = [0xD304E51C, 0xC2DF34A0, 0x5EBA7F1F, 0x2AB9650F, 0xD9CEF159, 0x4E9D3A6A, 0xCAC2E310, 0xC6B9BE7E,
           0xAD9FD58A, 0x0899E71B, 0xCA51A599, 0xC3FBD7C0, 0x751D2F26, 0x6CD0AC6B, 0xC58F3C1D, 0xE6D65414]


def test_compression_test_vector(oracle):
    """src/test_circuits/blake3.rs:2646-2746: the only absolute golden value in the reference."""
    out = np.zeros(16, dtype=np.uint32)
    oracle.orc_blake3_compress_raw(np.array(KAT_STATE, dtype=np.uint32), np.array(KAT_MSG, dtype=np.uint32), out)
    assert [int(x) for x in out] == KAT_OUT


@pytest.mark.parametrize("n", [0, 1, 3, 63, 64, 65, 127, 128, 129, 1023, 1024, 1025, 2047, 2048, 2049, 3072, 3073,
                               4096, 5000, 8192, 8193, 16384 + 7, 21000])
def test_blake3_vs_reference_package(oracle, n):
    rng = np.random.default_rng(n)
    data = rng.integers(0, 256, size=max(n, 1), dtype=np.uint8)[:n].copy()
    out = np.zeros(32, dtype=np.uint8)
    oracle.orc_blake3(data if n else np.zeros(1, dtype=np.uint8), n, out)
    assert bytes(out) == b3(bytes(data))


def test_gen_pcs_refs_leaf_and_compress(oracle):
    """Inputs of gen_pcs_refs (src/types.rs:246-258); LEAF3/COMPRESS values quoted in SURVEY 8(c)."""
    for n in (3, 17, 22, 20):
        row = np.arange(1, n + 1, dtype=np.uint64)
        out = np.zeros(32, dtype=np.uint8)
        oracle.orc_hash_row(row, n, out)
        assert bytes(out) == leaf(range(1, n + 1))
        if n == 3:
            assert limbs(bytes(out)) == [4163513704854067712, 9384471110237386207, 13671380075168847140,
                                         1533933974187331481]
    dig = lambda xs: struct.pack("<4Q", *xs)
    l, r = dig([1, 2, 3, 4]), dig([5, 6, 7, 8])
    out = np.zeros(32, dtype=np.uint8)
    oracle.orc_compress(np.frombuffer(l, dtype=np.uint8).copy(), np.frombuffer(r, dtype=np.uint8).copy(), out)
    assert bytes(out) == b3(l + r)
    assert limbs(bytes(out)) == [16432952784711837466, 12565756115161032165, 6915939387221618258,
                                 11123773279136987111]


def py_mmcs_commit(mats):
    """Independent Python restatement of MerkleTreeMmcs::commit."""
    order = sorted(range(len(mats)), key=lambda i: -mats[i].shape[0])  # stable
    pos = 0
    max_h = mats[order[0]].shape[0]
    group = []
    while pos < len(order) and mats[order[pos]].shape[0] == max_h:
        group.append(mats[order[pos]]); pos += 1
    layers = [[leaf([v for m in group for v in m[i]]) for i in range(max_h)]]
    while len(layers[-1]) > 1:
        prev = layers[-1]
        nl = len(prev) // 2
        group = []
        while pos < len(order) and mats[order[pos]].shape[0] == nl:
            group.append(mats[order[pos]]); pos += 1
        nxt = []
        for i in range(nl):
            d = b3(prev[2 * i] + prev[2 * i + 1])
            if group:
                d = b3(d + leaf([v for m in group for v in m[i]]))
            nxt.append(d)
        layers.append(nxt)
    return layers


def test_gen_pcs_refs_merkle(oracle):
    """The 3-matrix mixed-height case of gen_pcs_refs (src/types.rs:260-281): heights 8/4/2,
    widths 2/3/1, opened at index 5."""
    m0 = np.zeros((8, 2), dtype=np.uint64); m0[5] = [11, 12]
    m1 = np.zeros((4, 3), dtype=np.uint64); m1[2] = [107, 108, 109]
    m2 = np.zeros((2, 1), dtype=np.uint64); m2[1] = [202]
    t = orc.MmcsTree(oracle, [m0, m1, m2])
    layers = py_mmcs_commit([m0, m1, m2])
    assert bytes(t.root) == layers[-1][0]
    got_layers = t.layers()
    assert len(got_layers) == 4
    for gl, el in zip(got_layers, layers):
        assert [bytes(x) for x in gl] == el
    opened, proof = t.open(5)
    assert [int(x) for x in opened] == [11, 12, 107, 108, 109, 202]
    assert [bytes(p) for p in proof] == [layers[0][4], layers[1][3], layers[2][0]]
    assert t.verify(5, opened, proof)
    bad = opened.copy(); bad[3] += 1
    assert not t.verify(5, bad, proof)
    assert not t.verify(4, opened, proof)


@pytest.mark.parametrize("shapes", [[(16, 1)], [(16, 3), (16, 2)], [(4, 5), (32, 1), (32, 9), (8, 2)],
                                    [(2, 140)], [(64, 14), (2, 1)], [(8, 300), (8, 1), (4, 129)]])
def test_mmcs_random(oracle, shapes):
    """Mixed heights, same-height concatenation, injection order = input order, rows > 1024 bytes
    (multi-chunk BLAKE3 leaves)."""
    rng = np.random.default_rng(len(shapes) * 31 + shapes[0][0])
    mats = [orc.rand_matrix(rng, h, w) for h, w in shapes]
    t = orc.MmcsTree(oracle, mats)
    layers = py_mmcs_commit(mats)
    assert bytes(t.root) == layers[-1][0]
    max_h = max(h for h, _ in shapes)
    for index in {0, 1, max_h // 2, max_h - 1}:
        opened, proof = t.open(index)
        exp = []
        for m in mats:
            exp += [int(v) for v in m[index >> ((max_h.bit_length() - 1) - (m.shape[0].bit_length() - 1))]]
        assert [int(x) for x in opened] == exp
        assert t.verify(index, opened, proof)
        if len(proof):
            p2 = proof.copy(); p2[0, 0] ^= 1
            assert not t.verify(index, opened, p2)
