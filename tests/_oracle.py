"""ctypes loader for the CPU oracle (oracle/liboracle.so). Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    build()  # make is a no-op when liboracle.so is newer than its sources
    L = C.CDLL(path)
    sig = {
        "orc_num_threads": (C.c_int, []),
        "orc_simd_level": (C.c_int, []),
        "orc_set_num_threads": (None, [C.c_int]),
        "orc_fp_mul": (C.c_uint64, [C.c_uint64, C.c_uint64]),
        "orc_fp_inv": (C.c_uint64, [C.c_uint64]),
        "orc_two_adic_generator": (C.c_uint64, [C.c_uint32]),
        "orc_fp2_mul": (None, [u64p, u64p, u64p]),
        "orc_fp2_inv": (None, [u64p, u64p]),
        "orc_dft_batch": (None, [u64p, C.c_uint64, C.c_uint64, u64p]),
        "orc_dft_batch_bitrev": (None, [u64p, C.c_uint64, C.c_uint64, u64p]),
        "orc_idft_batch": (None, [u64p, C.c_uint64, C.c_uint64, u64p]),
        "orc_coset_dft_batch": (None, [u64p, C.c_uint64, C.c_uint64, C.c_uint64, u64p]),
        "orc_coset_idft_batch": (None, [u64p, C.c_uint64, C.c_uint64, C.c_uint64, u64p]),
        "orc_coset_lde_batch_bitrev": (None, [u64p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, u64p]),
        "orc_lde_from_shifted_coefficients": (None, [u64p, C.c_uint64, C.c_uint64, C.c_uint32, u64p]),
        "orc_shifted_quotient_slices": (None, [u64p, C.c_uint64, C.c_uint64, C.c_uint64, u64p]),
        "orc_blake3": (None, [u8p, C.c_uint64, u8p]),
        "orc_blake3_compress_raw": (None, [u32p, u32p, u32p]),
        "orc_hash_row": (None, [u64p, C.c_uint64, u8p]),
        "orc_compress": (None, [u8p, u8p, u8p]),
        "orc_mmcs_commit": (C.c_void_p, [C.POINTER(C.c_void_p), u64p, u64p, C.c_uint64, u8p]),
        "orc_mmcs_num_layers": (C.c_uint64, [C.c_void_p]),
        "orc_mmcs_layer_len": (C.c_uint64, [C.c_void_p, C.c_uint64]),
        "orc_mmcs_layer": (None, [C.c_void_p, C.c_uint64, u8p]),
        "orc_mmcs_open": (None, [C.c_void_p, C.c_uint64, u64p, u8p]),
        "orc_mmcs_verify": (C.c_int, [u8p, u64p, u64p, C.c_uint64, C.c_uint64, u64p, u8p, C.c_uint64]),
        "orc_mmcs_free": (None, [C.c_void_p]),
        "orc_system_create": (C.c_void_p, [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
        "orc_system_create_from_graphs": (C.c_void_p, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                       C.c_uint32, C.c_uint32, C.c_uint32]),
        "orc_system_free": (None, [C.c_void_p]),
        "orc_graph_num_nodes": (C.c_uint64, [C.c_void_p, C.c_uint32]),
        "orc_graph_nodes": (None, [C.c_void_p, C.c_uint32, u8p, u32p, u32p, u64p, u32p]),
        "orc_graph_num_zeros": (C.c_uint64, [C.c_void_p, C.c_uint32]),
        "orc_graph_zeros": (None, [C.c_void_p, C.c_uint32, u32p]),
        "orc_stage2_trace": (None, [C.c_void_p, C.c_uint32, u64p, C.c_uint64, u64p, u64p, u64p, u64p]),
        "orc_claims_accumulator": (None, [u64p, C.c_uint64, C.c_uint64, u64p, u64p, u64p]),
        "orc_quotient_values": (None, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, u64p, u64p, u64p, u64p, u64p]),
        "orc_selectors_on_coset": (None, [C.c_uint32, C.c_uint32, C.c_uint64, u64p, u64p, u64p, u64p]),
        "orc_pcs_commit": (C.c_void_p, [C.POINTER(C.c_void_p), u64p, u64p, C.c_uint64, C.c_uint32, u8p]),
        "orc_mmcs_matrix": (None, [C.c_void_p, C.c_uint64, u64p]),
        "orc_last_error": (C.c_char_p, []),
        "orc_prove": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), u64p, u64p, u64p, C.c_uint64, C.POINTER(C.c_void_p),
                                C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
        "orc_bytes_free": (None, [C.c_void_p]),
        "orc_last_transcript": (C.c_uint64, [u64p, C.c_uint64, u64p, C.c_uint64, u64p]),
        "orc_preprocessed_commit": (C.c_int, [C.c_void_p, u8p]),
        "orc_verify": (C.c_int, [C.c_void_p, u64p, u64p, C.c_uint64, u8p, C.c_uint64]),
        "orc_pcs_example_prove": (C.c_int, [C.POINTER(C.c_void_p), u64p, u64p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                            C.c_uint32, C.c_uint32, C.c_uint32, u8p, u64p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
        "orc_gen_challenger_refs": (None, [u64p]),
        "orc_check_lowering": (C.c_uint64, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, u32p]),
        "orc_pcs_example_verify": (C.c_int, [u8p, u64p, u64p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                             C.c_uint32, C.c_uint32, u8p, C.c_uint64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


P = 2**64 - 2**32 + 1


def rand_matrix(rng, rows, cols):
    return (rng.integers(0, P, size=(rows, cols), dtype=np.uint64)).astype(np.uint64)


def coset_lde(L, m, added_bits, shift=7):
    rows, cols = m.shape
    out = np.empty((rows << added_bits, cols), dtype=np.uint64)
    L.orc_coset_lde_batch_bitrev(np.ascontiguousarray(m), rows, cols, added_bits, shift, out)
    return out


def dft_bitrev(L, m):
    out = np.empty_like(m)
    L.orc_dft_batch_bitrev(np.ascontiguousarray(m), m.shape[0], m.shape[1], out)
    return out


class MmcsTree:
    def __init__(self, L, mats):
        self.L = L
        self.mats = [np.ascontiguousarray(m, dtype=np.uint64) for m in mats]
        n = len(mats)
        ptrs = (C.c_void_p * n)(*[m.ctypes.data for m in self.mats])
        self.heights = np.array([m.shape[0] for m in self.mats], dtype=np.uint64)
        self.widths = np.array([m.shape[1] for m in self.mats], dtype=np.uint64)
        self.root = np.zeros(32, dtype=np.uint8)
        self.h = L.orc_mmcs_commit(ptrs, self.heights, self.widths, n, self.root)
        assert self.h, "orc_mmcs_commit failed"

    def layers(self):
        out = []
        for i in range(self.L.orc_mmcs_num_layers(self.h)):
            ln = self.L.orc_mmcs_layer_len(self.h, i)
            buf = np.zeros((ln, 32), dtype=np.uint8)
            self.L.orc_mmcs_layer(self.h, i, buf)
            out.append(buf)
        return out

    def open(self, index):
        total = int(self.widths.sum())
        opened = np.zeros(total, dtype=np.uint64)
        depth = int(self.heights.max()).bit_length() - 1
        proof = np.zeros((max(depth, 1), 32), dtype=np.uint8)
        self.L.orc_mmcs_open(self.h, index, opened, proof)
        return opened, proof[:depth]

    def verify(self, index, opened, proof):
        return bool(self.L.orc_mmcs_verify(self.root, self.heights, self.widths, len(self.mats), index,
                                           np.ascontiguousarray(opened), np.ascontiguousarray(proof).reshape(-1),
                                           len(proof)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_mmcs_free(self.h)
            self.h = None


def pcs_commit(L, mats, log_blowup):
    """Oracle Pcs::commit; returns (root bytes, handle). Free with L.orc_mmcs_free(handle)."""
    mats = [np.ascontiguousarray(m, dtype=np.uint64) for m in mats]
    n = len(mats)
    ptrs = (C.c_void_p * n)(*[m.ctypes.data for m in mats])
    hs = np.array([m.shape[0] for m in mats], dtype=np.uint64)
    ws = np.array([m.shape[1] for m in mats], dtype=np.uint64)
    root = np.zeros(32, dtype=np.uint8)
    h = L.orc_pcs_commit(ptrs, hs, ws, n, log_blowup, root)
    assert h, "orc_pcs_commit failed"
    return bytes(root), h


VERIFY_ERRORS = {0: "Ok", 1: "InvalidClaim", 2: "InvalidOpeningArgument", 3: "InvalidProofShape", 4: "InvalidSystem",
                 5: "OodEvaluationMismatch", 6: "UnbalancedChannel", -1: "Deserialize"}


def flatten_claims(claims):
    """list of 1-D arrays -> (flat u64 values, offsets[n+1])"""
    offs = np.zeros(len(claims) + 1, dtype=np.uint64)
    for i, c in enumerate(claims):
        offs[i + 1] = offs[i] + len(c)
    flat = np.concatenate([np.asarray(c, dtype=np.uint64).ravel() for c in claims]) if len(claims) else np.zeros(0, dtype=np.uint64)
    if flat.size == 0:
        flat = np.zeros(1, dtype=np.uint64)
    return np.ascontiguousarray(flat), offs


class OracleSystem:
    """The oracle's System: CPU prover (src/prover.rs) and restated verifier (src/verifier.rs)."""

    def __init__(self, L, kind, log_blowup=1, log_final_poly_len=0, max_log_arity=1, num_queries=100, commit_pow_bits=0,
                 query_pow_bits=0, graphs=None, preprocessed=None):
        self.L = L
        if graphs is None:
            self.h = L.orc_system_create(kind.encode(), log_blowup, log_final_poly_len, max_log_arity, num_queries,
                                         commit_pow_bits, query_pow_bits)
        else:  # compiled circuits, same descriptors as multi_stark_b200.System.from_graphs
            from multi_stark_b200.system import graph_descs
            arr, keep = graph_descs(graphs)
            pre = [None if p is None else np.ascontiguousarray(p, dtype=np.uint64) for p in (preprocessed or [None] * len(graphs))]
            ptrs = (C.c_void_p * len(graphs))(*[p.ctypes.data if p is not None else None for p in pre])
            hs = (C.c_uint64 * len(graphs))(*[p.shape[0] if p is not None else 0 for p in pre])
            self.h = L.orc_system_create_from_graphs(C.cast(arr, C.c_void_p), len(graphs), C.cast(ptrs, C.c_void_p),
                                                     C.cast(hs, C.c_void_p), log_blowup, log_final_poly_len, max_log_arity,
                                                     num_queries, commit_pow_bits, query_pow_bits)
        assert self.h, "orc_system_create failed"

    def prove(self, traces, claims):
        """traces: one (h x w) array per circuit (h = 0: inactive). Returns (proof bytes, stage ms[6])."""
        mats = [np.ascontiguousarray(t, dtype=np.uint64) for t in traces]
        n = len(mats)
        ptrs = (C.c_void_p * n)(*[m.ctypes.data if m.size else None for m in mats])
        hs = np.array([m.shape[0] for m in mats], dtype=np.uint64)
        flat, offs = flatten_claims(claims)
        out, ln = C.c_void_p(), C.c_uint64()
        ms = (C.c_double * 6)()
        rc = self.L.orc_prove(self.h, ptrs, hs, flat, offs, len(claims), C.byref(out), C.byref(ln), ms)
        if rc != 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        data = C.string_at(out.value, ln.value)
        self.L.orc_bytes_free(out)
        return data, list(ms)

    def last_transcript(self):
        """(challenges [(c0, c1)...] = beta, gamma, alpha, zeta, alpha_pcs, FRI betas; query indices) of the last prove()."""
        ch = np.zeros(2 * 128, dtype=np.uint64)
        idx = np.zeros(4096, dtype=np.uint64)
        nidx = np.zeros(1, dtype=np.uint64)
        n = int(self.L.orc_last_transcript(ch, 128, idx, 4096, nidx))
        return [(int(ch[2 * i]), int(ch[2 * i + 1])) for i in range(n)], [int(v) for v in idx[:int(nidx[0])]]

    def verify(self, claims, proof_bytes):
        flat, offs = flatten_claims(claims)
        buf = np.frombuffer(proof_bytes, dtype=np.uint8).copy() if len(proof_bytes) else np.zeros(1, dtype=np.uint8)
        return VERIFY_ERRORS.get(self.L.orc_verify(self.h, flat, offs, len(claims), buf, len(proof_bytes)), "Error")

    def preprocessed_commit(self):
        out = np.zeros(32, dtype=np.uint8)
        return bytes(out) if self.L.orc_preprocessed_commit(self.h, out) else None

    def close(self):
        if self.h:
            self.L.orc_system_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pcs_example_prove(L, mats, log_blowup=1, log_final_poly_len=0, num_queries=100, commit_pow=0, query_pow=0, num_open=2):
    """examples/pcs_example.rs on the CPU oracle: returns (root, zeta, bytes of opened values + FRI proof)."""
    mats = [np.ascontiguousarray(m, dtype=np.uint64) for m in mats]
    n = len(mats)
    ptrs = (C.c_void_p * n)(*[m.ctypes.data for m in mats])
    hs = np.array([m.shape[0] for m in mats], dtype=np.uint64)
    ws = np.array([m.shape[1] for m in mats], dtype=np.uint64)
    root = np.zeros(32, dtype=np.uint8)
    zeta = np.zeros(2, dtype=np.uint64)
    out, ln = C.c_void_p(), C.c_uint64()
    rc = L.orc_pcs_example_prove(ptrs, hs, ws, n, log_blowup, log_final_poly_len, num_queries, commit_pow, query_pow, num_open,
                                 root, zeta, C.byref(out), C.byref(ln))
    if rc != 0:
        raise RuntimeError(L.orc_last_error().decode())
    data = C.string_at(out.value, ln.value)
    L.orc_bytes_free(out)
    return bytes(root), (int(zeta[0]), int(zeta[1])), data


def pcs_example_verify(L, root, shapes, data, log_blowup=1, log_final_poly_len=0, num_queries=100, commit_pow=0, query_pow=0,
                       num_open=2):
    hs = np.array([h for h, _ in shapes], dtype=np.uint64)
    ws = np.array([w for _, w in shapes], dtype=np.uint64)
    buf = np.frombuffer(data, dtype=np.uint8).copy() if len(data) else np.zeros(1, dtype=np.uint8)
    r = np.frombuffer(bytes(root), dtype=np.uint8).copy()
    return L.orc_pcs_example_verify(r, hs, ws, len(shapes), log_blowup, log_final_poly_len, num_queries, commit_pow, query_pow,
                                    num_open, buf, len(data))
