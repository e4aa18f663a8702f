"""Independent Python restatement of the proof wire format, for the tests: `Proof` field order of
src/prover.rs:213-238 under bincode 2 `standard().with_little_endian().with_fixed_int_encoding()`
(src/prover.rs:241-243): u64-LE length prefixes, fixed-width LE integers, bool/u8 one byte, Option tag byte,
Goldilocks as canonical u64, extension element = 2 coordinates, 32-byte digests raw. The FRI part follows p3-fri 0.5.1's
`FriProof { commit_phase_commits, commit_pow_witnesses, query_proofs, final_poly, query_pow_witness }`,
`QueryProof { input_proof: Vec<BatchOpening{opened_values, opening_proof}>, commit_phase_openings:
Vec<CommitPhaseProofStep{log_arity, sibling_values, opening_proof}> }` (PARITY UNPINNED)."""
import struct


CAP_WIRE = False  # True: commitments as a Merkle cap Vec<[u8; 32]> (u64 length 1 + digest), see host/proof.hpp


class R:
    def __init__(self, b):
        self.b, self.o = b, 0

    def commitment(self):
        if CAP_WIRE:
            assert self.u64() == 1
        return self.digest()

    def u8(self):
        v = self.b[self.o]
        self.o += 1
        return v

    def u64(self):
        v = struct.unpack_from("<Q", self.b, self.o)[0]
        self.o += 8
        return v

    def ext(self):
        return (self.u64(), self.u64())

    def digest(self):
        v = bytes(self.b[self.o:self.o + 32])
        assert len(v) == 32
        self.o += 32
        return v

    def vec(self, f):
        return [f() for _ in range(self.u64())]


def _opened_round(r):
    return r.vec(lambda: r.vec(lambda: r.vec(r.ext)))


def parse(data):
    r = R(data)
    p = {}
    p["active"] = r.vec(r.u8)
    p["stage_1_trace"], p["stage_2_trace"], p["quotient_chunks"] = r.commitment(), r.commitment(), r.commitment()
    p["intermediate_accumulators"] = r.vec(r.ext)
    p["log_degrees"] = r.vec(r.u8)
    f = {}
    f["commit_phase_commits"] = r.vec(r.commitment)
    f["commit_pow_witnesses"] = r.vec(r.u64)

    def query():
        q = {}
        q["input_proof"] = r.vec(lambda: {"opened_values": r.vec(lambda: r.vec(r.u64)), "opening_proof": r.vec(r.digest)})
        q["commit_phase_openings"] = r.vec(lambda: {"log_arity": r.u8(), "sibling_values": r.vec(r.ext),
                                                    "opening_proof": r.vec(r.digest)})
        return q
    f["query_proofs"] = r.vec(query)
    f["final_poly"] = r.vec(r.ext)
    f["query_pow_witness"] = r.u64()
    p["opening_proof"] = f
    p["quotient_opened_values"] = _opened_round(r)
    p["preprocessed_opened_values"] = _opened_round(r) if r.u8() == 1 else None
    p["stage_1_opened_values"] = _opened_round(r)
    p["stage_2_opened_values"] = _opened_round(r)
    assert r.o == len(data), "trailing bytes"
    return p


class W:
    def __init__(self):
        self.out = bytearray()

    def u8(self, v):
        self.out.append(v)

    def u64(self, v):
        self.out += struct.pack("<Q", v)

    def ext(self, v):
        self.u64(v[0])
        self.u64(v[1])

    def digest(self, d):
        self.out += d

    def commitment(self, d):
        if CAP_WIRE:
            self.u64(1)
        self.out += d

    def vec(self, v, f):
        self.u64(len(v))
        for x in v:
            f(x)


def _w_opened_round(w, rnd):
    w.vec(rnd, lambda m: w.vec(m, lambda p: w.vec(p, w.ext)))


def serialize(p):
    w = W()
    w.vec(p["active"], w.u8)
    w.commitment(p["stage_1_trace"])
    w.commitment(p["stage_2_trace"])
    w.commitment(p["quotient_chunks"])
    w.vec(p["intermediate_accumulators"], w.ext)
    w.vec(p["log_degrees"], w.u8)
    f = p["opening_proof"]
    w.vec(f["commit_phase_commits"], w.commitment)
    w.vec(f["commit_pow_witnesses"], w.u64)

    def query(q):
        def bo(b):
            w.vec(b["opened_values"], lambda row: w.vec(row, w.u64))
            w.vec(b["opening_proof"], w.digest)
        w.vec(q["input_proof"], bo)

        def step(s):
            w.u8(s["log_arity"])
            w.vec(s["sibling_values"], w.ext)
            w.vec(s["opening_proof"], w.digest)
        w.vec(q["commit_phase_openings"], step)
    w.vec(f["query_proofs"], query)
    w.vec(f["final_poly"], w.ext)
    w.u64(f["query_pow_witness"])
    _w_opened_round(w, p["quotient_opened_values"])
    if p["preprocessed_opened_values"] is None:
        w.u8(0)
    else:
        w.u8(1)
        _w_opened_round(w, p["preprocessed_opened_values"])
    _w_opened_round(w, p["stage_1_opened_values"])
    _w_opened_round(w, p["stage_2_opened_values"])
    return bytes(w.out)
