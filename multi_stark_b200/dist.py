"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL on the GPU box, gloo in the CPU tests).

The commitment path shards by independent units -- whole proofs, or the circuit groups of SURVEY section 8(e) partitioning A --
so there is NO data-path collective: every rank runs the full single-GPU pipeline on its own units. Ranks exchange only
what the protocol makes public anyway (32-byte commitments / proof digests) and the timing needed for an honest aggregate
(max over ranks). The reference has no multi-process story at all (rayon inside one process)."""
import hashlib
import os
import sys
import time

import torch
import torch.distributed as dist

from ._ffi import check


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _device():
    """tensors for collectives live where the backend wants them"""
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def shard_units(n_units, world_size=None, r=None):
    """Contiguous balanced partition of unit ids [0, n_units): the first n_units % world ranks get one more."""
    w = world() if world_size is None else world_size
    r = rank() if r is None else r
    if w <= 0 or not 0 <= r < w:
        raise ValueError("bad rank / world size")
    base, rem = divmod(n_units, w)
    lo = r * base + min(r, rem)
    return range(lo, lo + base + (1 if r < rem else 0))


def barrier():
    if world() > 1:
        dist.barrier()


def max_over_ranks(value):
    """Device-side max of a per-rank scalar (elapsed ms): the job takes as long as its slowest rank."""
    if world() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value):
    if world() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_digests(digests):
    """All-gather of this rank's 32-byte digests (commitments or SHA-256 of proof bytes), one per local unit.
    Returns the list over all ranks in unit order. Every rank must contribute the same count per unit shard rule."""
    w = world()
    if w == 1:
        return list(digests)
    counts = torch.tensor([len(digests)], dtype=torch.int64, device=_device())
    all_counts = [torch.zeros_like(counts) for _ in range(w)]
    dist.all_gather(all_counts, counts)
    cap = max(int(c.item()) for c in all_counts)
    buf = torch.zeros((max(cap, 1), 32), dtype=torch.uint8)
    for i, d in enumerate(digests):
        if len(d) != 32:
            raise ValueError("digests must be 32 bytes")
        buf[i] = torch.frombuffer(bytearray(d), dtype=torch.uint8)
    buf = buf.to(_device())
    out = [torch.zeros_like(buf) for _ in range(w)]
    dist.all_gather(out, buf)
    res = []
    for r_, c in enumerate(all_counts):
        rows = out[r_].cpu()
        for i in range(int(c.item())):
            res.append(bytes(rows[i].tolist()))
    return res


def proof_digest(proof_bytes):
    return hashlib.sha256(proof_bytes).digest()


def prove_sharded(units, prove_fn):
    """Runs prove_fn(unit) for the units of this rank's shard; returns (local {unit_id: proof}, digests of ALL units in
    unit order, gathered over the ranks). `units` is the same list on every rank."""
    mine = shard_units(len(units))
    proofs = {u: prove_fn(units[u]) for u in mine}
    digests = gather_digests([proof_digest(proofs[u]) for u in mine])
    return proofs, digests


# ------------------------------------------------------------------------------------------------------------------
# One proof over several GPUs: the collectives the sharded prover (host/dist_backend.hpp) calls back into.
# ------------------------------------------------------------------------------------------------------------------
import ctypes as _C  # noqa: E402

_ALLGATHER = _C.CFUNCTYPE(_C.c_int, _C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_uint64)
_BCAST = _C.CFUNCTYPE(_C.c_int, _C.c_void_p, _C.c_void_p, _C.c_uint64, _C.c_int32)
_SENDRECV = _C.CFUNCTYPE(_C.c_int, _C.c_void_p, _C.c_void_p, _C.c_uint64, _C.c_int32, _C.c_int32)


class MshComm(_C.Structure):
    """`msh_comm` of host/dist_backend.hpp"""
    _fields_ = [("user", _C.c_void_p), ("rank", _C.c_int32), ("world", _C.c_int32), ("allgather_host", _ALLGATHER),
                ("bcast_host", _BCAST), ("sendrecv_dev", _SENDRECV)]


class _DevView:
    """A raw device range as something torch.as_tensor understands (zero copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2,
                                         "strides": None}


class TorchComm:
    """Collectives over the default torch.distributed group. Host buffers travel as CPU tensors over gloo, or staged through
    the device over NCCL; device buffers travel device-to-device over NCCL (NVLink), or staged through pinned host memory
    over gloo (the single-GPU tests run two ranks on one device that way)."""

    def __init__(self, ctx, group=None):
        import torch.distributed as dist
        self.ctx = ctx
        self.dist = dist
        self.group = group
        self.nccl = dist.get_backend(group) == "nccl"
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.bytes_dev = 0   # device bytes sent or received by this rank (reported by bench / tools)
        self.bytes_host = 0
        self.errors = []
        self.seconds = {}    # wall time spent inside the callbacks, by kind
        self._stage = None
        self.trace = bool(os.environ.get("MSH_TRACE"))
        self._cb = (_ALLGATHER(self._allgather), _BCAST(self._bcast), _SENDRECV(self._sendrecv))  # keep the thunks alive
        self.struct = MshComm(None, self.rank, self.world, *self._cb)

    # ---- host buffers
    def _host_tensor(self, ptr, nbytes):
        return torch.frombuffer((_C.c_uint8 * nbytes).from_address(ptr), dtype=torch.uint8)

    def _allgather(self, user, send, recv, nbytes):
        t0 = time.perf_counter()
        try:
            return self._allgather_impl(send, recv, nbytes)
        finally:
            self.seconds["allgather"] = self.seconds.get("allgather", 0.0) + time.perf_counter() - t0

    def _staging(self, nbytes):
        """persistent pinned-host and device staging buffers (grow-only): a fresh pageable tensor per call costs ms"""
        if self._stage is None or self._stage[0].numel() < nbytes:
            cap = 1 << max(16, int(nbytes - 1).bit_length())
            dev = torch.device("cuda", torch.cuda.current_device())
            self._stage = (torch.empty(cap, dtype=torch.uint8).pin_memory(), torch.empty(cap, dtype=torch.uint8, device=dev))
        return self._stage

    def _allgather_impl(self, send, recv, nbytes):
        try:
            nbytes = int(nbytes)
            s = self._host_tensor(send, nbytes)
            r = self._host_tensor(recv, nbytes * self.world)
            if self.nccl:
                tr = [time.perf_counter()]
                pin, dev = self._staging(nbytes * (self.world + 1))
                tr.append(time.perf_counter())
                total = nbytes * self.world
                _C.memmove(pin.data_ptr(), send, nbytes)  # (torch's CPU copy_ spins up its thread pool: ms for 1 MB)
                dev[total:total + nbytes].copy_(pin[:nbytes], non_blocking=True)
                tr.append(time.perf_counter())
                self.dist.all_gather_into_tensor(dev[:total], dev[total:total + nbytes], group=self.group)
                tr.append(time.perf_counter())
                pin[:total].copy_(dev[:total], non_blocking=True)
                torch.cuda.current_stream().synchronize()
                tr.append(time.perf_counter())
                _C.memmove(recv, pin.data_ptr(), total)
                tr.append(time.perf_counter())
                if self.trace:
                    print("[comm] rank %d allgather %d B: staging %.3f, h2d %.3f, nccl call %.3f, d2h+sync %.3f, copy out %.3f ms" % (
                        (self.rank, nbytes) + tuple((b - a) * 1e3 for a, b in zip(tr, tr[1:]))), file=sys.stderr, flush=True)
            else:
                self.dist.all_gather_into_tensor(r, s.clone(), group=self.group)
            self.bytes_host += nbytes * self.world
            return 0
        except Exception as e:  # never unwind into C++
            self.errors.append(repr(e))
            return 1

    def _bcast(self, user, buf, nbytes, root):
        t0 = time.perf_counter()
        try:
            nbytes = int(nbytes)
            t = self._host_tensor(buf, nbytes)
            if self.nccl:
                pin, dev = self._staging(nbytes)
                if self.rank == root:
                    _C.memmove(pin.data_ptr(), buf, nbytes)
                    dev[:nbytes].copy_(pin[:nbytes], non_blocking=True)
                self.dist.broadcast(dev[:nbytes], src=self._global(root), group=self.group)
                if self.rank != root:
                    pin[:nbytes].copy_(dev[:nbytes], non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                    _C.memmove(buf, pin.data_ptr(), nbytes)
            else:
                self.dist.broadcast(t, src=self._global(root), group=self.group)
            self.bytes_host += nbytes
            return 0
        except Exception as e:
            self.errors.append(repr(e))
            return 1
        finally:
            self.seconds["bcast"] = self.seconds.get("bcast", 0.0) + time.perf_counter() - t0

    def _global(self, r):
        return self.dist.get_global_rank(self.group, r) if self.group is not None else r

    # ---- device buffers
    def _sendrecv(self, user, dev, nbytes, src, dst):
        t0 = time.perf_counter()
        try:
            return self._sendrecv_impl(dev, nbytes, src, dst)
        finally:
            self.seconds["sendrecv"] = self.seconds.get("sendrecv", 0.0) + time.perf_counter() - t0

    def _sendrecv_impl(self, dev, nbytes, src, dst):
        try:
            nbytes = int(nbytes)
            self.bytes_dev += nbytes
            if self.nccl:
                t = torch.as_tensor(_DevView(dev, nbytes), device=torch.device("cuda", torch.cuda.current_device()))
                if self.rank == src:
                    self.dist.send(t, dst=self._global(dst), group=self.group)
                else:
                    self.dist.recv(t, src=self._global(src), group=self.group)
                torch.cuda.current_stream().synchronize()
                return 0
            # gloo: stage through host memory
            host = torch.empty(nbytes, dtype=torch.uint8)
            if self.rank == src:
                check(self.ctx.L.msgpu_memcpy_d2h(self.ctx.h, _C.c_void_p(host.data_ptr()), _C.c_void_p(dev), nbytes))
                self.dist.send(host, dst=self._global(dst), group=self.group)
            else:
                self.dist.recv(host, src=self._global(src), group=self.group)
                check(self.ctx.L.msgpu_memcpy_h2d(self.ctx.h, _C.c_void_p(dev), _C.c_void_p(host.data_ptr()), nbytes))
            return 0
        except Exception as e:
            self.errors.append(repr(e))
            return 1


def assign_owners(heights, world_size):
    """Circuit -> rank for one sharded proof. Circuits of one trace height must share a rank (their rows share leaf digests
    and reduced openings); height classes are placed greedily, heaviest first, on the least loaded rank. `heights[i]` = trace
    rows of circuit i (0 = inactive; such circuits go to rank 0), weights = rows (a proxy for the LDE + Merkle work).
    Deterministic, so every rank computes the same assignment."""
    classes = {}
    for i, h in enumerate(heights):
        classes.setdefault(int(h), []).append(i)
    load = [0] * world_size
    owner = [0] * len(heights)
    for h in sorted(classes, reverse=True):
        if h == 0:
            continue
        r = min(range(world_size), key=lambda k: (load[k], k))
        load[r] += h * len(classes[h])
        for i in classes[h]:
            owner[i] = r
    return owner


class DistProver:
    """`System::prove_multiple_claims` (src/prover.rs:289-603) as ONE proof over the ranks of a torch.distributed group: each
    rank owns whole circuits (`owner[i]`), builds and commits their traces and answers their openings; the proof bytes are
    identical on every rank and identical to the single-GPU proof."""

    def __init__(self, ctx, system, owner, group=None):
        from . import _ffi
        self.H = _ffi.host_lib()
        self.ctx, self.system = ctx, system
        self.comm = TorchComm(ctx, group)
        self.owner = [int(o) for o in owner]
        if len(self.owner) != system.num_circuits:
            raise ValueError("one owner per circuit expected")
        arr = (_C.c_int32 * len(self.owner))(*self.owner)
        self.h = self.H.msh_dist_prover_create(system.h, ctx.h, _C.byref(self.comm.struct), arr)
        if not self.h:
            raise _ffi.MsgpuError(-3, (self.H.msh_last_error() or b"").decode() + " " + "; ".join(self.comm.errors))
        self.last_stage_ms = None

    def preprocessed_commit(self):
        import numpy as np
        out = np.zeros(32, dtype=np.uint8)
        return bytes(out) if self.H.msh_prover_preprocessed_commit(self.h, out.ctypes.data_as(_C.c_void_p)) else None

    def prove(self, traces, heights, claims):
        """traces[i]: (h x main_width) uint64 array for the circuits this rank owns, None otherwise; heights[i]: trace rows of
        EVERY circuit (0 = inactive); claims: (n, len) uint64 array, the same on every rank. Returns `Proof::to_bytes`."""
        import numpy as np
        from . import _ffi
        from .system import STAGE_NAMES
        n = self.system.num_circuits
        mats = [None if t is None else np.ascontiguousarray(t, dtype=np.uint64) for t in traces]
        for i in range(n):
            if self.owner[i] == self.comm.rank and heights[i] and (mats[i] is None or mats[i].shape[0] != heights[i]):
                raise ValueError("the trace of circuit %d (owned by this rank) is missing or has the wrong height" % i)
        ptrs = (_C.c_void_p * n)(*[m.ctypes.data if (m is not None and m.size and self.owner[i] == self.comm.rank) else None
                                   for i, m in enumerate(mats)])
        hs = (_C.c_uint64 * n)(*[int(h) for h in heights])
        cl = np.ascontiguousarray(claims, dtype=np.uint64)
        if cl.ndim != 2:
            raise ValueError("claims: an (n, len) array is expected")
        flat = cl.reshape(-1) if cl.size else np.zeros(1, dtype=np.uint64)
        out, ln = _C.c_void_p(), _C.c_uint64()
        ms = (_C.c_double * 6)()
        rc = self.H.msh_prove(self.h, ptrs, hs, flat.ctypes.data_as(_ffi.c_u64p), None, cl.shape[1], cl.shape[0],
                              _C.byref(out), _C.byref(ln), ms)
        if rc != 0:
            raise _ffi.MsgpuError(rc, (self.H.msh_last_error() or b"").decode() + " " + "; ".join(self.comm.errors))
        data = _C.string_at(out.value, ln.value)
        self.H.msh_bytes_free(out)
        self.last_stage_ms = dict(zip(STAGE_NAMES, ms))
        return data

    def close(self):
        if getattr(self, "h", None):
            self.H.msh_prover_free(self.h)
            self.h = None
