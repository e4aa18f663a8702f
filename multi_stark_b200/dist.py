"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL on the GPU box, gloo in the CPU tests).

The commitment path shards by independent units -- whole proofs, or the circuit groups of SURVEY section 8(e) partitioning A --
so there is NO data-path collective: every rank runs the full single-GPU pipeline on its own units. Ranks exchange only
what the protocol makes public anyway (32-byte commitments / proof digests) and the timing needed for an honest aggregate
(max over ranks). The reference has no multi-process story at all (rayon inside one process)."""
import hashlib

import torch
import torch.distributed as dist


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
