"""One rank of a column-block sharded prove of the wide AIR (tests/test_gpu_dist_prove.py, tools/wide_prove_sharded.py).
argv: backend log_rows width log_blowup out_prefix [reps] [num_queries]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import multi_stark_b200 as ms  # noqa: E402
from multi_stark_b200 import dist as msd  # noqa: E402


def main():
    backend, log_rows, width, lb, out_prefix = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    reps = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    nq = int(sys.argv[7]) if len(sys.argv) > 7 else 100
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    if backend == "nccl":
        os.environ.setdefault("TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING", "false")
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = ms.GpuContext(dev, stream=stream.cuda_stream)
    comm = msd.TorchComm(ctx)
    n = 1 << log_rows
    kw = dict(log_blowup=lb, num_queries=nq)
    system = ms.System("wide:%d" % width, **kw)
    prover = ms.Prover(ctx, system) if rank == 0 else None
    c0, c1 = msd.column_blocks(width, world)[rank]
    block = ctx.pinned_empty((n, c1 - c0))
    ms.wide_trace(n, width, out=block, cols=(c0, c1))
    times, tm = [], {}
    for it in range(reps):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        proof = msd.prove_wide_sharded(ctx, comm, prover, block, width, lb, timings=tm)
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
        dist.barrier()
    info = {"rank": rank, "ms": times, "timings": tm, "errors": comm.errors, "bytes_dev": comm.bytes_dev // reps,
            "launches": ctx.launches}
    if rank == 0:
        info["stages"] = prover.last_stage_ms
        info["proof_bytes"] = len(proof)
        with open(out_prefix + ".sharded.proof", "wb") as f:
            f.write(proof)
        if os.environ.get("WIDE_SINGLE", "1") == "1":
            full = ctx.pinned_empty((n, width))
            ms.wide_trace(n, width, out=full)
            ts = []
            for it in range(max(reps, 1)):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                want = prover.prove([full], [])
                ts.append((time.perf_counter() - t0) * 1e3)
            info["single_ms"] = ts
            info["single_stages"] = prover.last_stage_ms
            info["identical"] = want == proof
            with open(out_prefix + ".single.proof", "wb") as f:
                f.write(want)
        prover.close()
    with open("%s.rank%d.json" % (out_prefix, rank), "w") as f:
        json.dump(info, f)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
