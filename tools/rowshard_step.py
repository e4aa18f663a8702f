"""The bench step (both trace commitments of BASELINE configs[1]) as ONE job over the row shards of all ranks, with a per-kernel
device profile of rank 0 and the collective timings -- the strong-scaling leg of bench.py without the rest of it.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/rowshard_step.py [--log-rows 20] [--steps 10]
       MSGPU_P2P=0 keeps every exchange on NCCL (the peer-memory path is the default)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-rows", type=int, default=20)
    ap.add_argument("--log-blowup", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import multi_stark_b200 as ms
    from multi_stark_b200 import dist as msd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = ms.GpuContext(local_rank, stream=stream.cuda_stream)
    pa = argparse.Namespace(log_blowup=args.log_blowup)
    system = ms.System("u32_add", **bench.prove_params(pa))
    rs = msd.RowShardProver(ctx, system)
    stages = bench.u32_add_workload(args.log_rows, seed=0)
    hs = [[m.shape[0] for m in st] for st in stages]
    ws = [[m.shape[1] for m in st] for st in stages]

    def my_rows(m):
        r0, n = rs.block_rows(m.shape[0], m.shape[1])
        return np.ascontiguousarray(m[r0:r0 + n])
    blocks = [[my_rows(m) for m in st] for st in stages]
    pinned = [[ctx.pinned_copy(b) for b in st] for st in blocks]
    dev = [[ctx.upload(b) for b in st] for st in blocks]

    def step(host):
        return [rs.commit(pinned[i] if host else dev[i], hs[i], ws[i], host) for i in range(len(stages))]

    out = {"world": world, "peer_memory": rs.peer_memory, "log_rows": args.log_rows}
    for host in (False, True):
        for _ in range(args.warmup):
            roots = step(host)
        msd.barrier()
        torch.cuda.synchronize()
        rs.comm.seconds.clear()
        ctx.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            roots = step(host)
        e1.record()
        torch.cuda.synchronize()
        prof = ctx.profile_end()
        key = "host" if host else "resident"
        out[key + "_ms_per_step"] = msd.max_over_ranks(e0.elapsed_time(e1)) / args.steps
        out[key + "_kernels_ms_per_step"] = {"%s/%s" % (r["stage"], r["kernel"]): round(r["ms"] / args.steps, 4) for r in prof}
        out[key + "_kernel_sum_ms"] = round(sum(r["ms"] for r in prof) / args.steps, 4)
        out[key + "_comm_host_ms_per_step"] = {k: round(v * 1e3 / args.steps, 4) for k, v in rs.comm.seconds.items()}
        msd.barrier()
    elems = bench.committed_elements(stages, args.log_blowup)
    out["gelem_per_s_resident"] = elems / (out["resident_ms_per_step"] / 1e3) / 1e9
    out["gelem_per_s_host"] = elems / (out["host_ms_per_step"] / 1e3) / 1e9
    if rank == 0:
        pcs = ms.GpuPcs(ctx, args.log_blowup)
        want = []
        for st in stages:
            root, pd = pcs.commit([np.ascontiguousarray(m) for m in st])
            want.append(bytes(root))
            pd.free()
        out["roots_equal_single_gpu_commit"] = want == roots
        print(json.dumps(out))
    rs.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
