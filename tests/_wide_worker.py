"""One rank of a column-block sharded commit (tests/test_gpu_dist_prove.py, tools/wide_commit_sharded.py).
argv: backend log_rows width log_blowup out_prefix [reps]"""
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
    c0, c1 = msd.column_blocks(width, world)[rank]
    # every rank generates its own column block of the same counter-based matrix: entry (r, c) = splitmix64(r * width + c) mod p
    full_needed = rank == 0 and os.environ.get("WIDE_SINGLE", "1") == "1"
    idx = (np.arange(n, dtype=np.uint64)[:, None] * np.uint64(width) + np.arange(c0, c1, dtype=np.uint64)[None, :])
    block = ctx.pinned_copy(splitmix(idx) % np.uint64(ms.P))
    times, tm = [], {}
    handle = None
    for it in range(reps):
        if handle is not None:
            handle.free()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        root, handle = msd.commit_wide_sharded(ctx, comm, block, width, lb, timings=tm if it == reps - 1 else None)
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    H = n << lb
    rng = np.random.default_rng(7)
    indices = sorted({0, 1, H - 1, H // 2, H // world - 1, H // world % H} | {int(x) for x in rng.integers(0, H, size=8)})
    rows, paths = handle.open_batch(indices)
    info = {"rank": rank, "root": root.hex(), "ms": times, "timings": tm, "bytes_dev": comm.bytes_dev // reps,
            "launches": ctx.launches, "errors": comm.errors}
    np.savez("%s.rank%d.npz" % (out_prefix, rank), rows=rows, paths=paths, indices=np.array(indices, dtype=np.uint64))
    if full_needed:
        idx = (np.arange(n, dtype=np.uint64)[:, None] * np.uint64(width) + np.arange(width, dtype=np.uint64)[None, :])
        full = ctx.pinned_copy(splitmix(idx) % np.uint64(ms.P))
        del idx
        pcs = ms.GpuPcs(ctx, lb)
        ts = []
        for it in range(max(reps, 1)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            want_root, pd = pcs.commit([full])
            ts.append((time.perf_counter() - t0) * 1e3)
            if it + 1 < max(reps, 1):
                pd.free()
        wrows, wpaths = pd.open_batch(indices)
        info["single_root"] = bytes(want_root).hex()
        info["single_ms"] = ts
        info["openings_identical"] = bool(np.array_equal(wrows, rows) and np.array_equal(wpaths, paths))
        pd.free()
    with open("%s.rank%d.json" % (out_prefix, rank), "w") as f:
        json.dump(info, f)
    dist.barrier()
    handle.free()
    ctx.close()
    dist.destroy_process_group()


def splitmix(x):
    x = x + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


if __name__ == "__main__":
    main()
