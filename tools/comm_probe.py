"""Latency of the sharded prover's collectives (TorchComm callbacks) at the sizes a proof uses. Run one process per GPU:
RANK/WORLD_SIZE/MASTER_* in the environment (tools/dist_prove.py style) or under torchrun."""
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import multi_stark_b200 as ms  # noqa: E402
from multi_stark_b200 import dist as msd  # noqa: E402

rank = int(os.environ["RANK"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = ms.GpuContext(rank, stream=stream.cuda_stream)
comm = msd.TorchComm(ctx)
st = comm.struct
w = comm.world
for nbytes in (32, 4096, 1 << 20, 2 << 20):
    send = np.zeros(nbytes, dtype=np.uint8)
    recv = np.zeros(nbytes * w, dtype=np.uint8)
    for it in range(4):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st.allgather_host(None, send.ctypes.data, recv.ctypes.data, nbytes)
        t1 = time.perf_counter()
        st.bcast_host(None, send.ctypes.data, nbytes, 0)
        t2 = time.perf_counter()
        if rank == 0:
            print("host %8d B  it %d  allgather %.3f ms  bcast %.3f ms" % (nbytes, it, (t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
for nbytes in (1 << 20, 32 << 20):
    buf = ctx.malloc(nbytes)
    for it in range(4):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st.sendrecv_dev(None, buf, nbytes, 1, 0)
        t1 = time.perf_counter()
        if rank == 0:
            print("dev  %8d B  it %d  sendrecv %.3f ms" % (nbytes, it, (t1 - t0) * 1e3), flush=True)
print(rank, comm.errors)
dist.destroy_process_group()
