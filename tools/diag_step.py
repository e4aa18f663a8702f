"""Where does a resident bench step spend its time? Wall clock per call next to the per-launch CUDA events.
usage: python tools/diag_step.py [log_rows] [--sampler]"""
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402
import multi_stark_b200 as ms  # noqa: E402

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20
stages = bench.u32_add_workload(log_rows)
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = ms.GpuContext(0, stream=stream.cuda_stream)
pcs = ms.GpuPcs(ctx, 1)
res = [[torch.from_numpy(m.view("int64")).cuda() for m in st] for st in stages]
dev = [[(t.data_ptr(), t.shape[0], t.shape[1]) for t in st] for st in res]
torch.cuda.synchronize()
sampler = None
if "--sampler" in sys.argv:
    sampler = bench.ClockSampler(0)
    sampler.start()
    time.sleep(1.0)
nv = None
if "--nvml" in sys.argv:
    import threading
    import pynvml
    pynvml.nvmlInit()
    hdl = pynvml.nvmlDeviceGetHandleByIndex(0)
    nv = {"stop": False, "sm": [], "reasons": 0}

    def poll():
        while not nv["stop"]:
            nv["sm"].append(pynvml.nvmlDeviceGetClockInfo(hdl, pynvml.NVML_CLOCK_SM))
            nv["reasons"] |= pynvml.nvmlDeviceGetCurrentClocksEventReasons(hdl)
            time.sleep(0.05)
    th = threading.Thread(target=poll, daemon=True)
    th.start()
    time.sleep(0.5)
for it in range(5):
    ctx.profile_begin()
    t0 = time.perf_counter()
    walls = []
    for st in dev:
        a = time.perf_counter()
        root, pd = pcs.commit_dev(st)
        b = time.perf_counter()
        pd.free()
        c = time.perf_counter()
        walls.append((round((b - a) * 1e3, 3), round((c - b) * 1e3, 3)))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    prof = ctx.profile_end()
    print("step %d: wall %.3f ms, kernels %.3f ms, (commit, free) per stage %s" % (it, (t1 - t0) * 1e3, sum(r["ms"] for r in prof), walls))
if sampler:
    print(sampler.stop())
if nv:
    nv["stop"] = True
    th.join()
    print("nvml", len(nv["sm"]), sorted(set(nv["sm"])), hex(nv["reasons"]))
