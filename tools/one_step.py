"""Exactly two bench steps (stage-1 + stage-2 commit of BASELINE configs[1]) for ncu captures:
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum ... python tools/one_step.py"""
import sys

sys.path.insert(0, ".")
import bench  # noqa: E402
import multi_stark_b200 as ms  # noqa: E402

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20
stages = bench.u32_add_workload(log_rows)
ctx = ms.GpuContext(0)
pcs = ms.GpuPcs(ctx, 1)
dev = [[(ctx.upload(m), m.shape[0], m.shape[1]) for m in st] for st in stages]
for _ in range(2):
    for st in dev:
        root, pd = pcs.commit_dev(st)
        pd.free()
print("steps 2")
