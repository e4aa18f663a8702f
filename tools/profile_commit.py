"""Small driver for ncu captures: one warm-up + one measured Pcs::commit of a 2^log_rows x w matrix."""
import sys
import numpy as np
sys.path.insert(0, ".")
import multi_stark_b200 as ms

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20
w = int(sys.argv[2]) if len(sys.argv) > 2 else 14
lb = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rng = np.random.default_rng(0)
m = rng.integers(0, ms.P, size=(1 << log_rows, w), dtype=np.uint64)
ctx = ms.GpuContext(0)
pcs = ms.GpuPcs(ctx, lb)
d = ctx.upload(m)
for _ in range(2):
    root, pd = pcs.commit_dev([(d, m.shape[0], w)])
    pd.free()
print(bytes(root).hex())
