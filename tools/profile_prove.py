"""Per-stage / per-kernel timing of one device prove() (CUDA events per launch via msgpu_profile_*).
usage: python tools/profile_prove.py [log_rows] [log_blowup] [reps]"""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import multi_stark_b200 as ms  # noqa: E402

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = ms.GpuContext(0)
system = ms.System("u32_add", log_blowup=lb, num_queries=100)
t0 = time.perf_counter()
prover = ms.Prover(ctx, system)
byte, add, claims = ms.u32_add_workload(1 << log_rows)
if "--pageable" not in sys.argv:
    byte, add, claims = ctx.pinned_copy(byte), ctx.pinned_copy(add), ctx.pinned_copy(claims)
print("setup %.1f ms" % ((time.perf_counter() - t0) * 1e3))
for i in range(reps):
    t0 = time.perf_counter()
    proof = prover.prove([byte, add], claims)
    dt = (time.perf_counter() - t0) * 1e3
    print("prove 2^%d: %.2f ms wall, %d proof bytes; stages: %s" % (log_rows, dt, len(proof), json.dumps({k: round(v, 2) for k, v in prover.last_stage_ms.items()})))
ctx.profile_begin()
l0 = ctx.launches
proof = prover.prove([byte, add], claims)
prof = ctx.profile_end()
print("launches", ctx.launches - l0)
tot = sum(r["ms"] for r in prof)
print("sum of kernel ms %.3f" % tot)
for r in sorted(prof, key=lambda r: -r["ms"]):
    print("%-8s %-28s x%-4d %8.3f ms" % (r["stage"], r["kernel"], r["launches"], r["ms"]))
