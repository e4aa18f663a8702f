"""BASELINE configs[2] on ONE GPU: the wide synthetic AIR (W columns x 2^log_rows rows, log_blowup 2, no lookups):
prove() from pinned host memory, per-stage times, proof checked by the restated verifier.
usage: python tools/wide_prove.py [log_rows=20] [width=256] [log_blowup=2]"""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import multi_stark_b200 as ms  # noqa: E402
from tests import _oracle as orc  # noqa: E402

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20
width = int(sys.argv[2]) if len(sys.argv) > 2 else 256
lb = int(sys.argv[3]) if len(sys.argv) > 3 else 2
kw = dict(log_blowup=lb, num_queries=100)
ctx = ms.GpuContext(0)
system = ms.System("wide:%d" % width, **kw)
prover = ms.Prover(ctx, system)
t0 = time.perf_counter()
trace = ctx.pinned_empty((1 << log_rows, width))
ms.wide_trace(1 << log_rows, width, out=trace)
print("trace %.1f MB generated in %.1f s" % (trace.nbytes / 1e6, time.perf_counter() - t0), flush=True)
best = None
for i in range(3):
    t0 = time.perf_counter()
    proof = prover.prove([trace], [])
    dt = (time.perf_counter() - t0) * 1e3
    best = dt if best is None else min(best, dt)
    print("prove %d x 2^%d (B=%d): %.2f ms, %d proof bytes, stages %s" % (
        width, log_rows, 1 << lb, dt, len(proof), json.dumps({k: round(v, 2) for k, v in prover.last_stage_ms.items()})), flush=True)
ctx.profile_begin()
proof = prover.prove([trace], [])
prof = ctx.profile_end()
for r in sorted(prof, key=lambda r: -r["ms"])[:12]:
    print("%-10s %-26s x%-4d %9.3f ms" % (r["stage"], r["kernel"], r["launches"], r["ms"]))
L = orc.lib()
S = orc.OracleSystem(L, "wide:%d" % width, **kw)
print("verified:", S.verify([], proof))
elems = (width + 2 + 4) * (1 << (log_rows + lb))
print(json.dumps({"width": width, "log_rows": log_rows, "log_blowup": lb, "prove_ms": best, "committed_elements": elems,
                  "gelem_per_s_whole_prove": elems / best / 1e6}))
