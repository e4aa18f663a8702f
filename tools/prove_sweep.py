"""BASELINE metric (1): prove() wall ms (pinned host traces and claims in, Proof::to_bytes out) for the U32-add system at
2^16 .. 2^24 rows on one GPU, per-stage split, proofs checked by the restated verifier up to 2^22 rows.
usage: python tools/prove_sweep.py [--out gpurun_out/prove_sweep.json] [--max-log-rows 24]"""
import argparse
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

import multi_stark_b200 as ms  # noqa: E402
from tests import _oracle as orc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/prove_sweep.json")
ap.add_argument("--max-log-rows", type=int, default=24)
ap.add_argument("--log-blowup", type=int, default=1)
args = ap.parse_args()
kw = dict(log_blowup=args.log_blowup, num_queries=100)
ctx = ms.GpuContext(0)
system = ms.System("u32_add", **kw)
prover = ms.Prover(ctx, system)
L = orc.lib()
S = orc.OracleSystem(L, "u32_add", **kw)
rows = []
for log_rows in range(16, args.max_log_rows + 1, 2):
    byte, add, claims = ms.u32_add_workload(1 << log_rows)
    byte, add, claims = ctx.pinned_copy(byte), ctx.pinned_copy(add), ctx.pinned_copy(claims)
    times = []
    for it in range(5 if log_rows <= 22 else 3):
        t0 = time.perf_counter()
        proof = prover.prove([byte, add], claims)
        times.append((time.perf_counter() - t0) * 1e3)
    ctx.profile_begin()
    prover.prove([byte, add], claims)
    kern = sum(r["ms"] for r in ctx.profile_end())
    rec = {"log_rows": log_rows, "prove_ms": float(np.min(times[1:])), "prove_ms_median": float(np.median(times[1:])),
           "stages_ms": {k: round(v, 3) for k, v in prover.last_stage_ms.items()}, "kernel_ms": kern,
           "h2d_mb": (byte.nbytes + add.nbytes + claims.nbytes) / 1e6, "proof_bytes": len(proof),
           "committed_elements": (43 << log_rows) * (1 << args.log_blowup)}
    if log_rows <= 22:
        rec["verified"] = S.verify(claims, proof) == "Ok"
    rows.append(rec)
    print(json.dumps(rec), flush=True)
    del byte, add, claims
json.dump({"system": "u32_add", "params": kw, "rows": rows}, open(args.out, "w"), indent=1)
