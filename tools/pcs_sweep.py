"""BASELINE configs[4]: standalone PCS sweep (examples/pcs_example.rs shape): commit + open of one 2^log_n x w Goldilocks
matrix, log_blowup 1..3, 100 queries, opened at zeta twice. Reports per-stage times, LDE+Merkle Gelem/s, fractions of the
HBM and integer-pipe rooflines, and checks every opening with the restated verifier (test infrastructure, untimed).
usage: python tools/pcs_sweep.py [--quick] [--out gpurun_out/pcs_sweep.json]"""
import argparse
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import torch  # noqa: E402

import multi_stark_b200 as ms  # noqa: E402
from tests import _oracle as orc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="gpurun_out/pcs_sweep.json")
    ap.add_argument("--max-lde-gb", type=float, default=36.0)
    args = ap.parse_args()
    L = orc.lib()
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = ms.GpuContext(0, stream=stream.cuda_stream)
    peaks = ctx.measure_int_peak()
    try:
        hbm = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        hbm = 6650.0
    sizes = [16, 20] if args.quick else [16, 18, 20, 22, 24]
    widths = [1, 16] if args.quick else [1, 16, 64]
    blowups = [1] if args.quick else [1, 2, 3]
    rows = []
    for log_n in sizes:
        for w in widths:
            for lb in blowups:
                n = 1 << log_n
                lde_gb = n * w * 8 * (1 << lb) / 1e9
                if lde_gb * 1.6 + n * w * 8 * 2 / 1e9 > args.max_lde_gb:
                    continue
                g = torch.Generator(device="cuda")
                g.manual_seed(log_n * 100 + w * 10 + lb)
                m = torch.randint(0, 2**62, (n, w), dtype=torch.int64, device="cuda", generator=g)
                torch.cuda.synchronize()
                pcs = ms.GpuPcs(ctx, lb)
                arg = [(m.data_ptr(), n, w)]
                root, pd = pcs.commit_dev(arg)  # warm-up: builds the twiddle tables of this size
                pd.free()
                reps = 3 if log_n >= 22 else 5
                best, prof_best = None, None
                for _ in range(reps):
                    ctx.profile_begin()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    root, pd = pcs.commit_dev(arg)
                    e1.record()
                    torch.cuda.synchronize()
                    prof = ctx.profile_end()
                    t = e0.elapsed_time(e1)
                    if best is None or t < best:
                        best, prof_best = t, prof
                    pd.free()
                lde_ms = sum(r["ms"] for r in prof_best if r["stage"] == "lde")
                mk_ms = sum(r["ms"] for r in prof_best if r["stage"] == "merkle")
                root, pd = pcs.commit_dev(arg)
                params = dict(log_blowup=lb, num_queries=100)
                t_open = None
                for _ in range(2):
                    ch = ms.Challenger(**params)
                    ch.observe(bytes(root))
                    zeta = ch.sample_algebra_element()
                    t0 = time.perf_counter()
                    data, ms5 = ms.pcs_open(ctx, [(pd, [[zeta, zeta]])], ch)
                    dt = (time.perf_counter() - t0) * 1e3
                    t_open = dt if t_open is None else min(t_open, dt)
                    ch.close()
                ok = orc.pcs_example_verify(L, bytes(root), [(n, w)], data, log_blowup=lb, num_queries=100)
                pd.free()
                del m
                elems = n * w << lb
                alg_lde = 8 * n * w * (1 + (1 << lb))
                alg_mk = 8 * elems + 32 * (n << lb) + 96 * ((n << lb) - 1)
                row = {"log_n": log_n, "w": w, "log_blowup": lb, "commit_ms": best, "lde_ms": lde_ms, "merkle_ms": mk_ms,
                       "open_ms": t_open, "open_phases_ms": ms5, "gelem_s": elems / best / 1e6,
                       "lde_hbm_frac": alg_lde / (lde_ms * 1e-3) / 1e9 / hbm if lde_ms else None,
                       "merkle_hbm_frac": alg_mk / (mk_ms * 1e-3) / 1e9 / hbm if mk_ms else None,
                       "proof_bytes": len(data), "verified": ok == 1}
                rows.append(row)
                print("2^%-2d x %-3d B=%d  commit %8.3f ms (lde %8.3f, merkle %7.3f)  %6.2f Gelem/s  lde %4.1f%% / merkle %4.1f%% of HBM"
                      "  open %7.2f ms  proof %7d B  verified=%s" % (log_n, w, 1 << lb, best, lde_ms, mk_ms, row["gelem_s"],
                                                                  100 * (row["lde_hbm_frac"] or 0), 100 * (row["merkle_hbm_frac"] or 0),
                                                                  t_open, len(data), ok == 1), flush=True)
    json.dump({"hbm_gbs": hbm, "int_peaks_ginst_s": peaks, "rows": rows}, open(args.out, "w"), indent=1)
    assert all(r["verified"] for r in rows)


if __name__ == "__main__":
    main()
