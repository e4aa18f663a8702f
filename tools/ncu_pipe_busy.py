"""Per-kernel integer-pipe utilisation from an `ncu --set full` report: ALU pipe and FMA-heavy pipe (the one that executes IMAD /
IMAD.WIDE) busy percentages, issue-slot utilisation and occupancy, averaged over the launches of each kernel.
usage: python tools/ncu_pipe_busy.py gpurun_out/prof_commit_r2_v2.ncu-rep profiles/pipe_busy_r2_v2.json"""
import csv
import json
import subprocess
import sys

rep, dst = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
WANT = {
    "alu_busy_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "fmaheavy_busy_pct": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram_read_pct": "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
}
agg = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("void ", "").split("(")[0].split("<")[0]
    grid = r[col["launch__grid_size"]] if "launch__grid_size" in col else "0"
    if float(grid.replace(",", "")) < 64:  # single-CTA tails say nothing about pipe limits
        continue
    d = agg.setdefault(name, {"launches": 0, **{k: 0.0 for k in WANT}})
    d["launches"] += 1
    for k, m in WANT.items():
        d[k] += float(r[col[m]].replace(",", "")) if m in col else float("nan")
res = {"source": rep.split("/")[-1], "what": "ncu --set full, Pcs::commit of 2^20 x 14 (tools/profile_commit.py 20 14 1); averages over launches",
       "kernels": {k: {"launches": d["launches"], **{m: round(d[m] / d["launches"], 1) for m in WANT}} for k, d in agg.items()}}
json.dump(res, open(dst, "w"), indent=1)
print(json.dumps(res["kernels"], indent=1))
