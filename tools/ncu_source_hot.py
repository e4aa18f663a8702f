"""Per-instruction hot spots of one kernel from an .ncu-rep (needs --import-source on / -lineinfo):
python tools/ncu_source_hot.py file.ncu-rep <kernel regex> [instance]"""
import collections
import csv
import re
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
inst = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
sec = sections[inst]
ix = {h: i for i, h in enumerate(sec["hdr"])}
data = sec["data"]
val = lambda r, c: int(r[ix[c]] or 0)
print(sec["name"], "instructions:", len(data), "samples:", sum(val(r, "# Samples") for r in data))
stalls = [c for c in sec["hdr"] if c.startswith("stall_") and "Not Issued" not in c]
tot = {c: sum(val(r, c) for r in data) for c in stalls}
print("stall samples:", {k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
for col in sorted(tot, key=lambda c: -tot[c])[:3]:
    print("--- top", col)
    for r in sorted(data, key=lambda r: -val(r, col))[:10]:
        print("  %6d  %s" % (val(r, col), r[ix["Source"]].strip()[:100]))
agg, cnt = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    if m:
        agg[m.group(1)] += val(r, "# Samples")
        cnt[m.group(1)] += val(r, "Instructions Executed")
print("executed warp-instructions by opcode:", cnt.most_common(14))
print("samples by opcode:", agg.most_common(10))
