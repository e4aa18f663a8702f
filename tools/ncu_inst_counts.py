"""From an ncu CSV (--metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv) of tools/one_step.py, write profiles/step_counters_<tag>.json: per-kernel and per-stage thread instructions,
DRAM traffic and (cold, serialised) time per bench step. bench.py uses the instruction counts -- they do not depend on
the data -- to turn its live CUDA-event times into an achieved integer-instruction rate."""
import csv
import json
import sys

src, dst, steps = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 2
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
iname, imetric, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    k = r[iname].split("(")[0].replace("void ", "")
    k = k.split("<")[0]
    v = float(r[ival].replace(",", ""))
    u = r[iunit]
    if r[imetric].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if r[imetric].startswith("gpu__time"):
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    d = agg.setdefault(k, {"launches": 0, "warp_inst": 0.0, "us": 0.0, "dram_bytes": 0.0})
    if r[imetric] == "smsp__inst_executed.sum":
        d["warp_inst"] += v
        d["launches"] += 1
    elif r[imetric].startswith("gpu__time"):
        d["us"] += v
    else:
        d["dram_bytes"] += v
stage_of = {"k_ntt_strided": "lde", "k_ntt_block": "lde", "k_lde_mid": "lde", "k_hash_rows_staged": "merkle", "k_hash_rows_direct": "merkle",
            "k_merkle_subtree": "merkle"}
out = {"steps_captured": steps, "kernels": {}, "stages": {}}
for k, d in agg.items():
    e = {"launches_per_step": d["launches"] / steps, "thread_inst_per_step": 32 * d["warp_inst"] / steps,
         "ncu_us_per_step": d["us"] / steps, "dram_bytes_per_step": d["dram_bytes"] / steps}
    out["kernels"][k] = e
    st = stage_of.get(k)
    if st:
        s = out["stages"].setdefault(st, {"thread_inst_per_step": 0.0, "ncu_us_per_step": 0.0, "dram_bytes_per_step": 0.0})
        for f in s:
            s[f] += e[f]
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out["stages"]))
