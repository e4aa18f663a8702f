"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`): one line per kernel launch with the metrics that
matter for this integer / HBM bound code.  usage: python tools/ncu_summary.py file.ncu-rep [max_rows]"""
import csv
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"), ("gpu__time_duration.sum", "us"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy_busy%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "st_long"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "st_short"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "st_math"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "st_wait"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "st_bar"),
    ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "st_notsel"),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "st_lg"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "st_mio"),
]


def main():
    rep = sys.argv[1]
    limit = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(k), n) for k, n in WANT if k in hdr]
    for r in rows[2:2 + limit]:
        parts = []
        for i, n in idx:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = ("%.3g" % f) if abs(f) < 1e6 else ("%.4g" % f)
            except ValueError:
                v = v[:40]
            parts.append("%s=%s%s" % (n, v, units[i] if n in ("us", "dram_rd", "dram_wr") else ""))
        print("  ".join(parts))


if __name__ == "__main__":
    main()
