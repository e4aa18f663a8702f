"""GPU idle gaps of one prove(): reads the `[timeline]` lines (MSGPU_TIMELINE=1) of tools/profile_prove.py from stdin and prints
every gap above a threshold between the end of one launch and the start of the next, with the launches on either side."""
import re
import sys

thr = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
rows = []
for line in sys.stdin:
    m = re.match(r"\[timeline\] (\S*)\s+(\S+)\s+host \+\s*([\d.]+) us\s+gpu \+\s*([\d.]+) us\s+dur\s+([\d.]+) us", line)
    if m:
        rows.append((m.group(1), m.group(2), float(m.group(3)), float(m.group(4)), float(m.group(5))))
# keep the last proof only: gpu offsets restart at 0 for every profile
start = max(i for i, r in enumerate(rows) if r[3] == 0.0)
rows = rows[start:]
busy = sum(r[4] for r in rows)
end = rows[-1][3] + rows[-1][4]
print("launches %d, span %.1f us, busy %.1f us, idle %.1f us" % (len(rows), end, busy, end - busy))
tot = 0.0
for a, b in zip(rows, rows[1:]):
    gap = b[3] - (a[3] + a[4])
    if gap > thr:
        tot += gap
        print("gap %8.1f us at +%9.1f us  after %-9s %-22s before %-9s %-22s" % (gap, a[3] + a[4], a[0], a[1], b[0], b[1]))
print("gaps above %.0f us: %.1f us" % (thr, tot))
