"""Opcode mix of one kernel from `cuobjdump -sass`: usage python tools/sass_mix.py <object> <substring of the mangled name>"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, mix = None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            mix[m.group(1).split(".")[0]] += 1
tot = sum(mix.values())
print(pat, "static instructions:", tot)
for op, n in mix.most_common(25):
    print("  %-10s %6d  %5.1f%%" % (op, n, 100.0 * n / tot))
