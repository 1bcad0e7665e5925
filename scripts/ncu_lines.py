"""Attribute ncu warp-stall samples (SASS source page) to CUDA source lines via nvdisasm line info.
usage: python scripts/ncu_lines.py <report.ncu-rep> <kernel-regex> <cubin> <mangled-substring> [top]"""
import collections
import csv
import re
import subprocess
import sys

rep, kre, cubin, fn = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "demangled", "-k", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hdr[0]]
body = rows[hdr[0] + 1:(hdr[1] - 1 if len(hdr) > 1 else len(rows))]
si = h.index("Warp Stall Sampling (All Samples)")
ie = h.index("Instructions Executed")
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines, cur, infn, line = [], None, False, None
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        infn = fn in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(line)
print("sass instrs: ncu %d, nvdisasm %d" % (len(body), len(lines)))
agg = collections.Counter()
ex = collections.Counter()
for r, ln in zip(body, lines):
    agg[ln] += int(r[si])
    ex[ln] += int(r[ie])
tot = sum(agg.values())
src = {}
for (f, n), v in agg.most_common(top):
    if f not in src:
        import glob
        c = glob.glob("/root/repo/**/" + f, recursive=True)
        src[f] = open(c[0]).read().splitlines() if c else []
    text = src[f][n - 1].strip()[:100] if n - 1 < len(src[f]) else ""
    print("%5.1f%% %6d ex=%8d %s:%d  %s" % (100.0 * v / tot, v, ex[(f, n)], f, n, text))
