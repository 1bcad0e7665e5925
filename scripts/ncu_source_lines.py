"""Per-source-line stall samples and executed warp instructions of one kernel, straight from an ncu report captured with
`--set full --import-source on` (no cubin needed: the report carries the source).  Runs anywhere ncu is installed.
usage: python scripts/ncu_source_lines.py <report.ncu-rep> <kernel-regex> [top=40] > profiles/<name>.txt"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kre],
                     capture_output=True, text=True).stdout
agg, path, fns, hdr = {}, "", set(), None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fns.add(r[1])
    elif r[0] == "Line No":
        hdr = r
    elif hdr is not None and len(r) == len(hdr) and r[0].isdigit():
        si, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        try:
            st = int(r[si]) if r[si] not in ("-", "") else 0
            ex = int(r[ie]) if r[ie] not in ("-", "") else 0
        except ValueError:
            continue
        a = agg.setdefault((path, int(r[0]), r[1].strip()), [0, 0])
        a[0] += st
        a[1] += ex
tot, toti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print("# %s  kernel regex %r" % (rep.split("/")[-1], kre))
for f in sorted(fns):
    print("#   " + f[:150])
print("# total: %d stall samples, %d warp instructions executed" % (tot, toti))
print("%6s %6s %9s %6s  %-24s source" % ("stall", "stall%", "instr", "instr%", "file:line"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6d %5.1f%% %9d %5.1f%%  %-24s %s" % (a[0], 100.0 * a[0] / max(tot, 1), a[1], 100.0 * a[1] / max(toti, 1),
                                              "%s:%d" % (key[0], key[1]), key[2][:100]))
