"""SASS evidence for the Blackwell path: per kernel of libhps_b200.so, the number of tcgen05 MMA (UTCHMMA / UTCQMMA ...),
TMEM load / store (LDTM / STTM), bulk-copy = TMA (UBLKCP), mbarrier (SYNCS), global reduction (REDG.E.ADD.F32; the four-wide form red.global.add.v4.f32 prints as `REDG.E.ADD.F32 [..], R..` with a 128-bit register group) and warp-reduce
(REDUX) instructions, from `cuobjdump -sass` (no GPU needed; B200_PROFILING.md names the mnemonics).
usage: python scripts/sass_digest.py > profiles/sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hierarchical-point-slam_b200", "libhps_b200.so")
PATTERNS = [("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"),
            ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"), ("UBLKCP (cp.async.bulk)", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"),
            ("REDG (red.global.add incl. .v4.f32)", r"\bREDG\.E\.ADD\.F32"), ("REDUX", r"\bREDUX"), ("HMMA (mma.sync)", r"\bHMMA"),
            ("MUFU", r"\bMUFU"), ("instructions", r"^\s+/\*[0-9a-f]{4}\*/")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    counts, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        for label, pat in PATTERNS:
            if re.search(pat, line):
                counts[name][label] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(counts), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    print("# cuobjdump -sass hierarchical-point-slam_b200/libhps_b200.so   (arch: %s; %d kernels)" % (", ".join(arch), len(counts)))
    print("# per kernel: occurrences of the Blackwell-specific instructions in the SASS text (static counts, not executions)")
    cols = [l for l, _ in PATTERNS]
    print("%-58s " % "kernel" + " ".join("%9s" % c.split(" ")[0][:9] for c in cols))
    tot = collections.Counter()
    for (mangled, c), nice in zip(counts.items(), demangled):
        short = re.sub(r"\(.*", "", nice).replace("void ", "").replace("hps::", "")
        print("%-58s " % short[:58] + " ".join("%9d" % c[l] for l in cols))
        tot.update(c)
    print("%-58s " % "TOTAL" + " ".join("%9d" % tot[l] for l in cols))
    print("# legend: " + "; ".join(cols))
    if not tot["UTC*MMA (tcgen05.mma)"] or not tot["LDTM (tcgen05.ld)"] or not tot["UBLKCP (cp.async.bulk)"]:
        sys.exit("expected tcgen05 / TMA instructions are missing from the SASS")


if __name__ == "__main__":
    main()
