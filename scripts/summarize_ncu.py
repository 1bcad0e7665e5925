"""Turn gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep into text summaries under profiles/.
usage: python scripts/summarize_ncu.py <tag> "<description of the captured kernel>" """
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
desc = sys.argv[2] if len(sys.argv) > 2 else ""
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

path = os.path.join(G, "launches_%s.csv" % tag)
if os.path.exists(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, n = collections.OrderedDict(), 0
    for row in csv.DictReader(lines):
        v, unit = float(row["Metric Value"]), row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        a = agg.setdefault(row["Kernel Name"][:110], [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    mine = sum(a[1] for k, a in agg.items() if ("hps::" in k or "tc::" in k))
    out = ["# ncu launch list (%s): ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised" % tag,
           "# command: python bench.py --steps 3 --warmup 3 (ncu -s skip -c count window; includes the untimed L2-flush fills)",
           "# launches %d, total %.1f us; hps:: kernels %.1f us = %.1f%% of the window (rest = PyTorch glue)" % (n, tot, mine, 100 * mine / max(tot, 1e-9)),
           "# count   total_us   share   kernel"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%6d %10.1f %6.1f%%  %s" % (a[0], a[1], 100 * a[1] / tot, k))
    open(os.path.join(P, "%s_launches.txt" % tag), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:16]))

rep = os.path.join(G, "prof_%s.ncu-rep" % tag)
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h, units = r[0], r[1]
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
            "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
    o = ["# ncu --set full --clock-control none --import-source on (%s): %s" % (tag, desc)]
    for v in r[2:]:
        o.append("# ---- launch ----")
        for k in keys:
            for i, name in enumerate(h):
                if name == k:
                    o.append("%-85s %s %s" % (name, v[i], units[i]))
    open(os.path.join(P, "%s_full.txt" % tag), "w").write("\n".join(o) + "\n")
    print("\n".join(o))
