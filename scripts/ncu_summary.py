#!/usr/bin/env python
"""Digest ncu artefacts into the small text files kept under profiles/.

  ncu_summary.py rep  <file.ncu-rep> [...]   key metrics of every captured launch (from `--set full`)
  ncu_summary.py list <launches.csv>         per-kernel count / total / share from a
                                             `--metrics gpu__time_duration.sum` launch list
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def rep(path):
    # a .ncu-rep, or the `ncu -i rep --page raw --csv` export of one (taken on the GPU box: the reports are ~12 MB each)
    out = open(path).read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print("%s: no launches" % path)
        return
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("## %s" % path.split("/")[-1])
    for r in rows[2:]:
        print("kernel: %s" % r[idx["Kernel Name"]][:110])
        for k in KEYS:
            if k in idx:
                print("  %-84s %s %s" % (k, r[idx[k]], units[idx[k]]))
        if "dram__bytes_read.sum" in idx:
            def mb(v, u):
                v = float(v.replace(",", ""))
                return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            tr = mb(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + mb(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            print("  %-84s %.3f MB" % ("dram traffic (read+write)", tr))
    print()


def launch_list(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        ns = float(r[-1].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    print("## %s  (%d launches, %.1f us total; per-launch times are cold-cache/serialised: compare SHARES)" % (path.split("/")[-1], len(rows), total / 1e3))
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-60s n=%-4d total=%9.1f us  avg=%8.2f us  share=%5.1f%%" % (k[:60], n, ns / 1e3, ns / 1e3 / n, 100 * ns / total))


if __name__ == "__main__":
    if len(sys.argv) < 3:
        sys.exit(__doc__)
    if sys.argv[1] == "rep":
        for p in sys.argv[2:]:
            rep(p)
    else:
        for p in sys.argv[2:]:
            launch_list(p)
