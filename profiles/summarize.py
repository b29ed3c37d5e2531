"""Turns gpurun_out/*.ncu-rep and launch lists into the small text summaries committed under profiles/.

usage: python profiles/summarize.py <round-tag>   (reads gpurun_out/launches_<tag>.csv, prof_*_<tag>.ncu-rep)
"""
import collections
import csv
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fp64.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]


def launches(tag):
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}[r[ui]]
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(ROOT, "profiles", f"launches_{tag}_summary.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.2f} ms total\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{a[1] / tot * 100:6.2f}%  n={a[0]:5d}  avg={a[1] / a[0] / 1e3:10.1f} us  {k}\n")
    subprocess.call(["cp", path, os.path.join(ROOT, "profiles", f"launches_{tag}.csv")])


def reports(tag):
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_*_{tag}.ncu-rep"))):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            continue
        h, units = rows[0], rows[1]
        name = os.path.basename(rep).replace(".ncu-rep", "")
        with open(os.path.join(ROOT, "profiles", name + "_summary.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on; {len(rows) - 2} launch(es) of "
                    f"{rows[2][h.index('Kernel Name')][:80]}\n")
            for k in KEYS:
                if k in h:
                    i = h.index(k)
                    f.write(f"{k:75s} {units[i]:16s} " + "  ".join(r[i] for r in rows[2:]) + "\n")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    launches(tag)
    reports(tag)
