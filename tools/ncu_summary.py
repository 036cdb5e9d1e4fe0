#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches_summary.txt
  python tools/ncu_summary.py report   gpurun_out/prof_score.ncu-rep profiles/r01_score_kernel_ncu.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "lts__t_bytes.sum", "sm__cycles_active.avg",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg, total = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row["Metric Unit"]
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).strip()
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold cache: compare SHARES)\n")
        f.write("# source: %s ; total device time %.1f us over %d launches\n" % (src, total, sum(a[0] for a in agg.values())))
        f.write("%-72s %7s %12s %11s %7s\n" % ("kernel", "n", "total_us", "avg_us", "share"))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s %7d %12.1f %11.1f %6.2f%%\n" % (k[:72], n, t, t / n, 100 * t / total))
        # the bench's "step" is one acquisition pass: K_* generation, the contraction kernel, its finish and merge
        step = {k: v for k, v in agg.items() if re.search(r"kst_kernel|score_kernel|score_finish_kernel|merge_best_kernel", k)}
        if step:
            st = sum(v[1] for v in step.values())
            f.write("\n# kernels of the scoring step only (bench.py roofline.share_of_step is score_kernel's share of these)\n")
            for k, (n, t) in sorted(step.items(), key=lambda kv: -kv[1][1]):
                f.write("%-72s %7d %12.1f %11.1f %6.2f%%\n" % (k[:72], n, t, t / n, 100 * t / st))


def report(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on ; source report: %s\n" % src)
        name_col = hdr.index("Kernel Name")
        f.write("# launches captured: %s\n" % ", ".join(sorted({r[name_col][:60] for r in data})))
        for i, h in enumerate(hdr):
            if h in KEEP:
                f.write("%-86s %-10s %s\n" % (h, units[i], "  ".join(r[i] for r in data)))


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
