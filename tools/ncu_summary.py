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
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
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


HBM_PEAK_GBS = 6546.6          # MEASURED_PEAKS.json (copy bandwidth, this pool's B200)
FP64_ALU_DFMA_PER_S = 148 * 64 * 1.965e9     # non-tensor fp64 pipe: 64 DFMA / clk / SM


def _num(x):
    try:
        return float(x.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def report(src, dst):
    """One LABELLED block per kernel (mean over its launches in the capture), followed by derived roofline
    fractions: DRAM bytes / duration against the measured HBM copy peak, and the fp64 / DMMA pipe utilisation
    ncu reports (= fraction of the fp64-ALU / FP64 tensor roof for that instruction mix)."""
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    groups = collections.OrderedDict()
    for r in data:
        groups.setdefault(re.sub(r"\(.*", "", r[name_col]).strip(), []).append(r)
    col = {h: i for i, h in enumerate(hdr)}

    def mean(rs, metric):
        if metric not in col:
            return None
        vals = [_num(r[col[metric]]) for r in rs]
        vals = [v for v in vals if v is not None]
        return sum(vals) / len(vals) if vals else None

    def scaled(rs, metric, to):
        """Mean of `metric` converted to unit `to` (ncu picks per-column units)."""
        v = mean(rs, metric)
        if v is None:
            return None
        u = units[col[metric]]
        f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
             "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}.get(u, 1.0)
        return v * f / to
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on ; source report: %s\n" % src)
        f.write("# one block per kernel; every value is the mean over that kernel's launches in the capture\n")
        f.write("# peaks: HBM copy %.1f GB/s (MEASURED_PEAKS.json); fp64 ALU %.1f TDFMA/s (148 SM x 64 / clk x 1.965 GHz)\n"
                % (HBM_PEAK_GBS, FP64_ALU_DFMA_PER_S / 1e12))
        for name, rs in groups.items():
            f.write("\n== %s   (%d launch%s)\n" % (name[:110], len(rs), "" if len(rs) == 1 else "es"))
            for h in KEEP:
                if h in col:
                    v = mean(rs, h)
                    if v is not None:
                        f.write("   %-84s %-16s %.6g\n" % (h, units[col[h]], v))
            t = scaled(rs, "gpu__time_duration.sum", 1.0)
            rd, wr = scaled(rs, "dram__bytes_read.sum", 1.0), scaled(rs, "dram__bytes_write.sum", 1.0)
            if t and rd is not None and wr is not None:
                gbs = (rd + wr) / t / 1e9
                f.write("   -> derived: DRAM read+write %.1f MB in %.1f us = %.0f GB/s = %.3f of the HBM copy peak\n"
                        % ((rd + wr) / 1e6, t * 1e6, gbs, gbs / HBM_PEAK_GBS))
            p64 = mean(rs, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
            pd = mean(rs, "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active")
            if p64 is not None:
                f.write("   -> derived: fp64 ALU pipe %.1f %% of its peak (the fp64-ALU roof for this instruction mix)%s\n"
                        % (p64, "; DMMA sub-pipe %.1f %% (FP64 tensor roof)" % pd if pd else ""))


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
