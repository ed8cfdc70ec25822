#!/usr/bin/env python
"""Turn an ncu report into the text summaries kept under profiles/.

  python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_zonal   [--launches gpurun_out/launches.csv]
writes <prefix>_metrics.csv (the raw-page metrics the roofline and stall analysis use), <prefix>_phases.txt
(warp instructions and stall samples per source line group of rs_zonal.cu) and <prefix>_launches.txt.
Needs the ncu CLI (no GPU) and the kernel built with -lineinfo.
"""
import collections
import csv
import io
import os
import subprocess
import sys

KEEP = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_op_shared_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.max", "sm__cycles_active.min", "sm__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
)


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    with open(prefix + "_metrics.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "unit", "value"])
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
            for h, u, v in zip(hdr, units, vals):
                if h in KEEP or h.startswith("smsp__average_warps_issue_stalled"):
                    w.writerow([name[:60], h, u, v])
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    h = src[2]
    ci, cs = h.index("Instructions Executed"), h.index("# Samples")
    cur, agg, text = None, collections.OrderedDict(), {}
    for r in src[3:]:
        if len(r) <= ci:
            continue
        if r[0].strip():
            try:
                cur = int(r[0])
                text[cur] = ",".join(r[1:4])[:110]
            except ValueError:
                pass
            continue
        try:
            a = agg.setdefault(cur, [0, 0])
            a[0] += int(r[ci]); a[1] += int(r[cs])
        except (ValueError, TypeError):
            pass
    ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    with open(prefix + "_phases.txt", "w") as f:
        f.write(f"# {os.path.basename(rep)}: {ti / 1e9:.3f} G warp instructions, {ts} stall samples; source lines with >= 0.5 % of either\n")
        f.write("# line  inst%  samples%  source (rs_zonal.cu unless an intrinsic header)\n")
        for ln, a in sorted(agg.items()):
            if a[0] >= 0.005 * ti or a[1] >= 0.005 * ts:
                f.write(f"{ln:5d}  {a[0] / ti * 100:5.1f}  {a[1] / ts * 100:5.1f}  {text.get(ln, '')}\n")
    if "--launches" in sys.argv:
        lp = sys.argv[sys.argv.index("--launches") + 1]
        lines = [l for l in open(lp) if not l.startswith("==")]
        rows = list(csv.reader(lines))
        hh = rows[0]
        ki, vi, ui = hh.index("Kernel Name"), hh.index("Metric Value"), hh.index("Metric Unit")
        ag = collections.OrderedDict()
        for r in rows[1:]:
            if len(r) <= vi:
                continue
            v = float(r[vi].replace(",", "")) / {"ns": 1e6, "us": 1e3, "ms": 1.0}.get(r[ui], 1.0)
            a = ag.setdefault(r[ki][:90], [0, 0.0]); a[0] += 1; a[1] += v
        tot = sum(a[1] for a in ag.values())
        with open(prefix + "_launches.txt", "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none over the bench command (cold-cache, serialised)\n")
            for k, a in ag.items():
                f.write(f"{a[0]:4d} launches {a[1]:10.3f} ms {a[1] / tot * 100:5.1f} %  {a[1] / a[0]:9.4f} ms/launch  {k}\n")


if __name__ == "__main__":
    main()
