#!/usr/bin/env python
"""Summarises ncu output into small text files for profiles/ (the .ncu-rep files stay in gpurun_out/).

  python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.md
  python tools/ncu_summary.py full gpurun_out/prof_x.ncu-rep [...]        > profiles/rNN_full.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % (int/DPX min-add live here)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe % (POPC lives here)"),
    ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("sm__cycles_elapsed.avg", "SM elapsed cycles (avg)"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe cycles active %"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp cycles per issued instruction"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = next(i for i, r in enumerate(rows) if r[0] == "ID")
    head, rows = rows[h], rows[h + 1:]
    ki, vi, gi, bi = head.index("Kernel Name"), head.index("Metric Value"), head.index("Grid Size"), head.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        agg.setdefault((name, r[gi], r[bi]), []).append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    print("| kernel | grid | block | launches | avg us | share of GPU time |")
    print("|---|---|---|---:|---:|---:|")
    for (name, g, b), v in agg.items():
        print(f"| `{name}` | {g} | {b} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / total:.1f}% |")
    print(f"\ntotal kernel time in the capture: {total / 1e6:.3f} ms over {sum(len(v) for v in agg.values())} launches "
          "(ncu serialises launches and runs them cold: compare shares, not absolutes)")


def full(paths):
    for p in paths:
        # (a .csv argument is the `ncu -i X.ncu-rep --page raw --csv` page made on the GPU box: full-set reports of batched
        #  workloads are too large to travel back)
        out = open(p).read() if p.endswith(".csv") else subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        head, units, rows = rows[0], rows[1], rows[2:]
        print(f"### {p}\n")
        for r in rows:
            print(f"**`{r[head.index('Kernel Name')]}`**\n")
            for k, label in KEYS:
                if k in head:
                    i = head.index(k)
                    print(f"- {label}: {r[i]} {units[i]}  (`{k}`)")
            stalls = [(float(r[i] or 0), head[i]) for i in range(len(head)) if head[i].startswith("smsp__average_warps_issue_stalled_") and head[i].endswith("_per_issue_active.ratio")]
            stalls.sort(reverse=True)
            if stalls:
                print("- top stall reasons (warps stalled per issue-active cycle): " + ", ".join(f"{n.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, n in stalls[:5]))
            print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
