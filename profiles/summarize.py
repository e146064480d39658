#!/usr/bin/env python
"""Turn ncu reports / launch lists brought back in gpurun_out/ into the small text summaries kept under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv > profiles/r01_launches.md
  python profiles/summarize.py kernel gpurun_out/prof.ncu-rep  > profiles/r01_kernel.md
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "sm__cycles_elapsed.avg",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        d[r[ki]].append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in d.values())
    print("| kernel | launches | avg us | total us | share |\n|---|---|---|---|---|")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k[:110]}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / 1e3:.1f} | {100 * sum(v) / total:.1f}% |")


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    name_i = h.index("Kernel Name")
    for r in rows[2:]:
        print(f"### `{r[name_i][:140]}`\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in h:
                print(f"| {k} | {r[h.index(k)]} | {rows[1][h.index(k)]} |")
        stalls = [(float(r[i]), n) for i, n in enumerate(h) if n.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in n and r[i]]
        tot = sum(s for s, _ in stalls) or 1
        print("\nwarp-state samples: " + ", ".join(f"{n.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * s / tot:.0f}%" for s, n in sorted(stalls, reverse=True)[:7]) + "\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
