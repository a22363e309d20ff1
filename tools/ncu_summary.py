#!/usr/bin/env python
"""Condense an ncu report (.ncu-rep) into the few numbers DESIGN.md and bench.py quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt

One block per captured launch: duration, DRAM bytes read/written (the `traffic` of bench.py's roofline),
executed warp instructions, issue-slot, ALU-pipe and FMA-pipe utilisation, occupancy, registers, grid,
shared-memory bank conflicts and the stall reasons above 0.1 warps per issue.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe (INT32/logic)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (of ncu's peak)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: {len(rows) - 2} launches (ncu --set full --clock-control none; times are cold-cache, serialised)")
    for r in rows[2:]:
        print(f"\n== {r[col['Kernel Name']]}")
        for key, label in KEYS:
            if key in col:
                print(f"  {label:34s} {r[col[key]]} {units[col[key]]}   [{key}]")
        stalls = []
        for h, i in col.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.1:
                    stalls.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        print("  stalls (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)))


if __name__ == "__main__":
    main()
