#!/usr/bin/env python
"""Condense an `.ncu-rep` (captured with `ncu --set full`) into the per-kernel numbers we track.

    python tools/ncu_summary.py gpurun_out/step_full.ncu-rep > profiles/step_full_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__cluster_dim_x",
    "sm__cycles_active.avg",
    "smsp__inst_executed.sum",
    "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def main(path: str) -> None:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    total = 0.0
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        print("----")
        print(f"  {'Kernel Name':74s} {r[col['Kernel Name']][:64]}")
        for k in ("Grid Size", "Block Size"):
            if k in col:
                print(f"  {k:74s} {r[col[k]]}")
        for k in KEEP:
            if k in col:
                print(f"  {k:74s} {r[col[k]]} {units[col[k]]}")
        try:
            v, u = float(r[col["gpu__time_duration.sum"]].replace(",", "")), units[col["gpu__time_duration.sum"]]
            total += v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0) if u != "us" else v
        except (KeyError, ValueError):
            pass
    print("----")
    print(f"total gpu__time_duration over the listed launches: {total:.1f} us (serialised, cold caches: compare shares)")


if __name__ == "__main__":
    main(sys.argv[1])
