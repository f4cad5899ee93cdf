#!/usr/bin/env python
"""Render the JSON of tools/bench_ps.py as the multi-writer PS bandwidth table (markdown).

    python tools/ps_table.py gpurun_out/bench_ps_w8.json > profiles/r2/ps_bandwidth_8gpu.md
"""
import json
import sys


def main(path):
    d = json.load(open(path))
    rows = d["results"]
    ops, sizes, writers = [], [], []
    for r in rows:
        if r["op"] not in ops:
            ops.append(r["op"])
        if r["n"] not in sizes:
            sizes.append(r["n"])
        if r["writers"] not in writers:
            writers.append(r["writers"])
    writers.sort()
    print(f"# Parameter-server push / pull over NVLink, {d['world']} GPUs (rank 0 owns the center, ranks 1..w write at once)\n")
    print("Device-timed (CUDA events, max over the active ranks), 20 iterations after 3 warm-ups (5 at 100 M). Cell: "
          "`time per op | GB/s per writer per direction | GB/s total into/out of the PS GPU`. One direction moves "
          "4N bytes; `exchange` / `elastic` / `strict` move 4N each way.\n")
    for n in sizes:
        print(f"## N = {n:,} fp32 ({4 * n / 2**20:.0f} MiB)\n")
        print("| op | " + " | ".join(f"{w} writer{'s' if w > 1 else ''}" for w in writers) + " |")
        print("|---|" + "---|" * len(writers))
        for op in ops:
            cells = []
            for w in writers:
                rr = [r for r in rows if r["n"] == n and r["op"] == op and r["writers"] == w]
                cells.append(f"{rr[0]['us']:.1f} us \\| {rr[0]['GBps_per_writer_per_dir']:.0f} \\| "
                             f"{rr[0]['GBps_ps_ingress_total']:.0f}" if rr else "")
            print(f"| `{op}` | " + " | ".join(cells) + " |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
