#!/usr/bin/env python
"""Per-kernel summary + top stall sites of an ncu report (needs -lineinfo and --import-source on).

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [kernel-regex] [top-n]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
cols = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "smsp__inst_executed.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
idx = [(c, h.index(c)) for c in cols if c in h]
seen = set()
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    if pat and pat not in name:
        continue
    key = name[:40]
    if key in seen:
        continue
    seen.add(key)
    print({c.split(".")[0][-34:]: r[i][:44] for c, i in idx})
if not pat:
    sys.exit(0)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
si, so = h.index("# Samples"), h.index("Source")
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data, agg = [], collections.Counter()
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) < len(h):
        continue
    try:
        n = int(r[si])
    except ValueError:
        continue
    d = {h[i][6:]: int(r[i]) for i in stall if r[i] not in ("0", "")}
    for k, v in d.items():
        agg[k] += v
    data.append((n, len(data), r[so].strip(), d))
tot = sum(d[0] for d in data) or 1
print("instructions", len(data), "samples", tot, "stalls", agg.most_common(8))
for n, i, s, d in sorted(data, reverse=True)[:topn]:
    print(f"{i:6d} {100 * n / tot:5.1f}%  {s[:84]:84s} {d}")
