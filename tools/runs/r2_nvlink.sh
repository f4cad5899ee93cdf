#!/bin/bash
mkdir -p gpurun_out
ncu --query-metrics 2>/dev/null | grep -i -E "nvl" | head -80 > gpurun_out/nvlink_metric_names.txt; wc -l gpurun_out/nvlink_metric_names.txt
timeout 300 python tools/ncu_nvlink_ps.py --numel 12000000 2>&1 | tee gpurun_out/ps_single_process_2gpu.txt
M=$(grep -o -E "nvl[rt]x__bytes[a-z_]*" gpurun_out/nvlink_metric_names.txt | sort -u | sed 's/$/.sum/' | paste -sd, -)
echo "metrics: $M"
timeout 600 ncu --set full --section Nvlink --section Nvlink_Tables --section Nvlink_Topology ${M:+--metrics $M} --clock-control none -k regex:"ps_" -c 8 -o gpurun_out/r2_ps_nvlink -f \
   python tools/ncu_nvlink_ps.py --numel 12000000 --iters 1 > gpurun_out/ncu_ps_nvlink.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_ps_nvlink.log
ls -la gpurun_out/r2_ps_nvlink.ncu-rep
