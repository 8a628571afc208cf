#!/bin/bash
# Turns the .ncu-rep files of scripts/gpu_profile_final.sh into the tracked summaries under profiles/ (run where ncu is installed; no GPU needed).
set -eu
TAG=${1:-r1_final}
cd "$(dirname "$0")/.."
KEEP='Kernel Name|Grid Size|Block Size|gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_tensor|sm__cycles_elapsed.avg|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__shared_mem_per_block_dynamic|smsp__inst_executed.sum|smsp__issue_active.avg.pct_of_peak_sustained_active|lts__throughput.avg.pct_of_peak_sustained_elapsed|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'
for n in k1 gemm; do
  ncu -i gpurun_out/final_$n.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys, re
rows = list(csv.reader(sys.stdin)); hdr = rows[0]; keep = re.compile(r'^($KEEP)')
idx = [i for i, h in enumerate(hdr) if keep.match(h) and 'per_second' not in h and '.max' not in h and '.min' not in h]
w = csv.writer(sys.stdout)
for r in rows: w.writerow([r[i] for i in idx])
" > profiles/${TAG}_ncu_$n.csv
done
cp gpurun_out/final_launches.csv profiles/${TAG}_launches.csv
python scripts/launch_summary.py profiles/${TAG}_launches.csv 7 > profiles/${TAG}_launches_summary.txt
ls -la profiles/${TAG}*
