#!/bin/bash
# ncu metric pass over the resize kernels of one mip chain (writes gpurun_out/resize_metrics.csv)
timeout 300 ncu --metrics gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:resize_pass -c 4 --csv --log-file gpurun_out/resize_metrics.csv \
  python bench.py --format BC1_RGB --size 4096 --mips --mipgen --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/resize_metrics.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
for r in rows[h + 1:]:
    print(r[0], r[4][24:48], r[8], r[-3], r[-1])
PY
