#!/bin/bash
# instruction-fetch stall check of the encoder kernels (writes gpurun_out/stall_<fmt>.csv)
for f in "$@"; do
  kind="noise+grad"; [ "$f" = "BC6H" ] && kind="hdr"
  timeout 200 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -s 1 -c 1 --csv --log-file gpurun_out/stall_$f.csv python tools/prof_one.py $f 2048 $kind > /dev/null 2>&1
  python - "$f" <<'PY'
import csv, sys
rows = list(csv.reader(open("gpurun_out/stall_%s.csv" % sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
print(sys.argv[1], rows[h + 1][4][:40], " ".join("%s=%s" % (r[-3].split("__")[-1][:34], r[-1]) for r in rows[h + 1:]))
PY
done
