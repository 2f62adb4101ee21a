#!/bin/bash
# Scratch GPU session: parity tests, one sweep line (overlapped + serialised stage times), ncu counters of the point kernels.
#   bash tools/quick.sh <tag> [kernel-regex]
TAG=${1:-q}; KR=${2:-k_label_bev|k_quad_reduce}
mkdir -p gpurun_out/$TAG
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/sweep.py --frames 2048 --chunks 1024 --streams 2 --reps 4 2>&1 | tee gpurun_out/$TAG/sweep.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['kfps'], 'kfps', d['ms_best'], 'ms'); print(' overlapped', d['stages']); print(' serial    ', d['serial'])
"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k "regex:$KR" -s 4 -c 4 --csv --log-file gpurun_out/$TAG/ncu_counters.csv python tools/sweep.py --frames 1024 --chunks 512 --streams 1 --reps 1 --warm 1 > /dev/null 2>&1
python - <<P
import csv
rows=[r for r in csv.reader(open('gpurun_out/$TAG/ncu_counters.csv')) if len(r)>10]
h=rows[0]; ix={n:i for i,n in enumerate(h)}
for r in rows[1:]:
    print(r[ix['Kernel Name']][:16], r[ix['Metric Name']], r[ix['Metric Value']])
P
