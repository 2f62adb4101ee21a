#!/bin/bash
# Scratch: run tools/sweep.py under a list of "ENV=.. ENV=.." settings (one per line on stdin), print kfps per setting.
#   echo "SSD_GPU_SPLIT=1 SSD_GPU_PRIO=1" | bash tools/knobs.sh [sweep args]
while read -r line; do
  [ -z "$line" ] && continue
  out=$(env $line python tools/sweep.py --frames 2048 --reps 4 "$@" 2>&1 | tail -n +1)
  echo "$out" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print('  ?', l[:160].rstrip()); continue
    print('%-70s s=%d c=%d  %.3f ms  %.1f kfps  %s' % ('''$line''', d['streams'], d['chunk'], d['ms_best'], d['kfps'], {k: round(v, 2) for k, v in d['serial'].items() if k in ('transform_bin', 'label_bev', 'outline', 'quad_reduce')}))
"
done
