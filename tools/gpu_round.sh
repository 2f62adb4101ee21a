#!/bin/bash
# One GPU-box session: parity tests, bench, stage sweep, ncu launch list + full capture of the point kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [quick]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/nproc.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
if [ "$2" != "quick" ]; then
echo "== sweep"; timeout 600 python tools/sweep.py --frames 1024 --chunks 8,32,128,256 --reps 5 2>&1 | tee $OUT/sweep.jsonl
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 3 --frames 512 --e2e-frames 16 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_transform_bin|k_label_bev|k_quad_reduce|k_outline' -s 8 -c 4 \
   -o $OUT/prof python bench.py --steps 1 --warmup 3 --frames 512 --e2e-frames 16 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
fi
