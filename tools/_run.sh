cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02i
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_label_bev|k_quad_reduce" -s 4 -c 2 -o gpurun_out/r02i/prof python tools/sweep.py --frames 1024 --chunks 512 --streams 1 --reps 1 --warm 1 > gpurun_out/r02i/ncu.log 2>&1; tail -2 gpurun_out/r02i/ncu.log | cut -c1-200
