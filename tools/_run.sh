cd $GRAFT_REPO_ROOT
SSD_GPU_LIB=build/libssd_wrecdbg.so python tools/sweep.py --frames 1024 --chunks 512 --reps 2 --warm 1 2>&1 | tail -1
