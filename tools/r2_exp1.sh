#!/bin/bash
# first GPU contact of round 2: BA parity tests with the tile solver, measured pipe peaks, timers
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
./build/peaks > $O/peaks.json 2>&1; cat $O/peaks.json
timeout 600 python -m pytest tests/test_gpu_pose_ba.py -x -q 2>&1 | tail -15 | tee $O/pytest_ba.txt
timeout 400 python tools/r2_ba_exp.py 0,1,2,3 2>&1 | tee $O/ba_exp.txt
timeout 300 python tools/ba_trace.py c4 2>&1 | tee $O/ba_trace.txt
