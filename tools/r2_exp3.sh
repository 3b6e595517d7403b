#!/bin/bash
O=gpurun_out/r2c; mkdir -p $O
./build/peaks > $O/peaks.json 2>&1; cat $O/peaks.json
timeout 600 python -m pytest tests/test_gpu_pose_ba.py -x -q 2>&1 | tail -15 | tee $O/pytest_ba.txt
timeout 400 python tools/r2_ba_exp.py 2,3 2>&1 | tee $O/ba_exp.txt
COSL_BA_SCHUR_PAIRS=1 timeout 400 python tools/r2_ba_exp.py 3 2>&1 | tee $O/ba_exp_pairs.txt
COSL_BA_SPECIAL=8 timeout 400 python tools/r2_ba_exp.py 3 2>&1 | tee $O/ba_exp_sp8.txt
COSL_BA_SPECIAL=40 timeout 400 python tools/r2_ba_exp.py 3 2>&1 | tee $O/ba_exp_sp40.txt
timeout 300 python tools/ba_trace.py c4 2>&1 | tee $O/ba_trace.txt
