#!/bin/bash
O=gpurun_out/r2n; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pose_ba.py -x -q 2>&1 | tail -4 | tee $O/pytest.txt
COSL_BA_TIMING=1 timeout 400 python tools/r2_ba_exp.py -1 2>&1 | grep -v "^\[ba timing\] [a-z ]*[0-9.]* ms$" | grep "visits\|depth" | tail -3 | tee $O/ba_exp_blk.txt
python tools/r2_local_ba.py 2>&1 | grep "{}" | tee $O/local.txt
