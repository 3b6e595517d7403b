#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
timeout 400 python tools/r2_ba_exp.py -1 2>&1 | tail -1 | tee $O/ba_exp.txt
python tools/r2_local_ba.py 2>&1 | grep "{}" | tee $O/local.txt
timeout 600 python -m pytest tests/test_gpu_pose_ba.py -x -q 2>&1 | tail -3
