#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pose_ba.py -x -q 2>&1 | tail -15 | tee $O/pytest_ba.txt
timeout 400 python tools/r2_ba_exp.py 3 2>&1 | tee $O/ba_exp.txt
timeout 300 python tools/ba_trace.py c4 2>&1 | tee $O/ba_trace.txt
