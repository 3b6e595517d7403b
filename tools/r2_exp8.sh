#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pose_ba.py tests/test_golden.py -x -q 2>&1 | tail -5 | tee $O/pytest.txt
timeout 400 python tools/r2_ba_exp.py 3 2>&1 | tee $O/ba_exp.txt
COSL_BA_NO_HOT=1 timeout 400 python tools/r2_ba_exp.py 3 2>&1 | tee $O/ba_exp_nohot.txt
COSL_BA_SCHUR_MINB=6 timeout 400 python tools/r2_ba_exp.py 3 2>&1 | tee $O/ba_exp_minb6.txt
COSL_BA_SCHUR_MINB=8 timeout 400 python tools/r2_ba_exp.py 3 2>&1 | tee $O/ba_exp_minb8.txt
timeout 300 python tools/ba_trace.py c4 2>&1 | tee $O/ba_trace.txt
