#!/bin/bash
O=gpurun_out/r2b; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pose_ba.py tests/test_pose_ref.py tests/test_gpu_ref_facade.py -x -q 2>&1 | tail -15 | tee $O/pytest_ba.txt
timeout 600 python -m pytest tests/test_gpu_klt.py -x -q -k "benchmarked" -s 2>&1 | tail -15 | tee $O/pytest_klt_cfg.txt
timeout 400 python tools/r2_ba_exp.py 2,3 2>&1 | tee $O/ba_exp.txt
COSL_BA_NO_SCRATCH=1 timeout 400 python tools/r2_ba_exp.py 3 2>&1 | tee $O/ba_exp_noscratch.txt
timeout 300 python tools/ba_trace.py c4 2>&1 | tee $O/ba_trace.txt
