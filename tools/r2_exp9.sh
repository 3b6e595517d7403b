#!/bin/bash
O=gpurun_out/r2i; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pose_ba.py tests/test_golden.py -x -q 2>&1 | tail -5 | tee $O/pytest.txt
timeout 400 python tools/r2_ba_exp.py 3 2>&1 | tee $O/ba_exp_mma.txt
COSL_BA_SCHUR_SIMT=1 timeout 400 python tools/r2_ba_exp.py 3 2>&1 | tee $O/ba_exp_simt.txt
timeout 400 python tools/r2_shard_exp.py 2>&1 | tee $O/shard_exp.txt
COSL_BA_SCHUR_SIMT=1 timeout 400 python tools/r2_shard_exp.py 2>&1 | tee $O/shard_exp_simt.txt
timeout 300 python tools/r2_local_ba.py 2>&1 | tee $O/local_ba.txt
ncu --set full --clock-control none --import-source on -k regex:ba_schur_mma -s 3 -c 1 -o $O/ba_schur_mma python tools/profile_ba.py 2 > $O/ncu_mma.log 2>&1
