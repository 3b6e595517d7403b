#!/bin/bash
O=gpurun_out/r2n; mkdir -p $O
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_pose_ba.py -x -q -k "local_parity_c2 or intercam" 2>&1 | tail -12 | tee $O/sanitizer.txt
timeout 900 python -m pytest tests/test_gpu_pose_ba.py -x -q 2>&1 | tail -8 | tee $O/pytest.txt
timeout 400 python tools/r2_ba_exp.py -1 2>&1 | tee $O/ba_exp_blk.txt
COSL_BA_SCHUR_PAIRS=1 timeout 400 python tools/r2_ba_exp.py -1 2>&1 | tee $O/ba_exp_st.txt
python tools/r2_local_ba.py 2>&1 | grep "{}" | tee $O/local.txt
