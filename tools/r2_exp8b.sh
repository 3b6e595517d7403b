#!/bin/bash
O=gpurun_out/r2p; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pose_ba.py tests/test_ba_plan.py -x -q 2>&1 | tail -4 | tee $O/pytest.txt
timeout 400 python tools/r2_ba_exp.py -1 2>&1 | tee $O/ba_exp.txt
timeout 300 python tools/ba_trace.py c4 2>&1 | tee $O/ba_trace.txt | head -14
python tools/r2_local_ba.py 2>&1 | grep "{}" | tee $O/local.txt
