#!/bin/bash
O=gpurun_out/r2k; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pose_ba.py tests/test_golden.py tests/test_shim.py -x -q 2>&1 | tail -5 | tee $O/pytest.txt
timeout 300 python tools/ba_setup_timing.py 2>&1 | grep "==\|work lists\|build_solver\|uploads" | tee $O/setup_timing.txt
