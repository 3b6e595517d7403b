#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_pose_ba.py -q -k "sba_signature" 2>&1 | grep -E "assert|passed|failed|rg|ro" | head -8; done | tee $O/sba.txt
timeout 600 python -m pytest tests/test_gpu_pose_ba.py tests/test_golden.py tests/test_shim.py -q 2>&1 | tail -5 | tee $O/pytest.txt
