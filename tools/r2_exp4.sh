#!/bin/bash
O=gpurun_out/r2d; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pose_ba.py -x -q 2>&1 | tail -15 | tee $O/pytest_ba.txt
timeout 400 python tools/r2_ba_exp.py 3 2>&1 | tee $O/ba_exp.txt
timeout 300 python tools/ba_trace.py c4 2>&1 | tee $O/ba_trace.txt
NCU="ncu --clock-control none"
timeout 300 $NCU --set full --import-source on -k regex:ba_tile_solve -s 3 -c 1 -o $O/ba_tile_solve python tools/profile_ba.py 2 > $O/ncu_tile.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:ba_schur_rows -s 3 -c 1 -o $O/ba_schur_rows python tools/profile_ba.py 2 > $O/ncu_rows.log 2>&1
tail -3 $O/ncu_tile.log $O/ncu_rows.log; ls -la $O
