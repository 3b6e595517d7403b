#!/bin/bash
O=gpurun_out/r2w; mkdir -p $O
for K in ba_tile_two ba_finish_cost ba_prep_kernel ba_back_cams_points; do
timeout 300 ncu --set full --import-source on --clock-control none -k regex:$K -s 6 -c 1 -o $O/$K python tools/r2_local_one.py c3 6 > $O/$K.log 2>&1
python tools/ncu_summary.py $O/$K.ncu-rep > $O/$K.summary.txt 2>/dev/null; head -6 $O/$K.summary.txt | tail -3
done
