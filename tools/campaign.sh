#!/bin/bash
# Round measurement campaign on one B200: tests, bench (both arms), ncu launch lists and full
# captures of the dominant kernels.  Outputs under gpurun_out/campaign/.
O=gpurun_out/campaign; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest_gpu.txt
timeout 400 python bench.py > $O/bench.json 2> $O/bench.err
timeout 400 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
NCU="ncu --clock-control none --cache-control none"
timeout 150 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/klt_launches.csv python tools/profile_klt.py 4 > /dev/null 2>&1
timeout 150 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/ba_launches.csv python tools/profile_ba.py 2 > /dev/null 2>&1
timeout 150 $NCU --set full --import-source on -k regex:klt_gain_fused -s 3 -c 1 -o $O/klt_gain_fused python tools/profile_klt.py 6 > /dev/null 2>&1
timeout 150 $NCU --set full --import-source on -k regex:klt_front -s 3 -c 1 -o $O/klt_front python tools/profile_klt.py 6 > /dev/null 2>&1
timeout 150 $NCU --set full --import-source on -k regex:ba_schur_pairs -s 2 -c 1 -o $O/ba_schur_pairs python tools/profile_ba.py 2 > /dev/null 2>&1
timeout 150 $NCU --set full --import-source on -k regex:ba_chol_potf2_inv -s 40 -c 1 -o $O/ba_chol_potf2_inv python tools/profile_ba.py 2 > /dev/null 2>&1
cat $O/pytest_gpu.txt; head -c 600 $O/bench.json; echo; head -c 400 $O/bench_reference.json; echo; ls -la $O
