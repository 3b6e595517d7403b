#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_klt.py tests/test_gpu_klt_edges.py tests/test_golden.py tests/test_gpu_pose_ba.py -x -q 2>&1 | tail -5 | tee $O/pytest.txt
timeout 300 python tools/r2_local_ba.py 2>&1 | tee $O/local_ba.txt
timeout 600 python bench.py --quick --no-cpu > $O/bench_quick.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench_quick.json')); print('klt ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['share_of_step']); print('ba', d['ba']['ms_per_trial'])"
COSL_KLT_NO_TMA=1 timeout 600 python bench.py --quick --no-cpu > $O/bench_quick_notma.json 2> $O/bench2.err; python -c "
import json; d=json.load(open('$O/bench_quick_notma.json')); print('NO TMA klt ms', d['ms_per_step'], d['roofline']['share_of_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $O/klt_launches.csv python tools/profile_klt.py 12 > /dev/null 2>&1; python tools/launch_summary.py $O/klt_launches.csv 2>/dev/null | tail -14
