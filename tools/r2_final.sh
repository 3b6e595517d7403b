#!/bin/bash
# round-end verification on one B200: GPU tests, smoke, both bench arms, ncu launch list of the bench
O=gpurun_out/r2z; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3 | tee $O/smoke.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
timeout 900 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -2 $O/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/bench_launches.csv python bench.py --steps 3 --warmup 3 --quick --no-cpu > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2z/bench.json"))
print("klt", d["value"], d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("pipelined",{}).get("ms_per_step"))
print("ba", d["ba"]["value"], d["ba"]["ms_per_trial"], d["ba"]["roofline"]["ms_per_trial_by_class"])
print("c3", d["c3_local_ba"]["us_per_trial"], "pipe", d["pipeline"]["value"], "c5", d["c5"]["fps_8cam_e2e"])
r=json.load(open("gpurun_out/r2z/bench_reference.json")); print("ref", r["value"], r.get("ba",{}).get("value"))
PY
