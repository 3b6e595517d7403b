#!/bin/bash
N=$(nvidia-smi -L | wc -l); O=gpurun_out/r2mg$N; mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "rc=$?"; tail -2 $O/bench_n$N.err
python - <<PY
import json
d=json.load(open("$O/bench_n$N.json"))
b=d["ba"]
print("N=$N klt", d["value"], "ms", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"])
print("ba it/s", b["value"], "ms/trial", b["ms_per_trial"], "parity", b.get("parity_rel_cost_diff_3_trials"))
print(b["roofline"]["ms_per_trial_by_class"])
print("c5", d.get("c5",{}).get("fps_8cam_e2e"), d.get("c5",{}).get("config"))
PY
