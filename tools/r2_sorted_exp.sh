#!/bin/bash
N=$(nvidia-smi -L | wc -l); O=gpurun_out/r2s$N; mkdir -p $O
for S in 0 1; do
COSL_BENCH_BA_SORTED=$S timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$S bench.py --gpus $N --steps 20 --warmup 5 --quick --no-cpu > $O/bench_s$S.json 2> $O/bench_s$S.err
python - <<PY
import json
d=json.load(open("$O/bench_s$S.json")); b=d["ba"]
print("sorted=$S N=$N ba", round(b["value"],1), round(b["ms_per_trial"],4), "parity", b.get("parity_rel_cost_diff_3_trials"), {k:round(v,4) for k,v in b["roofline"]["ms_per_trial_by_class"].items()})
PY
done
