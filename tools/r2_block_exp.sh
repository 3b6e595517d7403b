#!/bin/bash
N=$(nvidia-smi -L | wc -l); O=gpurun_out/r2b$N; mkdir -p $O
for S in 0 1; do
if [ $S = 1 ]; then export COSL_BA_BLOCKING_READBACK=1; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$S bench.py --gpus $N --steps 20 --warmup 5 --quick --no-cpu > $O/bench_b$S.json 2> $O/bench_b$S.err
python - <<PY
import json
d=json.load(open("$O/bench_b$S.json")); b=d["ba"]
print("blocking=$S N=$N ba", round(b["value"],1), round(b["ms_per_trial"],4), "parity", b.get("parity_rel_cost_diff_3_trials"))
PY
done
COSL_BA_BLOCKING_READBACK=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29525 tools/mgpu_ba_check.py 2>&1 | grep "MGPU_PARITY"
