#!/bin/bash
# N-GPU validation: the multi-GPU pytest files, the N-rank parity tool on time-ordered points, the bench at N
N=$(nvidia-smi -L | wc -l); O=gpurun_out/r2y$N; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -4 | tee $O/pytest_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/mgpu_ba_check.py 2>&1 | grep "multi-GPU\|MGPU_PARITY" | tee $O/mgpu_check.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "bench rc=$?"; tail -2 $O/bench_n$N.err
python - <<PY
import json
d=json.load(open("$O/bench_n$N.json")); b=d["ba"]
print("N=$N klt", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
print("ba", b["value"], b["ms_per_trial"], "parity", b.get("parity_rel_cost_diff_3_trials"), b["roofline"]["ms_per_trial_by_class"])
print("c5", d.get("c5",{}).get("fps_8cam_e2e"))
PY
