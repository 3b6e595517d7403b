#!/bin/bash
N=$(nvidia-smi -L | wc -l); O=gpurun_out/r2mgt$N; mkdir -p $O
export COSL_BA_TIMING_FINE=1 COSL_BA_SCHUR_SIMT=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR tools/mgpu_ba_time.py 2>&1 | grep "rank" | tee $O/default.txt
NCCL_NVLS_ENABLE=0 timeout 300 $TR tools/mgpu_ba_time.py 2>&1 | grep "rank" | tee $O/nvls0.txt
NCCL_P2P_DISABLE=1 timeout 300 $TR tools/mgpu_ba_time.py 2>&1 | grep "rank" | tee $O/p2p0.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw --format=csv
