#!/bin/bash
O=gpurun_out/r2f; mkdir -p $O
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -5 $O/bench.err
head -c 1500 $O/bench.json; echo
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; head -c 400 $O/bench_reference.json; echo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest_gpu.txt
