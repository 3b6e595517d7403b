#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
timeout 300 python tools/ba_setup_timing.py 2>&1 | grep -v "^$" | tee $O/setup_timing.txt
