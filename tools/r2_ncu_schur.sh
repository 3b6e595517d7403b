#!/bin/bash
O=gpurun_out/r2n; mkdir -p $O
timeout 900 ncu --set full --import-source on --clock-control none -k regex:${KERN:-ba_schur_blk} -s 4 -c 1 -o $O/${TAG:-ba_schur_blk} python tools/r2_ba_exp.py -1 > $O/ncu_st.log 2>&1
tail -3 $O/ncu_st.log
ncu -i $O/${TAG:-ba_schur_blk}.ncu-rep --page details --csv > $O/${TAG:-ba_schur_blk}.details.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r2n/"+__import__("os").environ.get("TAG","ba_schur_blk")+".details.csv")))
h=rows[0]; n=h.index("Metric Name"); v=h.index("Metric Value"); u=h.index("Metric Unit"); s=h.index("Section Name")
keep=("Duration","DRAM Throughput","Memory Throughput","L1/TEX Hit Rate","L2 Hit Rate","Compute (SM) Throughput","Registers Per Thread","Achieved Occupancy","Theoretical Occupancy","Executed Ipc Active","Issue Slots Busy","L1/TEX Cache Throughput","L2 Cache Throughput","Mem Busy","Max Bandwidth","No Eligible","Warp Cycles Per Issued Instruction","Shared Memory Configuration Size","Dynamic Shared Memory Per Block","Block Limit Registers","Block Limit Shared Mem","Mem Pipes Busy")
for r in rows[1:]:
    if any(k==r[n] for k in keep): print(f"{r[s][:34]:34s} {r[n]:44s} {r[v]:>14s} {r[u]}")
PY
