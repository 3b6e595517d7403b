#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
for c in c3 c2; do
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/local_${c}_launches.csv python tools/r2_local_one.py $c 6 > $O/local_${c}.log 2>&1
tail -2 $O/local_${c}.log
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$O/local_${c}_launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
seq=[(r[ki].split("(")[0][-40:], float(r[vi].replace(",",""))*(1e-3 if r[ui]=="ns" else 1)) for r in rows[1:]]
# last trial = kernels after the last ba_linearize_points
idx=[i for i,(k,_) in enumerate(seq) if "ba_linearize_points" in k]
last=seq[idx[-1]:]
print("$c last trial kernels (us):")
for k,v in last: print(f"  {k:42s} {v:8.2f}")
print("  sum", round(sum(v for _,v in last),1))
PY
done
