"""profiles/traffic.json: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of
the kernels captured with `ncu --set full`, parsed from profiles/<prefix>_<kernel>.summary.txt.
bench.py reads it to fill roofline.traffic.  usage: python tools/make_traffic.py r1 r2d"""
import glob
import json
import os
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
prefixes = sys.argv[1:] if len(sys.argv) > 1 else ["r1"]  # later prefixes override earlier ones
out = {}
for prefix, path in [(p, f) for p in prefixes for f in sorted(glob.glob(os.path.join(root, f"{p}_*.summary.txt")))]:
    name = os.path.basename(path)[len(prefix) + 1:-len(".summary.txt")]
    tot = 0.0
    for line in open(path):
        m = re.match(r"\s*dram__bytes_(read|write)\.sum\s+([0-9.,]+)\s+(\w+)", line)
        if m:
            tot += float(m.group(2).replace(",", "")) * UNIT.get(m.group(3), 1.0)
    out[name] = {"dram_bytes_per_launch": tot, "source": os.path.basename(path)}
json.dump(out, open(os.path.join(root, "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
