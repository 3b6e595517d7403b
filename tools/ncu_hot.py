"""Per-SASS-line hot spots of an .ncu-rep (needs --import-source on / -lineinfo):
usage: python tools/ncu_hot.py x.ncu-rep [topN]"""
import csv, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
S, N, X = ci["Source"], ci["# Samples"], ci["Instructions Executed"]
data = []
for r in rows[hi + 1:]:
    if len(r) <= N: continue
    try: data.append((int(r[N]), int(r[X] or 0), r[S].strip()))
    except ValueError: pass
tot = sum(d[0] for d in data); totx = sum(d[1] for d in data)
print(f"total samples {tot}, total warp-instructions {totx}, {len(data)} SASS lines")
for n, x, s in sorted(data, reverse=True)[:top]:
    print(f"{n:7d} {n / max(1, tot):6.1%}  exec {x:8d}  {s[:100]}")
