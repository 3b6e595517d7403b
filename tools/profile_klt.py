"""Small driver for ncu: KLT c3 -- first() + N next_dev() steps with frames resident in HBM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from coslam_b200 import api

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
seqs = bench.make_sequences(0)
grp = api.KltGroup(bench.klt_cfg(), bench.KLT_C, bench.KLT_W, bench.KLT_H, bench.KLT_L,
                   bench.KLT_FW, bench.KLT_FH)
dev = [[torch.from_numpy(seqs[c].frames[k]).cuda() for c in range(bench.KLT_C)] for k in range(4)]
grp.first([seqs[c].frames[0] for c in range(bench.KLT_C)])
for i in range(1, 1 + steps):
    grp.next_dev([t.data_ptr() for t in dev[i % 4]], bench.KLT_W)
grp.sync()
f, n = grp.fetch()
print("tracked", [(f[c]["status"] == 0).sum() for c in range(bench.KLT_C)], "launches",
      api.kernel_launch_count())
