#!/bin/bash
# run bench.py --no-cpu for the default library and every build/variants/*.so; prints one line each
for lib in coslam_b200/libcoslam_b200.so build/variants/*.so; do
  COSLAM_B200_LIB=$PWD/$lib timeout 200 python bench.py --no-cpu --quick 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); r=d['roofline']
print('$lib', 'step_ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4), 'track_ms', round(r['avg_ms_per_step'],4), {k: round(v*d['ms_per_step'],4) for k,v in r['share_of_step'].items()})"
done
