#!/bin/bash
# tower experiment: board ranges that fit the L2, with and without the input convolution inside the launch
for cfg in "1 2" "2 2" "1 2" "2 2"; do
  set -- $cfg
  AZ_TOWER_FUSED=$1 AZ_TOWER_SPLIT=$2 timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_split.json 2>> gpurun_out/b_split.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/b_split.json').read().strip().splitlines()[-1])
print("fused $1 split $2", round(d['value']), {k:round(v,1) for k,v in d['wave_phases_us'].items()}, d['clocks']['sm_mhz'])
PY
done
