#!/bin/bash
# k_advance experiments (profiles/r1_conv_ablation.md "later experiments")
for it in 2 1 2 1; do
  AZ_ADV_MAX_ITERS=$it timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_it$it.json 2>> gpurun_out/b_it.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/b_it$it.json').read().strip().splitlines()[-1])
print($it, round(d['value']), {k:round(v,1) for k,v in d['wave_phases_us'].items()}, d['clocks']['sm_mhz'])
PY
done
