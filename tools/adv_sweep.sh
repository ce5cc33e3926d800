#!/bin/bash
# input convolution inside the per-range tower launch, after the relaxed accumulator hand-back (A/B on one box)
for f in 1 2 1 2; do
  AZ_TOWER_FUSED=$f timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_f.json 2>> gpurun_out/b_f.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/b_f.json').read().strip().splitlines()[-1])
print("fused $f", round(d['value']), {k:round(v,1) for k,v in d['wave_phases_us'].items()}, d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
PY
done
