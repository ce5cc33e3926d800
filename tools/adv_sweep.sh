#!/bin/bash
# one tower launch that walks both board ranges vs one launch per range (A/B on one box)
AZ_TOWER_INKERNEL=1 timeout 300 python -m pytest tests/test_nn_gpu.py tests/test_search_gpu.py -x -q 2>&1 | tail -2
for k in 0 1 0 1; do
  AZ_TOWER_INKERNEL=$k timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ink.json 2>> gpurun_out/b_ink.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/b_ink.json').read().strip().splitlines()[-1])
print("inkernel $k", round(d['value']), {k:round(v,1) for k,v in d['wave_phases_us'].items()}, d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
PY
done
