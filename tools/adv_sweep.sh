#!/bin/bash
# relaxed vs release arrive when an epilogue warp hands its TMEM accumulator back (A/B on one box)
timeout 300 python -m pytest tests/test_tc_conv_gpu.py tests/test_nn_gpu.py -x -q 2>&1 | tail -2
for r in 1 0 1 0; do
  AZ_TC_RELEASE_ARRIVE=$r timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_rel.json 2>> gpurun_out/b_rel.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/b_rel.json').read().strip().splitlines()[-1])
print("release $r", round(d['value']), {k:round(v,1) for k,v in d['wave_phases_us'].items()}, d['clocks']['sm_mhz'])
PY
done
