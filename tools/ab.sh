#!/bin/bash
# interleaved A/B of environment switches on one box: tools/ab.sh "NAME=VAL ..." "NAME=VAL ..." ...   (each run: bench.py --quick)
# prints value (sims/s) and the per-wave phase times of every run, two passes
for pass in 1 2; do
  for cfg in "$@"; do
    out=$(env $cfg python bench.py --quick --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1)
    echo "$cfg | $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); w=d["wave_phases_us"]; print("%.4f M sims/s | adv %.1f in %.1f tower %.1f heads %.1f wave %.1f | clk %s" % (d["value"]/1e6, w["k_advance"], w["input_conv"], w["tower"], w["heads"], w["wave_total"], d["clocks"]["sm_mhz"]))')"
  done
done
