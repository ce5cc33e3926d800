#!/bin/bash
# A/B of two BUILDS on one box: tools/ab_build.sh "<nvcc defines A>" "<nvcc defines B>"   (e.g. "" "-DAZ_EPI_WARPS=8")
# runs bench.py --quick twice per build, interleaved; restores the default build at the end
cd "$(dirname "$0")/.."
run() { python bench.py --quick --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); w=d["wave_phases_us"]; print("%.4f M sims/s | adv %.1f in %.1f tower %.1f heads %.1f wave %.1f | clk %s" % (d["value"]/1e6, w["k_advance"], w["input_conv"], w["tower"], w["heads"], w["wave_total"], d["clocks"]["sm_mhz"]))'; }
for pass in 1 2; do
  for defs in "$@"; do
    AZ_NVCC_DEFINES="$defs" python alphazero-chess_b200/build.py > /dev/null
    echo "[$defs] $(run)"
  done
done
python alphazero-chess_b200/build.py > /dev/null
