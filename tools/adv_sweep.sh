#!/bin/bash
# k_advance experiments (profiles/r1_conv_ablation.md "later experiments"): register bound / resident warps per SM
for mb in 4 5 6 7; do
  AZ_ADV_MINB=$mb timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_mb$mb.json 2>> gpurun_out/b_mb.err
done
