#!/bin/bash
# k_advance experiments (profiles/r1_conv_ablation.md "later experiments"): simulations a game may finish per wave
for it in 8 1 2; do
  AZ_ADV_MAX_ITERS=$it timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_it$it.json 2>> gpurun_out/b_it.err
done
AZ_ADV_MAX_ITERS=1 timeout 300 python -m pytest tests/test_search_gpu.py tests/test_selfplay_gpu.py -x -q 2>&1 | tail -3
