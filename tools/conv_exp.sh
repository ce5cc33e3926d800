#!/bin/bash
# Ablation of the per-layer tcgen05 convolution (profiles/r1_conv_ablation.md).  AZ_DBG_CONV bit flags:
# 1 no activation TMA loads, 2 no MMA issue, 4 epilogue only signals, 8 no residual loads, 16 no output stores.
for d in 0 1 2 4 16 24; do
  echo "dbg=$d"
  AZ_DBG_CONV=$d timeout 120 python -m pytest tests/test_tc_conv_gpu.py -q -s -k speed 2>&1 | grep "conv3x3 tc"
done
