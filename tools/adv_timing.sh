#!/bin/bash
# Instrumented k_advance (per-phase clocks and a histogram of per-warp durations), then the normal library is rebuilt.
#   tools/adv_timing.sh [sims] [plies] [warm-up plies]      (run on the GPU box: gpurun -- tools/adv_timing.sh)
set -e
cd "$(dirname "$0")/.."
AZ_NVCC_DEFINES="-DAZ_ADV_TIMING" python alphazero-chess_b200/build.py > /dev/null
python tools/bench_short.py ${1:-200} ${2:-2} 0 ${3:-0} 2>&1 | tail -8
python alphazero-chess_b200/build.py > /dev/null
