#!/bin/bash
# Instrumented tcgen05 heads kernel (per-phase clocks of block 0: TMA warp, MMA warp, first warp of each epilogue group).
# The instrumented library is built HERE (no GPU needed), shipped with the gpurun snapshot, and the normal one rebuilt afterwards:
#   tools/htc_timing.sh        -> gpurun_out/r2_htc_timing.log
set -e
cd "$(dirname "$0")/.."
AZ_NVCC_DEFINES="-DAZ_HTC_TIMING" python alphazero-chess_b200/build.py > /dev/null
tools/gpurun_retry.sh --timeout 600 -- 'AZ_HEADS_TC=1 timeout 300 python tools/bench_short.py 24 2 > gpurun_out/r2_htc_timing.log 2>&1; tail -9 gpurun_out/r2_htc_timing.log' 2>&1 | tail -14
python alphazero-chess_b200/build.py > /dev/null
