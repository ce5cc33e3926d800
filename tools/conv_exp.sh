for d in 0 1 2 3 4 5 6 7; do echo "dbg=$d"; AZ_DBG_CONV=$d timeout 120 python -m pytest tests/test_tc_conv_gpu.py -q -s -k speed 2>&1 | grep "conv3x3 tc"; done
for g in 148 74 296; do echo "grid=$g"; AZ_DBG_GRID=$g timeout 120 python -m pytest tests/test_tc_conv_gpu.py -q -s -k speed 2>&1 | grep "conv3x3 tc"; done
