timeout 300 python -m pytest tests/test_tc_conv_gpu.py -x -q 2>&1 | tail -2
for d in 0 2 4 24; do echo "dbg=$d"; AZ_DBG_CONV=$d timeout 120 python -m pytest tests/test_tc_conv_gpu.py -q -s -k speed 2>&1 | grep "conv3x3 tc"; done
