for mb in 3 4 5 6; do
  AZ_ADV_MINB=$mb timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_adv$mb.json 2>> gpurun_out/b_adv.err
done
AZ_ADV_MINB=5 timeout 300 python -m pytest tests/test_search_gpu.py tests/test_selfplay_gpu.py -x -q 2>&1 | tail -3
