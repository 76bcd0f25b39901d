#!/bin/bash
# thread pencils: first light
set -u
O=gpurun_out/r2k; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "thread_pencils or reflection_pencils or fixture" > $O/pytest_tp.log 2>&1; echo "pytest tp rc=$?" | tee -a $O/summary.txt
for v in 0 1; do
  RT_B200_PENCIL_THREAD=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-accelerated > $O/bench_tp$v.json 2> $O/bench_tp$v.err; echo "bench tp=$v rc=$?" | tee -a $O/summary.txt
done
RT_B200_PENCIL_THREAD=1 RT_B200_LAUNCHLOG=1 timeout 120 python tools/prof_one.py 4 > $O/launchlog.txt 2>&1
