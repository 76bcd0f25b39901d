#!/bin/bash
# thread pencils forced on: the whole suite + 300 fuzz seeds; shape A/B
set -u
O=gpurun_out/r2l; mkdir -p $O
RT_B200_PENCIL_THREAD=1 timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu_tp.log 2>&1; echo "pytest (thread pencils on) rc=$?" | tee -a $O/summary.txt
RT_B200_PENCIL_THREAD=1 RT_FUZZ_SEEDS=300 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k fuzz > $O/pytest_fuzz300_tp.log 2>&1; echo "fuzz300 (thread pencils on) rc=$?" | tee -a $O/summary.txt
RT_B200_PENCIL_THREAD=1 timeout 300 python bench.py --workload dodge --steps 3 --warmup 3 --no-cpu-baseline --no-accelerated > $O/bench_dodge_tp1.json 2> $O/bench_dodge_tp1.err; echo "dodge rc=$?" | tee -a $O/summary.txt
RT_B200_PENCIL_THREAD=1 timeout 300 python bench.py --workload cube --steps 10 --warmup 3 --no-cpu-baseline --no-accelerated > $O/bench_cube_tp1.json 2> $O/bench_cube_tp1.err; echo "cube rc=$?" | tee -a $O/summary.txt
