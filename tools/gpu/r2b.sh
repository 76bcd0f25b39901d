#!/bin/bash
# Round 2, GPU call B (1 GPU): new tests, ubench variants, bench line with parity, C1/C3, rt_trace latency.
set -u
O=gpurun_out/r2b; mkdir -p $O
tools/ubench/pencil > $O/ubench_pencil.txt 2>&1
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
python bench.py --workload cube --steps 20 --warmup 3 --no-accelerated > $O/bench_cube.json 2> $O/bench_cube.err
RT_B200_GRAPH=0 python bench.py --workload cube --steps 20 --warmup 3 --no-accelerated --no-cpu-baseline > $O/bench_cube_nograph.json 2> $O/bench_cube_nograph.err
python bench.py --workload dodge --steps 3 --warmup 3 --no-accelerated > $O/bench_dodge.json 2> $O/bench_dodge.err
python tools/trace_latency.py > $O/trace_latency.json 2> $O/trace_latency.err
ls -la $O >> $O/summary.txt
