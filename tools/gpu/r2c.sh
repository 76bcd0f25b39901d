#!/bin/bash
# Round 2, GPU call C (1 GPU): reflection pencils -- tests, A/B on the headline frame, rt_trace latency.
set -u
O=gpurun_out/r2c; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
RT_FUZZ_SEEDS=200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k fuzz > $O/pytest_fuzz200.log 2>&1; echo "fuzz200 rc=$?" | tee -a $O/summary.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
RT_B200_PENCIL_REFLECT=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-accelerated > $O/bench_n1_noreflect.json 2> $O/bench_n1_noreflect.err
RT_B200_LAUNCHLOG=1 python tools/prof_one.py 4 > $O/launchlog.txt 2>&1
python bench.py --workload dodge --steps 3 --warmup 3 --no-accelerated --no-cpu-baseline > $O/bench_dodge.json 2> $O/bench_dodge.err
python tools/trace_latency.py > $O/trace_latency.json 2> $O/trace_latency.err
ls -la $O >> $O/summary.txt
