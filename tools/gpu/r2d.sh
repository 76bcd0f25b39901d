#!/bin/bash
# Round 2, GPU call D (1 GPU): dynamic shared memory + the scalar 8-ray / 6-ray pencil shapes, A/B on the headline frame.
set -u
O=gpurun_out/r2d; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
for cfg in 2,8,2 4,4,2 3,4,2; do
  RT_B200_PTUNE=$cfg python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-accelerated --no-parity > $O/bench_ptune_$cfg.json 2> $O/bench_ptune_$cfg.err; echo "bench $cfg rc=$?" | tee -a $O/summary.txt
done
RT_B200_PTUNE=4,4,2 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_ptune442.log 2>&1; echo "pytest 4,4,2 rc=$?" | tee -a $O/summary.txt
RT_B200_PTUNE=3,4,2 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_ptune342.log 2>&1; echo "pytest 3,4,2 rc=$?" | tee -a $O/summary.txt
RT_B200_PTUNE=4,4,2 python bench.py --workload dodge --steps 3 --warmup 3 --no-accelerated --no-cpu-baseline > $O/bench_dodge_442.json 2> $O/bench_dodge_442.err
ls -la $O >> $O/summary.txt
