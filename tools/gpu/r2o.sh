#!/bin/bash
set -u
N=8; O=gpurun_out/r2n_n8; mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$RUN bench.py --gpus $N --steps 10 --warmup 3 > $O/C2_bench_n$N.json 2> $O/C2_bench_n$N.err; echo "C2 rc=$?" | tee -a $O/summary.txt
python bench.py --gpus $N --single-process --steps 10 --warmup 3 --no-cpu-baseline --no-accelerated > $O/C2_bench_single_process_n$N.json 2> $O/C2_bench_single_process_n$N.err; echo "C2 single-process rc=$?" | tee -a $O/summary.txt
$RUN tools/run_config.py --workload sphere1m --frames 1 --modes brute_force > $O/C4_sphere1m_n$N.json 2> $O/C4_sphere1m_n$N.err; echo "C4 rc=$?" | tee -a $O/summary.txt
python -m pytest tests/test_gpu_app.py -m gpu -x -q > $O/pytest_multi_gpu.log 2>&1; echo "pytest multi-gpu rc=$?" | tee -a $O/summary.txt
