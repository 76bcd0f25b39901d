#!/bin/bash
# Round 2 measurement matrix at N GPUs of one box (north_star: C1..C5 at 1/2/4/8 GPUs).   bash tools/gpu/matrix.sh N [tag]
# Every JSON lands in gpurun_out/<tag>/ ; copy what is to be judged into profiles/.
set -u
N=${1:-1}; TAG=${2:-r2m_n$N}; O=gpurun_out/$TAG; mkdir -p $O
if [ "$N" = "1" ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"; fi
t0=$(date +%s)
# C2 (headline) through bench.py: value, e2e, parity against the reference's pinned frame, roofline, cpu_baseline
$RUN bench.py --gpus $N --steps 10 --warmup 3 > $O/C2_bench_n$N.json 2> $O/C2_bench_n$N.err; echo "C2 rc=$? $(( $(date +%s) - t0 )) s" | tee -a $O/summary.txt
# the same frame driven by ONE process (rt_init(N), ncclCommInitAll): the C++ drop-in's mode
if [ "$N" != "1" ]; then
  python bench.py --gpus $N --single-process --steps 10 --warmup 3 --no-cpu-baseline --no-accelerated > $O/C2_bench_single_process_n$N.json 2> $O/C2_bench_single_process_n$N.err; echo "C2 single-process rc=$? $(( $(date +%s) - t0 )) s" | tee -a $O/summary.txt
fi
# C1, C3 through bench.py (cpu_baseline with core count and 1-thread figure; parity where a pin exists)
$RUN bench.py --gpus $N --workload cube --steps 20 --warmup 3 > $O/C1_bench_n$N.json 2> $O/C1_bench_n$N.err; echo "C1 rc=$? $(( $(date +%s) - t0 )) s" | tee -a $O/summary.txt
$RUN bench.py --gpus $N --workload dodge --steps 5 --warmup 3 > $O/C3_bench_n$N.json 2> $O/C3_bench_n$N.err; echo "C3 rc=$? $(( $(date +%s) - t0 )) s" | tee -a $O/summary.txt
# C4: 1 M triangles, 3840x2160x16 -- one timed frame per mode (brute force, tile culling), oracle lattice check, CPU lattice timing at N = 1
CPUPIX=0; [ "$N" = "1" ] && CPUPIX=96
$RUN tools/run_config.py --workload sphere1m --frames 1 --cpu-pixels $CPUPIX > $O/C4_sphere1m_n$N.json 2> $O/C4_sphere1m_n$N.err; echo "C4 rc=$? $(( $(date +%s) - t0 )) s" | tee -a $O/summary.txt
# C5: sizes x samples; everything up to 3e8 samples per GPU with 2 timed frames, the larger ones with 1
$RUN tools/sweep.py --sizes 512,1024,2048,4096,8192 --pf 1,2,4,8 --max-samples 1.5e8 --frames 2 > $O/C5_n$N.jsonl 2> $O/C5_n$N.err; echo "C5 small rc=$? $(( $(date +%s) - t0 )) s" | tee -a $O/summary.txt
$RUN tools/sweep.py --sizes 2048,4096,8192 --pf 4,8 --min-samples 1.5e8 --max-samples ${MAXS:-6e8} --frames 1 >> $O/C5_n$N.jsonl 2>> $O/C5_n$N.err; echo "C5 large rc=$? $(( $(date +%s) - t0 )) s" | tee -a $O/summary.txt
if [ "$N" != "1" ]; then
  python -m pytest tests/test_gpu_app.py -m gpu -x -q > $O/pytest_multi_gpu.log 2>&1; echo "pytest multi-gpu rc=$? $(( $(date +%s) - t0 )) s" | tee -a $O/summary.txt
fi
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/nvidia_smi.csv 2>&1
