#!/bin/bash
# Round 2, GPU call F (1 GPU): ncu evidence for the shipped kernels.
set -u
O=gpurun_out/r2f; mkdir -p $O
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-accelerated --no-parity > $O/bench_plain.json 2> $O/bench_plain.err; echo "bench rc=$?" | tee -a $O/summary.txt
# launch list of the same command (the recipe's pass): per-launch times, cold-cache and serialised -- the SHARES must agree with the bench's
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_balls.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-accelerated --no-parity > $O/ncu_launches.log 2>&1
# counters of every scan launch of one frame
python tools/prof_one.py 4 > $O/prof_one_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"k_trace|k_shadow" --csv --log-file $O/launches_counters_balls.csv python tools/prof_one.py 4 > $O/ncu_counters.log 2>&1
# full captures: the mirror-pencil scan (k_trace<...,0,0,0,1>: non-primary pencil) and the level-1 generic scan of chunk 0, second frame
ncu --set full --clock-control none --import-source on -k regex:"k_trace<2, 8, 2, 0, 0, 0, 1>" -s 2 -c 1 -o $O/prof_mirror -f python tools/prof_one.py 4 > $O/ncu_full_mirror.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace<2, 8, 2, 0, 0, 0, 0>" -s 6 -c 1 -o $O/prof_bounce -f python tools/prof_one.py 4 > $O/ncu_full_bounce.log 2>&1
ls -la $O >> $O/summary.txt
