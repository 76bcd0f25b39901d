#!/bin/bash
# final kernels (thread pencils default): whole suite, smoke, N = 1 matrix into r2n_n1, ncu counters
set -u
O=gpurun_out/r2n_n1; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
bash tools/gpu/matrix.sh 1 r2n_n1
python tools/prof_one.py 4 > $O/prof_one_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"k_trace|k_shadow" --csv --log-file $O/launches_counters_balls.csv python tools/prof_one.py 4 > $O/ncu_counters.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_tp -s 2 -c 1 -o $O/prof_thread_pencil -f python tools/prof_one.py 4 > $O/ncu_full_tp.log 2>&1
