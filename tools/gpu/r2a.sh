#!/bin/bash
# Round 2, GPU call A (1 GPU): RT_OPT_PENCIL_ANY on the GPU, C1/C3 before/after, instruction-count captures.
set -u
O=gpurun_out/r2a; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_default.log 2>&1; echo "pytest default rc=$?" | tee -a $O/summary.txt
RT_B200_PENCIL_ANY=1 python -m pytest tests -m gpu -x -q > $O/pytest_pencil_any.log 2>&1; echo "pytest pencil_any rc=$?" | tee -a $O/summary.txt
RT_B200_PENCIL_ANY=1 RT_FUZZ_SEEDS=300 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k fuzz > $O/pytest_pencil_any_fuzz300.log 2>&1; echo "fuzz300 pencil_any rc=$?" | tee -a $O/summary.txt
for w in cube dodge; do
  python tools/run_config.py --workload $w --frames 5 --modes brute_force > $O/${w}_default.json 2> $O/${w}_default.err
  RT_B200_PENCIL_ANY=1 python tools/run_config.py --workload $w --frames 5 --modes brute_force > $O/${w}_pencil_any.json 2> $O/${w}_pencil_any.err
done
# executed-instruction counters of every scan launch of one 16-spp frame (second frame of prof_one.py)
python tools/prof_one.py 4 > $O/prof_one_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__inst_executed.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"k_trace|k_shadow" --csv --log-file $O/inst_counts.csv python tools/prof_one.py 4 > $O/ncu_inst.log 2>&1
# full capture of the largest pencil k_shadow launch (level 0 of chunk 0, second frame: 8 shadow launches per frame)
ncu --set full --clock-control none --import-source on -k regex:k_shadow -s 8 -c 1 -o $O/prof_shadow_r2a -f python tools/prof_one.py 4 > $O/ncu_full_shadow.log 2>&1
ls -la $O | tee -a $O/summary.txt
