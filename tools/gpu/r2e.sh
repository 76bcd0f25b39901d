#!/bin/bash
set -u
O=gpurun_out/r2e; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
python tools/trace_latency.py > $O/trace_latency.json 2> $O/trace_latency.err; echo "latency rc=$?" | tee -a $O/summary.txt
RT_B200_SMALL_TRACE=0 python tools/trace_latency.py > $O/trace_latency_wavefront.json 2> $O/trace_latency_wavefront.err
