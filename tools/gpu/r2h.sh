#!/bin/bash
# Round 2, GPU call H (1 GPU): the whole GPU suite (incl. the full-frame pins), then the matrix at N = 1.
set -u
O=gpurun_out/r2m_n1; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
python bench.py --impl reference --steps 3 --warmup 1 > $O/C2_bench_reference_arm.json 2> $O/C2_bench_reference_arm.err; echo "reference arm rc=$?" | tee -a $O/summary.txt
bash tools/gpu/matrix.sh 1 r2m_n1
