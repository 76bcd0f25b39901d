#!/bin/bash
set -u
O=gpurun_out/r2g; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 16 -c 1 -o $O/prof_mirror -f python tools/prof_one.py 4 > $O/ncu_full_mirror.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 12 -c 1 -o $O/prof_bounce -f python tools/prof_one.py 4 > $O/ncu_full_bounce.log 2>&1
