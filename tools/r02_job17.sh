#!/bin/bash
mkdir -p gpurun_out/r02; cd /root/repo; L=gpurun_out/r02/j17_times.log; : > $L
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02/j17_pytest.log 2>&1; tail -3 gpurun_out/r02/j17_pytest.log
for w in 1 8; do RTB200_PRINT_TAIL=1 python tools/render_once.py C2 3 0 $w >> $L 2>&1; done
RTB200_PRINT_TAIL=1 python tools/render_once.py C4M 3 >> $L 2>&1
grep -E "tail|Mrays" $L | cut -c1-250
