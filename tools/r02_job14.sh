#!/bin/bash
mkdir -p gpurun_out/r02; cd /root/repo; L=gpurun_out/r02/j14_times.log; : > $L
for w in 1 8; do RTB200_PRINT_TAIL=1 python tools/render_once.py C2 2 0 $w >> $L 2>&1; done
RTB200_PRINT_TAIL=1 python tools/render_once.py 800x600x16 2 >> $L 2>&1
cat $L
