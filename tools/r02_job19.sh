#!/bin/bash
mkdir -p gpurun_out/r02; cd /root/repo; L=gpurun_out/r02/j19.log; : > $L
for w in 8 1; do for a in 3 2 1; do RTB200_ASYNC_CTAS=$a python tools/pipeline_probe.py C2 $w 40 >> $L 2>&1; done; done
RTB200_ASYNC_CTAS=3 python tools/pipeline_probe.py C4M 8 40 >> $L 2>&1
RTB200_ASYNC_CTAS=2 python tools/pipeline_probe.py C4M 8 40 >> $L 2>&1
cat $L
