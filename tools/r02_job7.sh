#!/bin/bash
mkdir -p gpurun_out/r02; cd /root/repo
RTB200_PRINT_QUEUES=1 python tools/render_once.py C2 2 > gpurun_out/r02/j7_queues.log 2>&1
RTB200_PRINT_QUEUES=1 python tools/render_once.py C4M 2 >> gpurun_out/r02/j7_queues.log 2>&1
cat gpurun_out/r02/j7_queues.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rt_stream -s 1 -c 1 -f -o gpurun_out/r02/j7_stream_c2 python tools/render_once.py C2 2 > gpurun_out/r02/j7_ncu_c2.log 2>&1; echo "ncu rc=$?"
