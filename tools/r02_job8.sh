#!/bin/bash
mkdir -p gpurun_out/r02; cd /root/repo
RTB200_PRINT_QUEUES=1 python tools/render_once.py C2 3 > gpurun_out/r02/j8_queues.log 2>&1
RTB200_KERNEL=wavefront python tools/render_once.py C2 3 >> gpurun_out/r02/j8_queues.log 2>&1
cat gpurun_out/r02/j8_queues.log
timeout 600 ncu --metrics smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:rt_stream -s 1 -c 1 python tools/render_once.py C2 2 > gpurun_out/r02/j8_ncu.log 2>&1; grep -E "no_instruction|issue_active|duration" gpurun_out/r02/j8_ncu.log
