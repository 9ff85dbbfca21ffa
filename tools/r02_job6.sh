#!/bin/bash
# round 2, job 6: first run of the queue-driven kernel (rt_stream_kernel): parity tests under a timeout, then timings
mkdir -p gpurun_out/r02; cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02/j6_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02/j6_pytest.log
L=gpurun_out/r02/j6_times.log; : > $L
run() { echo "## $*" >> $L; env "$@" timeout 120 python tools/render_once.py C2 3 >> $L 2>&1; env "$@" timeout 120 python tools/render_once.py C4M 3 >> $L 2>&1; }
run RTB200_KERNEL=stream
run RTB200_KERNEL=stream RTB200_WF_SMEM=0
run RTB200_KERNEL=stream RTB200_WF_SMEM=1
run RTB200_KERNEL=stream RTB200_POOL_SLOTS=1024
run RTB200_KERNEL=wavefront
cat $L
