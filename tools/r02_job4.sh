#!/bin/bash
# round 2, job 4: tuning experiments (list capacities -> hierarchy in shared memory; leaf size; register budget)
mkdir -p gpurun_out/r02; cd /root/repo; L=gpurun_out/r02/j4_times.log; : > $L
run() { echo "## $*" >> $L; env "$@" python tools/render_once.py C2 3 >> $L 2>&1; env "$@" python tools/render_once.py C4M 3 >> $L 2>&1; }
P=/root/repo/rust-raytracer_b200
run RTB200_LIB=$P/librtb200.so
run RTB200_LIB=$P/librtb200.so RTB200_WF_MINB=2
run RTB200_LIB=$P/librtb200_caps.so
run RTB200_LIB=$P/librtb200_caps.so RTB200_WF_SMEM=0
run RTB200_LIB=$P/librtb200_caps.so RTB200_WF_MINB=2
run RTB200_LIB=$P/librtb200_k4.so
run RTB200_LIB=$P/librtb200_k16.so
cat $L
