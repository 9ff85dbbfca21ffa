#!/bin/bash
mkdir -p gpurun_out/r02; cd /root/repo; L=gpurun_out/r02/j5_times.log; : > $L
run() { echo "## $*" >> $L; env "$@" python tools/render_once.py C2 3 >> $L 2>&1; env "$@" python tools/render_once.py C4M 3 >> $L 2>&1; }
P=/root/repo/rust-raytracer_b200
run RTB200_LIB=$P/librtb200_caps.so RTB200_WF_SMEM=0 RTB200_WF_MINB=4
run RTB200_LIB=$P/librtb200_caps2.so RTB200_WF_SMEM=0
run RTB200_LIB=$P/librtb200_caps2.so RTB200_WF_SMEM=0 RTB200_WF_MINB=4
cat $L
