#!/bin/bash
mkdir -p gpurun_out/r02; cd /root/repo; L=gpurun_out/r02/j9_times.log; : > $L
run() { echo "## $*" >> $L; env "$@" timeout 120 python tools/render_once.py C2 3 >> $L 2>&1; env "$@" timeout 120 python tools/render_once.py C4M 3 >> $L 2>&1; }
P=/root/repo/rust-raytracer_b200
run RTB200_LIB=$P/librtb200.so
run RTB200_LIB=$P/librtb200_b128.so
run RTB200_LIB=$P/librtb200_b384.so
run RTB200_LIB=$P/librtb200_b768.so
grep -E "^##|Mrays|regs" $L
