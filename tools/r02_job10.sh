#!/bin/bash
mkdir -p gpurun_out/r02; cd /root/repo; L=gpurun_out/r02/j10_times.log; : > $L
P=/root/repo/rust-raytracer_b200
for lib in librtb200_b192.so librtb200_b384.so; do echo "## $lib" >> $L; RTB200_LIB=$P/$lib timeout 120 python tools/render_once.py C2 3 >> $L 2>&1; RTB200_LIB=$P/$lib timeout 120 python tools/render_once.py C4M 3 >> $L 2>&1; done
echo "## shards: what one GPU of an N-GPU run does (C2)" >> $L
for w in 1 2 4 8; do python tools/render_once.py C2 4 0 $w >> $L 2>&1; done
echo "## C3 / C5S shard of 8" >> $L
python tools/render_once.py C3 2 0 1 >> $L 2>&1
python tools/render_once.py C3 2 0 8 >> $L 2>&1
grep -E "^##|Mrays|regs" $L
