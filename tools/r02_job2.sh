#!/bin/bash
# round 2, job 2: first run of the BVH / warp-cooperative kernel: parity tests, then timings
mkdir -p gpurun_out/r02
cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02/j2_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02/j2_pytest.log
for c in C2 C3S C4M C5S; do timeout 300 python tools/render_once.py $c 3 >> gpurun_out/r02/j2_times.log 2>&1; done
cat gpurun_out/r02/j2_times.log
