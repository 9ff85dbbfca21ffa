#!/bin/bash
# round 2, job 3: full GPU tests of the BVH kernel + ncu full captures (C2, C4M)
mkdir -p gpurun_out/r02
cd /root/repo
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02/j3_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02/j3_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rt_wavefront -s 1 -c 1 -f -o gpurun_out/r02/j3_trace_c2 python tools/render_once.py C2 2 > gpurun_out/r02/j3_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rt_wavefront -s 1 -c 1 -f -o gpurun_out/r02/j3_trace_c4 python tools/render_once.py C4M 2 > gpurun_out/r02/j3_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
