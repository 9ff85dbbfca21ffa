#!/bin/bash
# round 2, job 1 (shipped round-1 kernel): tests, C1 diagnosis, sanitizer logs, ncu full of C2 and C4, C3/C4 timings
mkdir -p gpurun_out/r02
cd /root/repo
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r02/j1_pytest.log 2>&1; echo "pytest rc=$?"
timeout 300 python tools/r02_diag_c1.py > gpurun_out/r02/j1_diag_c1.log 2>&1; echo "diag rc=$?"
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r02/j1_memcheck.log python tools/r02_san.py > gpurun_out/r02/j1_memcheck_stdout.log 2>&1; echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02/j1_racecheck.log python tools/r02_san.py > gpurun_out/r02/j1_racecheck_stdout.log 2>&1; echo "racecheck rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rt_wavefront -s 1 -c 1 -f -o gpurun_out/r02/j1_trace_c2 python tools/render_once.py C2 2 > gpurun_out/r02/j1_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rt_wavefront -s 1 -c 1 -f -o gpurun_out/r02/j1_trace_c4 python tools/render_once.py C4M 2 > gpurun_out/r02/j1_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 300 python tools/render_once.py C3 2 > gpurun_out/r02/j1_c3.log 2>&1
timeout 300 python tools/render_once.py C4M 2 > gpurun_out/r02/j1_c4s.log 2>&1
tail -3 gpurun_out/r02/j1_pytest.log gpurun_out/r02/j1_diag_c1.log gpurun_out/r02/j1_c3.log gpurun_out/r02/j1_c4s.log
ls -la gpurun_out/r02
