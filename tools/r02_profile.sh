#!/bin/bash
# round 2 evidence job (shipped build): ncu --set full of the trace kernel on C2 and on the 10k-sphere scene, the launch list of a
# bench.py step, compute-sanitizer memcheck + racecheck on small scenes. Outputs -> gpurun_out/r02/final_*
mkdir -p gpurun_out/r02; cd /root/repo; O=gpurun_out/r02
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rt_wavefront -s 1 -c 1 -f -o $O/final_trace_c2 python tools/render_once.py C2 2 > $O/final_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rt_wavefront -s 1 -c 1 -f -o $O/final_trace_c4 python tools/render_once.py C4M 2 > $O/final_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/final_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 compute-sanitizer --tool memcheck --log-file $O/final_memcheck.log python tools/r02_san.py big > $O/final_memcheck_stdout.log 2>&1; echo "memcheck rc=$?"
timeout 1200 compute-sanitizer --tool racecheck --log-file $O/final_racecheck.log python tools/r02_san.py big > $O/final_racecheck_stdout.log 2>&1; echo "racecheck rc=$?"
tail -3 $O/final_memcheck.log $O/final_racecheck.log $O/final_memcheck_stdout.log
