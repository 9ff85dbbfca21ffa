#!/bin/bash
# Run under gpurun on ONE GPU: launch list + one full ncu capture of the trace kernel. Outputs -> gpurun_out/
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
# (1) every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# (2) the top kernel, full set, once
ncu --set full --clock-control none --import-source on -k regex:rt_wavefront -s 1 -c 1 -f -o gpurun_out/trace_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
