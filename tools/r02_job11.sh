#!/bin/bash
# round 2, job 11: bench.py shake-out on one GPU (both arms) + C3/C4 lines
mkdir -p gpurun_out/r02; cd /root/repo
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02/j11_bench_ref.json 2> gpurun_out/r02/j11_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02/j11_bench_c2.json 2> gpurun_out/r02/j11_bench_c2.err; echo "c2 rc=$?"
timeout 900 python bench.py --config C4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02/j11_bench_c4.json 2> gpurun_out/r02/j11_bench_c4.err; echo "c4 rc=$?"
tail -c 1500 gpurun_out/r02/j11_bench_c2.json; tail -5 gpurun_out/r02/j11_bench_c2.err; tail -5 gpurun_out/r02/j11_bench_c4.err; tail -3 gpurun_out/r02/j11_bench_ref.err
