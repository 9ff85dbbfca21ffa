#!/bin/bash
mkdir -p gpurun_out/r02; cd /root/repo; L=gpurun_out/r02/j12_times.log; : > $L
run() { echo "## $*" >> $L; env "$@" timeout 120 python tools/render_once.py C2 3 >> $L 2>&1; env "$@" timeout 120 python tools/render_once.py C4M 3 >> $L 2>&1; }
run A=1
run RTB200_WF_SMEM=1
run RTB200_WF_SMEM=7
run RTB200_WF_SMEM=3
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r02/j12_pytest.log 2>&1; tail -2 gpurun_out/r02/j12_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29777 bench.py --impl reference --steps 2 --warmup 1 --cpu-budget 30 > gpurun_out/r02/j12_ref_torchrun.json 2> gpurun_out/r02/j12_ref_torchrun.err; tail -c 700 gpurun_out/r02/j12_ref_torchrun.json
grep -E "^##|Mrays" $L
