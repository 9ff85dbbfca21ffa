#!/bin/bash
# 2-GPU check of the one-process multi-GPU entry point (threads + peer copies) and of bench.py's N=2 path
mkdir -p gpurun_out/r02; cd /root/repo; O=gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity.py tests/test_host_cli.py -m gpu -q -k "dist or multi or harness or cli" > $O/j16_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/j16_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 5 > $O/j16_bench_n2.json 2> $O/j16_bench_n2.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02/j16_bench_n2.json").read().strip().splitlines()[-1])
print("value %.0f ms/step %.2f e2e %.0f ms %.2f golden %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["golden"]))
P
tail -3 $O/j16_bench_n2.err
RTB200_GPUS=2 RTB200_STATS=1 ./rust-raytracer_b200/raytracer /dev/null /dev/null 2>&1 | head -2
