#!/bin/bash
# short re-run of the scaling end points (N = 1 and N = 8 on one box) after the shard-ring change in rtb200/dist.py
mkdir -p gpurun_out/r02; cd /root/repo; O=gpurun_out/r02
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/s18_C2_n1.json 2> $O/s18_C2_n1.err; echo "n1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 20 --warmup 5 > $O/s18_C2_n8.json 2> $O/s18_C2_n8.err; echo "n8 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29609 bench.py --gpus 8 --steps 40 --warmup 5 > $O/s18_C2_n8b.json 2> $O/s18_C2_n8b.err; echo "n8b rc=$?"
python - <<'P'
import json
for f in ("s18_C2_n1","s18_C2_n8","s18_C2_n8b"):
    d=json.loads(open(f"gpurun_out/r02/{f}.json").read().strip().splitlines()[-1])
    print(f, "value %.0f ms %.3f e2e %.0f ms %.3f golden %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["golden"]))
P
