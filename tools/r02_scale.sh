#!/bin/bash
# round 2, 8-GPU job: multi-GPU correctness tests, the C2 scaling run (N = 1, 2, 4, 8, back to back on one box) and the big configs at N = 8
mkdir -p gpurun_out/r02; cd /root/repo; O=gpurun_out/r02
nvidia-smi -L | head -8
run() { # N config steps warmup extra...
  N=$1; C=$2; K=$3; W=$4; shift 4
  if [ $N -eq 1 ]; then timeout 600 python bench.py --gpus 1 --config $C --steps $K --warmup $W "$@" > $O/scale_${C}_n$N.json 2> $O/scale_${C}_n$N.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --config $C --steps $K --warmup $W "$@" > $O/scale_${C}_n$N.json 2> $O/scale_${C}_n$N.err; fi
  echo "bench $C n=$N rc=$? $(tail -c 300 $O/scale_${C}_n$N.json | tr -d '\n' | cut -c1-200)"
}
run 1 C2 20 5 --no-cpu-baseline
run 2 C2 20 5
run 4 C2 20 5
run 8 C2 20 5
run 8 C3 5 3
run 8 C4 3 3
run 8 C5 3 3
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02/scale_*_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.0f e2e %.0f golden %s ms/step %.2f" % (d["value"], d["e2e"]["value"], d["golden"], d["ms_per_step"]))
    except Exception as e:
        print(f, "ERR", e)
P
