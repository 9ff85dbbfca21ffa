#!/bin/bash
# round 2: the single-GPU bench lines (default C2 line with the CPU leg, the reference arm, C3 / C4 / C5 at full size)
mkdir -p gpurun_out/r02; cd /root/repo; O=gpurun_out/r02
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > $O/n1_reference_arm.json 2> $O/n1_reference_arm.err; echo "ref rc=$?"
timeout 900 python bench.py > $O/n1_C2_full.json 2> $O/n1_C2_full.err; echo "c2 rc=$?"
for c in C3 C4 C5; do timeout 900 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > $O/n1_$c.json 2> $O/n1_$c.err; echo "$c rc=$?"; done
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02/n1_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.1f e2e %.1f golden %s ms/step %.2f" % (d["value"], d["e2e"]["value"], d.get("golden"), d["ms_per_step"]), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
P
