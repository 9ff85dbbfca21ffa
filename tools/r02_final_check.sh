#!/bin/bash
# what the driver runs at round end, on one GPU: the GPU tests, smoke(), the default bench line
mkdir -p gpurun_out/r02; cd /root/repo; O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q > $O/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/final_smoke.log
timeout 900 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02/final_bench.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f golden %s clocks %s cpu %.1f (%d threads)" % (d["value"], d["e2e"]["value"], d["golden"], d["clocks"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]))
print({k: d["roofline"][k] for k in ("frac","traffic","issue_active_pct","lanes_per_inst","barrier_stall_pct","profile")})
P
