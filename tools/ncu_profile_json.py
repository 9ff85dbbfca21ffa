#!/usr/bin/env python3
"""Distil one `ncu --set full` capture of the trace kernel into the JSON bench.py reads (profiles/kernel_profile.json).
usage: ncu_profile_json.py <rep> <config-name> <out.json> [kernel_info json from render_once]
bench.py only reports these numbers when registers / shared memory / grid of the RUNNING build equal the captured ones."""
import csv, json, subprocess, sys
rep, cfg, out = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, vals = rows[0], rows[1], rows[2]
def g(name, conv=float):
    i = hdr.index(name); v = vals[i].replace(",", "")
    return conv(v), units[i]
def bytes_of(name):
    v, u = g(name)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines())); sh = srows[1]; sd = srows[2:]
ismp = sh.index("# Samples")
ibar = [i for i, h in enumerate(sh) if h == "stall_barrier"]
tot = sum(int(r[ismp]) for r in sd)
bar = sum(int(r[ibar[0]] or 0) for r in sd) if ibar else 0
dur, du = g("gpu__time_duration.sum")
dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(du, 1.0)
smem, su = g("launch__shared_mem_per_block_dynamic")
smem_b = smem * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6}.get(su.split("/")[0], 1)
d = {
    "config": cfg, "kernel": vals[hdr.index("Kernel Name")], "source": rep.split("/")[-1],
    "registers": int(g("launch__registers_per_thread")[0]), "grid": int(g("launch__grid_size")[0]), "block": int(g("launch__block_size")[0]),
    "smem_bytes": int(round(smem_b)), "duration_ms_under_ncu": dur_ms,
    "dram_bytes_per_launch": bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum"),
    "dram_pct_of_peak": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")[0],
    "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active")[0],
    "lanes_per_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio")[0],
    "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active")[0],
    "inst_executed": g("smsp__inst_executed.sum")[0],
    "barrier_stall_pct": 100.0 * bar / max(tot, 1),
    "fp64_pipe_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")[0],
    "fma_pipe_pct": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")[0],
    "lsu_pipe_pct": g("sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active")[0],
}
db = {}
try:
    db = json.load(open(out))
except Exception:
    pass
db[cfg] = d
json.dump(db, open(out, "w"), indent=1, sort_keys=True)
print(json.dumps(d, indent=1))
