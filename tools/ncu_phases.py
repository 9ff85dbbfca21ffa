#!/usr/bin/env python3
"""Instruction/sample breakdown of a kernel by KERNEL-LEVEL source line ranges (inlined callees folded into their call site).
usage: ncu_phases.py rep lib.so kernel-substr file.cu name:lo-hi [name:lo-hi ...]"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, lib, kname, cufile = sys.argv[1:5]
phases = []
for a in sys.argv[5:]:
    nm, r = a.split(':'); lo, hi = r.split('-'); phases.append((nm, int(lo), int(hi)))
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(lib)}", shell=True, capture_output=True)
outer = None
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    infn = False; seq = []; last = None
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
        if m:
            if infn and seq: break
            infn = kname in m.group(1); seq = []; last = None; continue
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: last = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln): seq.append(last)
    if infn and seq: outer = seq; break
csvout = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvout.splitlines())); hdr = rows[1]; data = rows[2:]
iex = hdr.index("Instructions Executed"); ismp = hdr.index("# Samples"); ithr = hdr.index("Thread Instructions Executed")
stall = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
assert outer and len(outer) == len(data), (len(outer) if outer else None, len(data))
agg = collections.OrderedDict((nm, [0, 0, 0, 0, collections.Counter()]) for nm, _, _ in phases); agg["other"] = [0, 0, 0, 0, collections.Counter()]
tot = [0, 0]
base = os.path.basename(cufile)
for i, r in enumerate(data):
    key = "other"
    o = outer[i]
    if o and o[0] == base:
        for nm, lo, hi in phases:
            if lo <= o[1] <= hi: key = nm; break
    a = agg[key]; a[0] += int(r[iex]); a[1] += int(r[ismp]); a[2] += int(r[ithr]); a[3] += 1
    for h, c in stall.items():
        if r[c] not in ("", "0"): a[4][h] += int(r[c])
    tot[0] += int(r[iex]); tot[1] += int(r[ismp])
print(f"{'phase':12s} {'sass':>5s} {'instr%':>7s} {'smp%':>6s} {'thr/inst':>8s}  top stalls")
for nm, a in agg.items():
    top = ", ".join(f"{k[6:]}={100*v/max(a[1],1):.0f}%" for k, v in a[4].most_common(4))
    print(f"{nm:12s} {a[3]:5d} {100*a[0]/tot[0]:7.2f} {100*a[1]/tot[1]:6.2f} {a[2]/max(a[0],1):8.1f}  {top}")
print("total warp instrs", tot[0])
