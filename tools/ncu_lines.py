#!/usr/bin/env python3
"""Attribute an ncu source-page export (SASS, per-instruction counts) to CUDA source lines using nvdisasm line info.
usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top]"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(lib)}", shell=True, capture_output=True)
lines_of = None
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur = None; infn = False; seq = []
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
        if m:
            if infn and seq: break
            infn = kname in m.group(1); seq = []; cur = None; continue
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            inl = "inlined" in m.group(3)
            if not inl: cur = (os.path.basename(m.group(1)), int(m.group(2)))
            else: cur_inl = (os.path.basename(m.group(1)), int(m.group(2)))
            if inl: cur = cur  # keep outermost non-inlined location when present
            last = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            seq.append(last if 'last' in dir() else None)
    if infn and seq:
        lines_of = seq; break
csvout = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvout.splitlines()))
hdr = rows[1]; data = rows[2:]
iex = hdr.index("Instructions Executed"); ismp = hdr.index("# Samples"); ithr = hdr.index("Thread Instructions Executed")
print("sass instrs:", len(data), "line-info instrs:", len(lines_of) if lines_of else None)
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for i, r in enumerate(data):
    key = lines_of[i] if lines_of and i < len(lines_of) else None
    for k, c in enumerate((iex, ismp, ithr)):
        v = int(r[c]); agg[key][k] += v; tot[k] += v
src_cache = {}
def srcline(key):
    if not key: return ""
    f, l = key
    for root in ("rust-raytracer_b200/csrc",):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip()[:90] if l - 1 < len(src_cache[p]) else ""
    return ""
print(f"{'file:line':34s} {'instr%':>7s} {'smp%':>6s} {'thr/inst':>8s}  source")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    nm = f"{key[0]}:{key[1]}" if key else "?"
    print(f"{nm:34s} {100*v[0]/tot[0]:7.2f} {100*v[1]/tot[1]:6.2f} {v[2]/max(v[0],1):8.1f}  {srcline(key)}")
