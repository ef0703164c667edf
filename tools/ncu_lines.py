#!/usr/bin/env python
"""Per-source-line aggregation of an ncu source page (SASS view) using nvdisasm line info.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-name-substring> [top]
Prints, per file:line, the stall samples, warp instructions executed and average active threads.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# offsets -> (file, line) inside the chosen kernel
linemap, cur, inside = {}, ("?", 0), False
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*)", ln)
    if m:
        linemap[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
# several kernels may be in the report: take the first whose name matches
start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kname.replace("ILi0", "").split("IL")[0] in r[1].replace(" ", ""))
hdr = rows[start + 1]
iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
base = None
tot = [0, 0, 0]
for r in rows[start + 2:]:
    if not r or r[0] == "Kernel Name":
        break
    addr = int(r[0], 16)
    if base is None:
        base = addr
    key = linemap.get(addr - base, (("?", 0), ""))[0]
    a = agg[key]
    s, i, t = int(r[iS] or 0), int(r[iI] or 0), int(r[iT] or 0)
    a[0] += s; a[1] += i; a[2] += t
    tot[0] += s; tot[1] += i; tot[2] += t
    for c in stall_cols:
        v = int(r[c] or 0)
        if v:
            a[3][hdr[c]] += v
print(f"total samples {tot[0]}  warp-inst {tot[1]}  avg active threads {tot[2] / max(tot[1], 1):.2f}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{k[6:]}={v}" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:4])
    print(f"{key[0]}:{key[1]:<5d} samples {a[0]:7d} ({100 * a[0] / tot[0]:5.1f}%)  inst {a[1]:11d} ({100 * a[1] / tot[1]:5.1f}%)  act {a[2] / max(a[1], 1):5.1f}  {st}")
# ---- per-function summary (tools/regions.py maps file:line to the enclosing function / resume() stage)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from regions import region
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ra = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
for key, a in agg.items():
    name = region(root, key[0], key[1])
    ra[name][0] += a[0]; ra[name][1] += a[1]; ra[name][2] += a[2]
    for k2, v2 in a[3].items():
        ra[name][3][k2] += v2
print("---- functions (source files as they are NOW: re-run right after profiling)")
for nm, a in sorted(ra.items(), key=lambda kv: -kv[1][1])[:45]:
    st = ", ".join(f"{k[6:]}={100 * v / max(a[0], 1):.0f}%" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:4])
    print(f"{nm:44s} samples {100 * a[0] / tot[0]:5.1f}%  warp-inst {100 * a[1] / tot[1]:5.1f}%  act {a[2] / max(a[1], 1):5.1f}  {st}")
