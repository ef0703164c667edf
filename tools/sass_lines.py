#!/usr/bin/env python
"""Static SASS instruction count per source line (and per line range) of one kernel: sass_lines.py <lib.so> <kernel-substring> [top]"""
import os, re, subprocess, sys, tempfile
from collections import Counter
lib, kname = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
cnt, ops, cur, inside = Counter(), Counter(), ("?", 0), False
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        cnt[cur] += 1; ops[m.group(3).split(".")[0]] += 1
tot = sum(cnt.values())
print("total", tot)
print("opcodes:", ", ".join(f"{k}={v}" for k, v in ops.most_common(25)))
for key, v in cnt.most_common(top):
    print(f"{key[0]}:{key[1]:<5d} {v:5d} {100*v/tot:5.1f}%")
