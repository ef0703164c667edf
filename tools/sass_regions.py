#!/usr/bin/env python
"""Static SASS size per function / resume() stage of one kernel (16 bytes per instruction): sass_regions.py <lib.so> <kernel-substring>
Out-of-line device functions (ddiv, dsqrt, ion_n, ...) are separate .text sections: pass their name substring to size them."""
import os, re, subprocess, sys, tempfile
from collections import Counter
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from regions import region
lib, kname = sys.argv[1:3]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
cnt, cur, inside, secs = Counter(), ("?", 0), False, Counter()
sec = None
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        sec = ln.split(".text.")[1].split(",")[0]
        inside = kname in ln
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and sec:
        secs[sec] += 1
    if not inside:
        continue
    m2 = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m2:
        cur = (os.path.basename(m2.group(1)), int(m2.group(2))); continue
    if m:
        cnt[region(root, cur[0], cur[1])] += 1
tot = sum(cnt.values())
print(f"kernel {kname}: {tot} instructions = {tot * 16 / 1024:.1f} KB")
for k, v in cnt.most_common(60):
    print(f"  {k:48s} {v:6d}  {v * 16 / 1024:6.1f} KB")
print("sections (all .text of the library):")
for k, v in secs.most_common(30):
    print(f"  {k[:110]:110s} {v:6d} {v * 16 / 1024:6.1f} KB")
