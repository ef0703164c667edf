#!/usr/bin/env python
"""Instruction mix of one kernel from an ncu source page, by opcode class and by code region (tools/regions.py).
usage: ncu_mix.py <report.ncu-rep> <lib.so> <kernel-substring>"""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from regions import region
rep, lib, kname = sys.argv[1:4]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
linemap, cur, inside = {}, ("?", 0), False
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        inside = kname in ln; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*)", ln)
    if m: linemap[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = rows[1]
iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
def cls(op):
    o = op.split(".")[0]
    if o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"): return "fp64:" + o
    if o.startswith("MUFU"): return "mufu"
    if o in ("F2F", "I2F", "F2I", "FRND", "I2FP", "F2FP", "DMMA"): return "conv"
    if o in ("LDS", "STS", "LDSM"): return "smem"
    if o in ("LDG", "STG", "LD", "ST", "LDL", "STL", "ATOMG", "ATOMS", "RED", "LDC", "ULDC"): return "mem:" + o
    if o in ("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "WARPSYNC", "BAR", "BREAK", "JMP", "BRX", "YIELD", "NOP"): return "ctl:" + o
    if o.startswith("F") : return "fp32"
    return "int/other"
mix = defaultdict(lambda: defaultdict(int)); tot = defaultdict(int); samp = defaultdict(int)
base = None
for r in rows[2:]:
    if not r or r[0] == "Kernel Name": break
    addr = int(r[0], 16)
    if base is None: base = addr
    f, l = linemap.get(addr - base, ("?", 0))
    reg = region(root, f, l)
    src = r[1].strip()
    src = re.sub(r"^@!?U?P\d+\s+", "", src)
    op = src.split()[0] if src else "?"
    n = int(r[iI] or 0)
    grp = "R" if any(k in reg for k in ("fdiv", "ion_", "iterate_ne", "rhs_tail", "eval_request", "amrex_max0", "uvb_rho", "fast_log10", "div_delta_t", "load_request", "save_result", "sm_32_intrinsics", "phase R")) else ("S" if ("phase S" in reg or "sort_key" in reg) else "B")
    mix[grp][cls(op)] += n; tot[grp] += n; samp[grp] += int(r[iS] or 0)
T = sum(tot.values()); ST = sum(samp.values())
thr = defaultdict(int)
for r in rows[2:]:
    if not r or r[0] == "Kernel Name": break
print(f"# instruction mix by phase of {os.path.basename(rep)}: R = evaluation (RHS / EOS), B = bookkeeping (BDF state machine, load / store of cells), S = sort")
for g in sorted(mix):
    print(f"== group {g}: warp-inst {tot[g]} ({100*tot[g]/T:.1f}% of kernel), samples {100*samp[g]/ST:.1f}%")
    for c, n in sorted(mix[g].items(), key=lambda kv: -kv[1])[:14]:
        print(f"   {c:12s} {100*n/tot[g]:5.1f}%")
