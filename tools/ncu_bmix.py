import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict
sys.path.insert(0, "/root/repo/tools")
from regions import region
rep, lib, kname = sys.argv[1:4]
root = "/root/repo"
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
rows = list(csv.reader(out)); hdr = rows[1]
iI = hdr.index("Instructions Executed")
ops = defaultdict(int); regs = defaultdict(int); tot = 0
base=None
for r in rows[2:]:
    if not r or r[0] == "Kernel Name": break
    addr = int(r[0], 16)
    if base is None: base = addr
    f, l = linemap.get(addr - base, ("?", 0))
    reg = region(root, f, l)
    if any(k in reg for k in ("fdiv", "ion_", "iterate_ne", "rhs_tail", "eval_request", "amrex_max0", "uvb_rho", "fast_log10", "div_delta_t", "load_request", "save_result", "sm_32_intrinsics", "phase R", "phase S", "sort_key")): continue
    src = re.sub(r"^@!?U?P\d+\s+", "", r[1].strip()); op = src.split()[0] if src else "?"
    n = int(r[iI] or 0)
    ops[op.split(".")[0] + ("." + op.split(".")[1] if "." in op and op.split(".")[0] in ("IMAD","MOV","LOP3","SEL","ISETP","IADD3") else "")] += n; regs[reg] += n; tot += n
print("B total", tot)
for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:40]: print(f"  {k:16s} {100*v/tot:5.1f}%")
print("by region")
for k, v in sorted(regs.items(), key=lambda kv: -kv[1])[:30]: print(f"  {k:50s} {100*v/tot:5.1f}%")
