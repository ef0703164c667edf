"""Maps file:line to the enclosing function name (heuristic parser for hc_device.cuh / nyx_hc.cu)."""
import bisect, os, re
_cache = {}
def _load(path):
    starts, names = [], []
    pat = re.compile(r'^\s*(?:template\s*<[^>]*>\s*)?(?:HC_HD_NOINLINE|HC_HD|__device__ __forceinline__|__global__|static|inline)\b.*?([A-Za-z_][A-Za-z_0-9]*)\s*\(')
    for i, ln in enumerate(open(path), 1):
        m = pat.match(ln)
        if m and not ln.strip().startswith("//"):
            starts.append(i); names.append(m.group(1))
        m2 = re.match(r'\s*// ================= (.*)', ln)
        if m2:
            starts.append(i); names.append("resume:" + m2.group(1)[:28])
    return starts, names
def region(root, fname, line):
    cands = [os.path.join(root, "nyx_b200", "csrc", fname)]
    for c in cands:
        if os.path.exists(c):
            if c not in _cache: _cache[c] = _load(c)
            st, nm = _cache[c]
            k = bisect.bisect_right(st, line) - 1
            return f"{fname.split('.')[0][:6]}:{nm[k]}" if k >= 0 else fname
    return fname
