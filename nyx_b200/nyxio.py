"""Small on-disk formats either side of the HeatCool path; numpy only, used by tests / tools.

1. AMReX VisMF (MultiFab on disk), header version 1 -- what Nyx::init_zhi reads with VisMF::Read(zhi_from_file, nyx.inhomo_zhi_file)
   (Source/Initialization/Nyx_initdata.cpp:186-191).  `<name>_H`:
       version(1) / how(1 = NFiles) / ncomp / ngrow  [one integer, or an IntVect "(g,g,g)" in newer AMReX]
       BoxArray::writeOn:  "(n 0" / "((lox,loy,loz) (hix,hiy,hiz) (0,0,0))" x n / ")"
       n / "FabOnDisk: <data file> <byte offset>" x n
       blank / "n,ncomp" / n rows of per-component minima "v,v,...," / blank / "n,ncomp" / n rows of maxima
   and the data files hold FArrayBox::writeOn records (the same native FAB format as the hctest chunks, nyx_b200/hctest.py).
   Pinned on a VisMF the reference ships (Util/SliceUtils/slice_00340/Diag_x_*, copied to tests/golden/vismf_slice): offsets, boxes and the
   header's min / max tables against the data (tests/test_nyxio.py).
2. The use_typical_steps checkpoint files `first_max_steps`, `second_max_steps` (written Source/IO/Nyx_output.cpp:295-317, read back by
   Nyx::typical_values_post_restart, Source/Driver/Nyx.cpp:1625-1657): one integer and a newline each.  Quirk kept: the reference writes
   old_max_sundials_steps into BOTH files, and reads the second one into new_max_sundials_steps.
"""
import os
import re

import numpy as np

from . import hctest

_BOX = re.compile(r"\(\((-?\d+),(-?\d+),(-?\d+)\) \((-?\d+),(-?\d+),(-?\d+)\) \((\d+),(\d+),(\d+)\)\)")


def read_vismf(name):
    """name: path without the _H suffix.  -> dict(ncomp, ngrow, boxes [(lo, hi)], fabs [(ncomp, nz, ny, nx) arrays, ghost cells included],
    los [first cell of each FAB], fab_on_disk [(file, offset)], min, max [(nfabs, ncomp) arrays from the header])"""
    lines = open(name + "_H").read().split("\n")
    it = iter(lines)
    version = int(next(it))
    if version != 1:
        raise ValueError(f"{name}_H: VisMF header version {version} (only version 1, the one with min/max tables, is handled)")
    how = int(next(it))
    ncomp = int(next(it))
    g = next(it).strip()
    ngrow = int(g) if g.lstrip("-").isdigit() else int(re.match(r"\((-?\d+),", g).group(1))
    m = re.match(r"\((\d+) (\d+)", next(it))
    nboxes = int(m.group(1))
    boxes = []
    for _ in range(nboxes):
        v = [int(x) for x in _BOX.match(next(it).strip()).groups()]
        if v[6:] != [0, 0, 0]:
            raise ValueError("only cell-centred MultiFabs are handled")
        boxes.append((tuple(v[0:3]), tuple(v[3:6])))
    if next(it).strip() != ")":
        raise ValueError("BoxArray not closed")
    nfod = int(next(it))
    fod = []
    for _ in range(nfod):
        tag, fname, off = next(it).split()
        if tag != "FabOnDisk:":
            raise ValueError(tag)
        fod.append((fname, int(off)))

    def table():
        ln = next(it)
        while not ln.strip():
            ln = next(it)
        n, nc = (int(x) for x in ln.split(","))
        return np.array([[float(x) for x in next(it).rstrip(",").split(",")] for _ in range(n)]).reshape(n, nc)
    mn, mx = table(), table()
    d = os.path.dirname(name)
    cache, fabs, los = {}, [], []
    for (fname, off), (lo, hi) in zip(fod, boxes):
        buf = cache.setdefault(fname, open(os.path.join(d, fname), "rb").read())
        arr, flo, _ = hctest.read_fab(buf, off)
        want = tuple(h - l + 1 + 2 * ngrow for l, h in zip(lo, hi))
        if arr.shape != (ncomp,) + want[::-1] or flo != tuple(x - ngrow for x in lo):
            raise ValueError(f"FAB at {fname}:{off} is {arr.shape} from {flo}, the header says box {lo}..{hi} grown by {ngrow}")
        fabs.append(arr)
        los.append(flo)
    return dict(version=version, how=how, ncomp=ncomp, ngrow=ngrow, boxes=boxes, fabs=fabs, los=los, fab_on_disk=fod, min=mn, max=mx)


def write_vismf(name, boxes, fabs, ngrow=0, nfiles=1):
    """the inverse of read_vismf: fabs[i] covers boxes[i] grown by ngrow; FABs are dealt round-robin over `nfiles` data files"""
    ncomp = fabs[0].shape[0]
    base = os.path.basename(name)
    d = os.path.dirname(name)
    handles = [open(os.path.join(d, f"{base}_D_{i:05d}"), "wb") for i in range(nfiles)]
    fod = []
    for i, ((lo, hi), arr) in enumerate(zip(boxes, fabs)):
        f = handles[i % nfiles]
        fod.append((f"{base}_D_{i % nfiles:05d}", f.tell()))
        hctest.write_fab(f, arr, tuple(x - ngrow for x in lo))
    for f in handles:
        f.close()
    v = (slice(None),) + (slice(ngrow, -ngrow or None),) * 3
    with open(name + "_H", "w") as f:
        f.write(f"1\n1\n{ncomp}\n{ngrow}\n({len(boxes)} 0\n")
        for lo, hi in boxes:
            f.write(f"(({lo[0]},{lo[1]},{lo[2]}) ({hi[0]},{hi[1]},{hi[2]}) (0,0,0))\n")
        f.write(f")\n{len(boxes)}\n")
        for fname, off in fod:
            f.write(f"FabOnDisk: {fname} {off}\n")
        for red in (np.min, np.max):
            f.write(f"\n{len(boxes)},{ncomp}\n")
            for arr in fabs:
                f.write("".join(f"{red(arr[v][c]):.16e}," for c in range(ncomp)) + "\n")
        f.write("\n")


def coarse_zhi_for_box(vm, lo, hi, ratio, comp=0):
    """The part of Nyx::init_zhi before its cell loop, for one fine box: the coarse FAB over (box coarsened by ratio) filled from the file's
    FABs (BoxArray::coarsen + MultiFab::ParallelCopy, Nyx_initdata.cpp:181-191).  -> (array (1, nz, ny, nx), coarse lo)"""
    clo = tuple(int(np.floor(x / ratio)) for x in lo)          # amrex::coarsen rounds towards minus infinity
    chi = tuple(int(np.floor(x / ratio)) for x in hi)
    out = np.full((1,) + tuple(h - l + 1 for l, h in zip(clo, chi))[::-1], np.nan)
    for (blo, bhi), arr, flo in zip(vm["boxes"], vm["fabs"], vm["los"]):
        ilo = tuple(max(a, b) for a, b in zip(clo, blo))
        ihi = tuple(min(a, b) for a, b in zip(chi, bhi))
        if any(a > b for a, b in zip(ilo, ihi)):
            continue
        dst = tuple(slice(ilo[d] - clo[d], ihi[d] - clo[d] + 1) for d in (2, 1, 0))
        src = tuple(slice(ilo[d] - flo[d], ihi[d] - flo[d] + 1) for d in (2, 1, 0))
        out[(0,) + dst] = arr[(comp,) + src]
    if np.isnan(out).any():
        raise ValueError(f"the file does not cover the coarsened box {clo}..{chi}")
    return out, clo


def write_typical_steps(dirname, old_max_sundials_steps):
    """Source/IO/Nyx_output.cpp:295-317 -- both files receive old_max_sundials_steps"""
    for fn in ("first_max_steps", "second_max_steps"):
        with open(os.path.join(dirname, fn), "w") as f:
            f.write(f"{int(old_max_sundials_steps)}\n")


def read_typical_steps(dirname):
    """Nyx::typical_values_post_restart (Source/Driver/Nyx.cpp:1625-1657) -> (old_max_sundials_steps, new_max_sundials_steps)"""
    out = []
    for fn in ("first_max_steps", "second_max_steps"):
        with open(os.path.join(dirname, fn)) as f:
            out.append(int(f.read().split()[0]))
    return tuple(out)
