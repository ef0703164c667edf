"""The reference's `hctest` wire format: the isolation-test snapshot that Nyx::integrate_state_struct dumps with
nyx.hctest_example_write = 1 and Exec/HeatCoolTests replays (Source/HeatCool/f_rhs_struct.H:587-697 sdc_writeOn / sdc_readFrom,
Source/HeatCool/integrate_state_with_source_3d.cpp:82-125).  Reader and writer for test / bench harnesses, so that snapshots of real runs
can be pushed through the CUDA path (tools/hctest_replay.py; compared with the oracle in tests/test_hctest_format.py); the dump/replay hooks themselves stay with the reference.

On disk, for step N and MFIter index i:
  <prefix>Chunk.N.i   six FABs back to back -- S_old, D_old, S_new, hydro_src, reset_src, IR (f_rhs_struct.H:639-644) -- each
                      FArrayBox::writeOn in the native binary format (AMReX_FArrayBox.cpp:915-922, AMReX_FabConv.cpp):
                      "FAB ((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))((lox,loy,loz) (hix,hiy,hiz) (0,0,0)) ncomp\\n"
                      followed by prod(hi-lo+1)*ncomp little-endian IEEE doubles, x fastest, component slowest;
  <prefix>BADMAP.N    BoxArray::writeOn + DistributionMapping::writeOn: "(nboxes 0\\n((lo) (hi) (0,0,0))\\n ... )(nboxes\\nrank\\n ... )";
  <prefix>inputs.N    the ParmParse table plus the replay keys nyx.initial_z, nyx.final_z, nyx.fixed_dt, ... ("key = value" lines).
"""
import os
import re

import numpy as np

REAL_DESCRIPTOR = "((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))"   # IEEE double, little endian (RealDescriptor of a native x86-64 build)
FAB_ORDER = ("s_old", "diag", "s_new", "hydro_src", "reset_src", "ir")        # order inside a chunk
_HDR = re.compile(rb"FAB \(\(8, \(64 11 52 0 1 12 0 1023\)\),\(8, \(8 7 6 5 4 3 2 1\)\)\)"
                  rb"\(\((-?\d+),(-?\d+),(-?\d+)\) \((-?\d+),(-?\d+),(-?\d+)\) \((\d+),(\d+),(\d+)\)\) (\d+)\n")


def fab_header(lo, hi, ncomp):
    return f"FAB {REAL_DESCRIPTOR}(({lo[0]},{lo[1]},{lo[2]}) ({hi[0]},{hi[1]},{hi[2]}) (0,0,0)) {ncomp}\n".encode()


def write_fab(f, arr, lo):
    """arr: (ncomp, nz, ny, nx) float64 whose first cell is `lo`."""
    arr = np.ascontiguousarray(arr, dtype="<f8")
    ncomp, nz, ny, nx = arr.shape
    f.write(fab_header(lo, (lo[0] + nx - 1, lo[1] + ny - 1, lo[2] + nz - 1), ncomp))
    f.write(arr.tobytes())


def read_fab(buf, pos):
    """-> (array (ncomp, nz, ny, nx), lo, new position); raises ValueError on anything but a native little-endian double FAB."""
    m = _HDR.match(buf, pos)
    if not m:
        raise ValueError(f"no native double-precision FAB header at byte {pos}: {bytes(buf[pos:pos + 80])!r}")
    lox, loy, loz, hix, hiy, hiz, t0, t1, t2, ncomp = (int(x) for x in m.groups())
    if (t0, t1, t2) != (0, 0, 0):
        raise ValueError("only cell-centred FABs occur in a hctest chunk")
    nx, ny, nz = hix - lox + 1, hiy - loy + 1, hiz - loz + 1
    n = nx * ny * nz * ncomp
    start = m.end()
    if start + 8 * n > len(buf):
        raise ValueError("truncated FAB")
    arr = np.frombuffer(buf, dtype="<f8", count=n, offset=start).reshape(ncomp, nz, ny, nx).copy()
    return arr, (lox, loy, loz), start + 8 * n


def write_chunk(path, fabs, los):
    """fabs / los: dicts keyed by FAB_ORDER."""
    with open(path, "wb") as f:
        for k in FAB_ORDER:
            write_fab(f, fabs[k], los[k])


def _read_bytes(path):
    """the file, or its xz-compressed copy `path`.xz (how committed fixtures are stored)"""
    if not os.path.exists(path) and os.path.exists(path + ".xz"):
        import lzma
        with lzma.open(path + ".xz", "rb") as f:
            return f.read()
    return open(path, "rb").read()


def read_fabs(path, names):
    """a file of len(names) FABs back to back -> (fabs, los) keyed by `names`"""
    buf = _read_bytes(path)
    pos, fabs, los = 0, {}, {}
    for k in names:
        fabs[k], los[k], pos = read_fab(buf, pos)
    if pos != len(buf):
        raise ValueError(f"{path}: {len(buf) - pos} trailing bytes after the {len(names)} FABs")
    return fabs, los


def read_chunk(path):
    buf = _read_bytes(path)
    pos, fabs, los = 0, {}, {}
    for k in FAB_ORDER:
        fabs[k], los[k], pos = read_fab(buf, pos)
    if pos != len(buf):
        raise ValueError(f"{path}: {len(buf) - pos} trailing bytes after the six FABs")
    return fabs, los


def write_badmap(path, boxes, ranks=None):
    ranks = ranks or [0] * len(boxes)
    with open(path, "w") as f:
        f.write(f"({len(boxes)} 0\n")
        for lo, hi in boxes:
            f.write(f"(({lo[0]},{lo[1]},{lo[2]}) ({hi[0]},{hi[1]},{hi[2]}) (0,0,0))\n")
        f.write(f")({len(boxes)}\n")
        for r in ranks:
            f.write(f"{r}\n")
        f.write(")")


def read_badmap(path):
    txt = open(path).read()
    m = re.match(r"\((\d+) \d+\n", txt)
    if not m:
        raise ValueError(f"{path}: not a BoxArray::writeOn stream")
    n = int(m.group(1))
    boxes = [tuple((int(a), int(b), int(c)) for a, b, c in (g[0:3], g[3:6]))
             for g in re.findall(r"\(\((-?\d+),(-?\d+),(-?\d+)\) \((-?\d+),(-?\d+),(-?\d+)\) \(0,0,0\)\)", txt)]
    if len(boxes) != n:
        raise ValueError(f"{path}: {len(boxes)} boxes, header says {n}")
    tail = txt[txt.rindex(")("):]
    ranks = [int(x) for x in tail[2:-1].split()[1:]]
    return boxes, ranks


def write_inputs(path, table):
    with open(path, "w") as f:
        for k, v in table.items():
            f.write(f"{k} = {v}\n")


def read_inputs(path):
    """"key = value" lines of a ParmParse dump (later entries win, as in ParmParse)."""
    out = {}
    for ln in open(path):
        ln = ln.split("#", 1)[0].strip()
        if "=" in ln:
            k, v = ln.split("=", 1)
            out[k.strip()] = v.strip()
    return out


def write_fixture(dirname, step, boxes, chunks, inputs, prefix=""):
    """chunks: one (fabs, los) per box, in MFIter order.  Writes what sdc_writeOn writes."""
    os.makedirs(dirname, exist_ok=True)
    base = os.path.join(dirname, prefix)
    table = dict(inputs)
    table.update({"nyx.hctest_filename_inputs": f"{base}inputs.{step}", "nyx.hctest_filename_badmap": f"{base}BADMAP.{step}",
                  "nyx.hctest_filename_chunk": f"{base}Chunk.{step}.", "nyx.hctest_endIndex": len(boxes), "nyx.hctest_example_write": 0,
                  "nyx.hctest_example_read": 1, "nyx.do_dm_particles": 0, "nyx.do_hydro": 0, "nyx.hctest_example_index": step})
    write_inputs(f"{base}inputs.{step}", table)
    write_badmap(f"{base}BADMAP.{step}", boxes)
    for i, (fabs, los) in enumerate(chunks):
        write_chunk(f"{base}Chunk.{step}.{i}", fabs, los)


def read_fixture(dirname, step, prefix=""):
    base = os.path.join(dirname, prefix)
    inputs = read_inputs(f"{base}inputs.{step}")
    boxes, ranks = read_badmap(f"{base}BADMAP.{step}")
    chunks = [read_chunk(f"{base}Chunk.{step}.{i}") for i in range(len(boxes))]
    for (lo, hi), (fabs, los) in zip(boxes, chunks):   # every FAB contains its valid box
        for k in FAB_ORDER:
            nz, ny, nx = fabs[k].shape[1:]
            l = los[k]
            if not all(l[d] <= lo[d] and hi[d] <= l[d] + (nx, ny, nz)[d] - 1 for d in range(3)):
                raise ValueError(f"FAB {k} does not cover box {lo}-{hi}")
    return dict(inputs=inputs, boxes=boxes, ranks=ranks, chunks=chunks,
                z=float(inputs["nyx.initial_z"]), z_end=float(inputs["nyx.final_z"]), dt=float(inputs["nyx.fixed_dt"]))


def params_from_inputs(inputs):
    """The nyx.* keys of the path (ode_eos_setup, f_rhs_struct.H:45-101; Nyx.cpp:474-568) -> keyword arguments of HcParams."""
    g = lambda k, d: float(inputs.get(k, d))   # noqa: E731
    kw = dict(rtol=g("nyx.sundials_reltol", 1e-4), atol_factor=g("nyx.sundials_abstol", 1e-4), h_species=g("nyx.h_species", 0.76),
              gamma_minus_1=g("nyx.gamma", 5.0 / 3.0) - 1.0, uvb_density_A=g("nyx.uvb_density_A", 1.0), uvb_density_B=g("nyx.uvb_density_B", 0.0),
              zhi_flash=g("nyx.reionization_zHI_flash", -1.0), zheii_flash=g("nyx.reionization_zHeII_flash", -1.0),
              T_zhi=g("nyx.reionization_T_zHI", 0.0), T_zheii=g("nyx.reionization_T_zHeII", 0.0))
    kw["inhomo_reion"] = int(float(inputs.get("nyx.inhomo_reion", 0)))
    kw["use_constraint"] = int(float(inputs.get("nyx.use_sundials_constraint", 0)))
    kw["use_typical_steps"] = int(float(inputs.get("nyx.use_typical_steps", 0)))
    return kw
