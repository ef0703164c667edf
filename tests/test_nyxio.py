"""the on-disk formats around the inhomogeneous-reionization input and the use_typical_steps restart state
(nyx_b200/nyxio.py), pinned on a VisMF MultiFab that the reference itself ships (Util/SliceUtils/slice_00340/Diag_x_*, a copy lives in
tests/golden/vismf_slice), and the init_zhi cell loop: oracle on the CPU, CUDA kernel on the GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from nyx_b200 import capi, nyxio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "vismf_slice", "Diag_x")


def test_vismf_reader_on_the_reference_fixture():
    vm = nyxio.read_vismf(FIXTURE)
    assert (vm["version"], vm["how"], vm["ncomp"], vm["ngrow"]) == (1, 1, 2, 0) and len(vm["boxes"]) == 16
    assert vm["boxes"][0] == ((16, 0, 0), (16, 7, 7)) and vm["boxes"][-1] == ((16, 24, 24), (16, 31, 31))
    assert vm["fab_on_disk"][1] == ("Diag_x_D_00001", 1112) and vm["fab_on_disk"][-1] == ("Diag_x_D_00007", 3348)
    # the header's min / max tables were computed by the reference from the same data: every FAB, both components, to the printed digits
    for i, arr in enumerate(vm["fabs"]):
        assert arr.shape == (2, 8, 8, 1)
        for c in range(2):
            assert float(f"{arr[c].min():.16e}") == vm["min"][i, c] and float(f"{arr[c].max():.16e}") == vm["max"][i, c]
    # every byte of the data files is accounted for by exactly one FAB record
    sizes = {}
    for (fname, off), arr in zip(vm["fab_on_disk"], vm["fabs"]):
        sizes.setdefault(fname, []).append((off, arr.size * 8))
    for fname, recs in sizes.items():
        recs.sort()
        total = os.path.getsize(os.path.join(os.path.dirname(FIXTURE), fname))
        ends = [recs[i + 1][0] for i in range(len(recs) - 1)] + [total]
        for (off, nbytes), end in zip(recs, ends):
            hdr = end - off - nbytes
            assert 70 < hdr < 110          # the FAB header line


def test_vismf_round_trip(tmp_path):
    vm = nyxio.read_vismf(FIXTURE)
    # byte-identical rewrite of the reference's files: same header text, same data files (4 files, FABs dealt as the header says)
    name = str(tmp_path / "Diag_x")
    by_file = {}
    for (fname, off), arr, lo in zip(vm["fab_on_disk"], vm["fabs"], vm["los"]):
        by_file.setdefault(fname, []).append((off, arr, lo))
    for fname, recs in by_file.items():
        with open(tmp_path / fname, "wb") as f:
            for off, arr, lo in sorted(recs, key=lambda r: r[0]):
                assert f.tell() == off
                nyxio.hctest.write_fab(f, arr, lo)
        assert open(tmp_path / fname, "rb").read() == open(os.path.join(os.path.dirname(FIXTURE), fname), "rb").read()
    # our own writer -> reader
    rng = np.random.default_rng(5)
    boxes = [((0, 0, 0), (7, 3, 3)), ((8, 0, 0), (15, 3, 3)), ((0, 4, 0), (15, 7, 3))]
    fabs = [rng.standard_normal((1,) + tuple(h - l + 1 + 2 for l, h in zip(lo, hi))[::-1]) for lo, hi in boxes]
    nyxio.write_vismf(str(tmp_path / "zhi"), boxes, fabs, ngrow=1, nfiles=2)
    back = nyxio.read_vismf(str(tmp_path / "zhi"))
    assert back["boxes"] == boxes and back["ngrow"] == 1 and all(np.array_equal(a, b) for a, b in zip(back["fabs"], fabs))
    for i, arr in enumerate(fabs):
        assert back["min"][i, 0] == float(f"{arr[0, 1:-1, 1:-1, 1:-1].min():.16e}")
    # header text of the writer == the reference's for the fixture's own contents
    nyxio.write_vismf(str(tmp_path / "again"), vm["boxes"], vm["fabs"], ngrow=0, nfiles=1)
    ours = open(str(tmp_path / "again_H")).read().split("\n")
    theirs = open(FIXTURE + "_H").read().split("\n")
    assert len(ours) == len(theirs)
    for a, b in zip(ours, theirs):
        assert a == b or a.startswith("FabOnDisk:")


def _zhi_case(tmp_path, ratio=4):
    """a coarse z_HI field on disk (two boxes), a fine level of three ragged boxes with one ghost cell in diag"""
    rng = np.random.default_rng(9)
    cboxes = [((0, 0, 0), (3, 7, 7)), ((4, 0, 0), (7, 7, 7))]
    cfabs = [rng.uniform(5.5, 12.0, (1, 8, 8, 4)) for _ in cboxes]
    nyxio.write_vismf(str(tmp_path / "zhi.bin"), cboxes, cfabs)
    vm = nyxio.read_vismf(str(tmp_path / "zhi.bin"))
    fine = [((0, 0, 0), (15, 31, 31)), ((16, 0, 0), (31, 15, 31)), ((16, 16, 0), (31, 31, 31))]
    diag = [rng.standard_normal((3,) + tuple(h - l + 3 for l, h in zip(lo, hi))[::-1]) for lo, hi in fine]
    coarse = [nyxio.coarse_zhi_for_box(vm, lo, hi, ratio) for lo, hi in fine]
    return vm, fine, diag, coarse, ratio


def test_init_zhi_oracle(tmp_path, port):
    vm, fine, diag, coarse, ratio = _zhi_case(tmp_path)
    full = np.concatenate([f[0] for f in vm["fabs"]], axis=2)          # the coarse field over the whole domain, (nz, ny, nx)
    for (lo, hi), d, (z, zlo) in zip(fine, diag, coarse):
        before = d.copy()
        port.init_zhi(d, tuple(x - 1 for x in lo), z, zlo, lo, hi, ratio)
        k, j, i = np.meshgrid(*(np.arange(lo[a], hi[a] + 1) for a in (2, 1, 0)), indexing="ij")
        assert np.array_equal(d[2, 1:-1, 1:-1, 1:-1], full[k // ratio, j // ratio, i // ratio])
        d2 = d.copy(); d2[2, 1:-1, 1:-1, 1:-1] = before[2, 1:-1, 1:-1, 1:-1]
        assert np.array_equal(d2, before)                               # nothing else touched


@pytest.mark.gpu
def test_init_zhi_on_device(tmp_path, hc_lib, port):
    import torch
    vm, fine, diag, coarse, ratio = _zhi_case(tmp_path)
    dd = [torch.from_numpy(d).cuda() for d in diag]
    zd = [torch.from_numpy(z).cuda() for z, _ in coarse]
    hc_lib.init_zhi_batch([capi.fab_of_torch(d, tuple(x - 1 for x in lo)) for d, (lo, hi) in zip(dd, fine)],
                          [capi.fab_of_torch(z, zlo) for z, (_, zlo) in zip(zd, coarse)], ratio, [capi.make_box(lo, hi) for lo, hi in fine])
    torch.cuda.synchronize()
    for (lo, hi), d, (z, zlo), got in zip(fine, diag, coarse, dd):
        port.init_zhi(d, tuple(x - 1 for x in lo), z, zlo, lo, hi, ratio)
        assert np.array_equal(got.cpu().numpy(), d)
    with pytest.raises(capi.HcError, match="does not cover"):
        hc_lib.init_zhi_batch([capi.fab_of_torch(dd[0], (-1, -1, -1))], [capi.fab_of_torch(zd[0], (1, 0, 0))], ratio, [capi.make_box(*fine[0])])


def test_typical_steps_files(tmp_path, built):
    """python mirror and the C++ drop-in functions write / read the same files; the reference's quirk (old_max in both) is kept"""
    nyxio.write_typical_steps(str(tmp_path), 17)
    assert open(tmp_path / "first_max_steps").read() == "17\n" and open(tmp_path / "second_max_steps").read() == "17\n"
    assert nyxio.read_typical_steps(str(tmp_path)) == (17, 17)
    lib = C.CDLL(built.build_dropin_check())
    lib.nyxref_set.argtypes = [C.c_char_p, C.c_char_p]
    lib.nyxref_get_max_steps.restype = C.c_long
    d2 = tmp_path / "chk"
    d2.mkdir()
    lib.nyxref_set(b"nyx.use_typical_steps", b"0")
    assert lib.nyxref_write_typical_steps(str(d2).encode()) == 0 and not (d2 / "first_max_steps").exists()
    lib.nyxref_set(b"nyx.use_typical_steps", b"1")
    lib.nyxref_set(b"nyx.old_max_sundials_steps", b"23")
    lib.nyxref_set(b"nyx.new_max_sundials_steps", b"41")
    assert lib.nyxref_write_typical_steps(str(d2).encode()) == 0
    assert open(d2 / "first_max_steps").read() == "23\n" and open(d2 / "second_max_steps").read() == "23\n"
    open(d2 / "second_max_steps", "w").write("29\n")
    lib.nyxref_set(b"nyx.old_max_sundials_steps", b"3")
    assert lib.nyxref_read_typical_steps(str(d2).encode()) == 0
    assert (lib.nyxref_get_max_steps(0), lib.nyxref_get_max_steps(1)) == (23, 29) == nyxio.read_typical_steps(str(d2))
    assert lib.nyxref_read_typical_steps(str(tmp_path / "missing").encode()) == -1
    lib.nyxref_set(b"nyx.use_typical_steps", b"0")
