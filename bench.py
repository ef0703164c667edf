#!/usr/bin/env python
"""bench.py -- HeatCool cell-updates/s on synthetic Lyman-alpha fields (BASELINE.json metric).

One "step" = one pass of the hot path over the rank's boxes of a 512^3-per-GPU field (lognormal density, log-uniform 1e3..1e7 K, z = 3,
TREECOOL_middle), through the C-ABI of include/nyx_hc.h.  The headline (`value`, `e2e`, `roofline`) is the Strang half-step
Nyx::integrate_state_vec (BASELINE.json's metric names integrate_state_vode); the same line carries, under `paths`, BOTH entry points --
`vec` and `struct` (the SDC step Nyx::integrate_state_struct, what the shipped Exec/LyA inputs run) -- each with value, roofline and e2e, and
for N > 1 under `strong` the SAME global 512^3 field sharded by box over the N GPUs (weak scaling stays the headline so that SCALE records
remain comparable).  `value` = device-resident throughput (CUDA events on the launching stream, max over ranks), `e2e` = the same call on
pinned HOST FABs (H2D + kernel + D2H inside the timed region), `roofline` = algorithmic FP64 flops / event time against the DFMA peak
measured in the same run (plus HBM GB/s), `cpu_baseline` = the reference's own OpenMP implementation (oracle/_ref, else the C port) on a
bounded sample of the same boxes.  `--impl reference` times only that CPU arm.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout.  Libraries write banners to file descriptor 1 (NCCL prints "NCCL version ..." there when
# NCCL_DEBUG is set): keep a private handle on the real stdout for the result line and point fd 1 at stderr for everything else.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

METRIC = "heatcool_cell_updates_per_s"
UNIT = "cell-updates/s"
TREECOOL = os.path.join(ROOT, "tests", "golden", "TREECOOL_middle")   # the reference's Exec/LyA/TREECOOL_middle (data file)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=512, help="cells per side of the (per-GPU) domain")
    ap.add_argument("--box", type=int, default=128, help="amr.max_grid_size")
    ap.add_argument("--z", type=float, default=3.0)
    ap.add_argument("--path", default="vec", choices=["vec", "struct"], help="the path of the headline numbers (the other one goes under `paths`)")
    ap.add_argument("--paths", default="both", choices=["both", "one"], help="one: measure only --path")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling leg")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workload
def workload_name(args, path=None):
    what = "integrate_state_vec Strang half-step dt/2" if (path or args.path) == "vec" else "integrate_state_struct SDC step dt"
    return f"Exec/LyA {args.n}^3 per GPU synthetic lognormal, z={args.z:g}, max_grid_size {args.box}, {what}"


def config_of(args, path, world, ncell_local=None, nb=None, extra=None):
    """the `config` object: the same keys from the GPU arm and the reference arm"""
    n = args.n
    nboxes = len(range(0, n, args.box)) ** 3
    cfg = {"workload": workload_name(args, path), "cells_per_gpu": ncell_local if ncell_local is not None else n ** 3,
           "boxes_per_gpu": nb if nb is not None else nboxes, "path": path, "z": args.z, "rtol": 1e-4, "atol_factor": 1e-4,
           "parallelism": "boxes sharded over %d GPU(s), no data-path collective, one scalar all-reduce of diagnostics" % world}
    if extra:
        cfg.update(extra)
    return cfg


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and with it the first-touch placement of the pinned host FABs) to the cores of the NUMA node the GPU hangs off:
    with 8 ranks per node the host side of the end-to-end leg otherwise funnels 8 x 10 GB per step through one node."""
    info = {"node": None}
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        if bus is None:
            return info
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        info["node"] = node
        if node < 0:
            return info
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
    except Exception as e:  # noqa: BLE001
        info["error"] = repr(e)[:80]
    return info


def box_seed(global_box_index):
    return 20240601 + int(global_box_index)


def make_box_fields(args, gidx, lo, hi):
    """numpy FABs of one box (no ghost cells): the same generator the parity tests use."""
    from nyx_b200 import synth
    shape = (hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, hi[2] - lo[2] + 1)
    return synth.make_fab(shape, seed=box_seed(gidx), z=args.z)


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md clocks line), through NVML."""

    def __init__(self, index, period=0.2):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.sm_max, self.power = [], set(), None, []
        self.stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost", 0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self.stop_evt.wait(self.period)

    def result(self):
        self.stop_evt.set()
        if self.is_alive():
            self.join(timeout=2.0)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "power_w_max": max(self.power) if self.power else None, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU arm
_HOST_THREADS = None


def cpu_reference_run(args, path, boxes_idx, boxes, budget_s, min_boxes=1, calibrated=None):
    """Times the reference's own CPU implementation (oracle/_ref OpenMP build; C port if it is absent) on as many
    boxes of the workload as fit `budget_s` (at least `min_boxes`).  Returns (cells_per_s, info dict)."""
    from nyx_b200 import synth
    from oracle import pyref
    a = 1.0 / (1.0 + args.z)
    dt = synth.step_dt(args.z)
    kind, cores, ref, port = "reference", 1, None, None
    # all the host threads the box offers (torchrun exports OMP_NUM_THREADS=1 to its children; NYX_REF_THREADS overrides).  Counted BEFORE the
    # library is loaded: with OMP_PROC_BIND set, libgomp pins the calling thread to one core as it initialises
    global _HOST_THREADS
    if _HOST_THREADS is None:
        _HOST_THREADS = len(os.sched_getaffinity(0))
    try:
        ref = pyref.Reference("omp")
        want = int(os.environ.get("NYX_REF_THREADS", "0")) or _HOST_THREADS
        ref.set("omp.num_threads", want)
        cores = ref.max_threads()
    except (FileNotFoundError, OSError):
        kind = "port"
        port = pyref.Port()

    def run_boxes(idx_list):
        fields = [make_box_fields(args, boxes_idx[i], *boxes[i]) for i in idx_list]
        cells = sum(f[0].shape[1] * f[0].shape[2] * f[0].shape[3] for f in fields)
        t0 = time.perf_counter()
        if path == "vec":
            if ref is not None:
                ref.stats_reset()
                ref.integrate_state_vec([boxes[i][0] + boxes[i][1] for i in idx_list], [f[0] for f in fields], [f[1] for f in fields], a, 0.5 * dt)
            else:
                for i, f in zip(idx_list, fields):
                    port.integrate_state_vec(f[0], f[1], boxes[i][0], boxes[i][1], a, 0.5 * dt, want_stats=False)
        else:
            a_end = synth.a_after(args.z, dt)
            for i, f in zip(idx_list, fields):
                s_old, diag = f
                s_new = s_old.copy()
                hs = np.zeros_like(s_old)
                rs = np.zeros((1,) + s_old.shape[1:])
                ir = np.zeros((1,) + s_old.shape[1:])
                if ref is not None:
                    ref.stats_reset()
                    ref.integrate_state_struct([boxes[i][0] + boxes[i][1]], [s_old], [s_new], [diag], [hs], [ir], [rs], a, a_end, dt, 0)
                else:
                    port.integrate_state_struct(s_old, s_new, diag, hs, rs, ir, boxes[i][0], boxes[i][1], a, a_end, dt, 0, want_stats=False)
        return cells, time.perf_counter() - t0

    # calibrate on one box (once per process), then size the sample to the budget
    if calibrated is None:
        c1, t1 = run_boxes([0])
    else:
        c1, t1 = calibrated
    nb = int(max(min_boxes, min(len(boxes), budget_s / max(t1, 1e-3))))
    nb = min(nb, len(boxes))
    if nb > 1 or calibrated is not None:
        cells, t = run_boxes(list(range(nb)))
    else:
        cells, t = c1, t1
    info = {"kind": kind, "cores": cores, "sample": f"{nb} box(es) of the workload ({cells} cells) in {t:.2f} s; first box alone {t1:.2f} s",
            "seconds": t, "cells": cells, "calibrated": (c1, t1)}
    return cells / t, info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # OMP_PROC_BIND=close, all host threads.  Must be in the environment before libgomp initialises (oracle/_ref is
    # loaded below); torchrun's OMP_NUM_THREADS=1 is overridden through omp_set_num_threads in cpu_reference_run.
    global _HOST_THREADS
    _HOST_THREADS = len(os.sched_getaffinity(0))
    os.environ["OMP_PROC_BIND"] = "close"
    os.environ.pop("OMP_NUM_THREADS", None)
    from nyx_b200 import sharded
    boxes = sharded.box_list(args.n, args.box)
    idx = list(range(len(boxes)))
    # each timed step: >= 10 s of the reference's work (or >= 8 boxes), the whole run bounded to a few minutes
    per_step = min(12.0, max(4.0, 240.0 / max(1, args.steps + args.warmup)))
    vals, info, cal = [], None, None
    for s in range(args.warmup + args.steps):
        v, info = cpu_reference_run(args, args.path, idx, boxes, per_step, min_boxes=6, calibrated=cal)
        cal = info["calibrated"]
        if s >= args.warmup:
            vals.append(v)
    value = float(np.median(vals)) if vals else 0.0
    cells_per_step = info["cells"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * cells_per_step / value if value else None, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_of(args, args.path, world, extra={"note": "reference CPU/OpenMP implementation (oracle/_ref: the reference's own translation units + CVODE) on a "
                                                               "bounded sample of the workload's boxes per timed step; OMP_PROC_BIND=close"}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"],
                             "min": float(min(vals)) if vals else None, "median": value, "max": float(max(vals)) if vals else None},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
COMPS = {"vec": {"state": 6, "diag": 2}, "struct": {"state": 6, "diag": 2, "s_new": 6, "hydro_src": 6, "reset_src": 1, "ir": 1}}
MUTATED = {"vec": ("state", "diag"), "struct": ("s_new", "diag", "ir")}
# what hc_integrate_*_host moves per cell (components of 8 bytes): Strang: rho, rho_E, rho_e in, rho_E, rho_e, T, Ne out (diag is a pure output)
H2D_COMPS = {"vec": 3, "struct": 2 + 2 + 3 + 2 + 1}      # S_old(rho, rho e), diag(T, Ne), S_new(rho, rho E, rho e), hydro_src(rho, rho e), reset_src; I_R is a pure output
D2H_COMPS = {"vec": 4, "struct": 2 + 1 + 2}


def traffic_from_capture(path, ncells):
    """DRAM read+write bytes of ONE launch of the dominant kernel from the committed `ncu --set full` capture of this command
    (profiles/r2_traffic.json: bytes per launch at a stated cell count); scaled only if the cell count differs."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            t = json.load(f)[path]
        return t["dram_bytes_per_launch"] * (ncells / t["cells"]), t.get("source", "profiles/r2_traffic.json")
    except (OSError, KeyError, ValueError, TypeError):
        return None, None


class PathBench:
    """buffers and timed legs of one entry point (vec | struct) over a list of boxes of this rank"""

    def __init__(self, ctx, path, box_ids, gidx):
        import torch
        from nyx_b200 import capi
        self.ctx, self.path = ctx, path
        args, boxes, dev = ctx["args"], ctx["boxes"], ctx["dev"]
        self.torch = torch
        self.mine, self.gidx = box_ids, gidx
        self.shapes = [tuple(h - l + 1 for l, h in zip(*boxes[i])) for i in box_ids]
        self.ncell = sum(s[0] * s[1] * s[2] for s in self.shapes)
        comps = COMPS[path]
        self.comps = comps

        def alloc(ncomp, pinned):
            tot = sum(ncomp * s[0] * s[1] * s[2] for s in self.shapes)
            return torch.empty(tot, dtype=torch.float64, pin_memory=True) if pinned else torch.empty(tot, dtype=torch.float64, device=dev)

        def views(buf, ncomp):
            out, off = [], 0
            for s in self.shapes:
                n = ncomp * s[0] * s[1] * s[2]
                out.append(buf[off:off + n].view(ncomp, s[2], s[1], s[0]))
                off += n
            return out

        self.host = {k: alloc(c, True) for k, c in comps.items()}
        host_v = {k: views(self.host[k], c) for k, c in comps.items()}
        t0 = time.perf_counter()
        for b, i in enumerate(box_ids):
            st, dg = ctx["field"](gidx[b], *boxes[i])
            host_v["state"][b].copy_(torch.from_numpy(st))
            host_v["diag"][b].copy_(torch.from_numpy(dg))
            if path == "struct":
                host_v["s_new"][b].copy_(torch.from_numpy(st))
        if path == "struct":
            self.host["hydro_src"].zero_(); self.host["reset_src"].zero_(); self.host["ir"].zero_()
        self.t_gen = time.perf_counter() - t0
        self.mut = MUTATED[path]
        self.devb = {k: alloc(c, False) for k, c in comps.items()}
        for k in comps:
            self.devb[k].copy_(self.host[k], non_blocking=True)
        torch.cuda.synchronize()
        dev_v = {k: views(self.devb[k], c) for k, c in comps.items()}
        self.pristine_dev = {k: self.devb[k].clone() for k in self.mut}
        self.pristine_host = None
        los = [boxes[i][0] for i in box_ids]
        self.tiles = [capi.make_box(*boxes[i]) for i in box_ids]
        self.dfab = {k: [capi.fab_of_torch(v, lo) for v, lo in zip(dev_v[k], los)] for k in comps}
        self.hfab = {k: [capi.make_fab(v.data_ptr(), lo, (s[0], s[1], s[2]), comps[k]) for v, lo, s in zip(host_v[k], los, self.shapes)] for k in comps}
        self.stream = torch.cuda.current_stream()

    def step_device(self):
        c, hc, d = self.ctx, self.ctx["hc"], self.dfab
        if self.path == "struct":
            return hc.integrate_struct_batch(d["state"], d["diag"], d["s_new"], d["hydro_src"], d["reset_src"], d["ir"], self.tiles,
                                             c["a"], c["a_end"], c["dt"], 0, stream=self.stream.cuda_stream)
        return hc.integrate_vec_batch(d["state"], d["diag"], self.tiles, c["a"], 0.5 * c["dt"], stream=self.stream.cuda_stream)

    def step_host(self):
        c, hc, h = self.ctx, self.ctx["hc"], self.hfab
        if self.path == "struct":
            return hc.integrate_struct_host(h["state"], h["diag"], h["s_new"], h["hydro_src"], h["reset_src"], h["ir"], self.tiles, c["a"], c["a_end"], c["dt"], 0)
        return hc.integrate_vec_host(h["state"], h["diag"], self.tiles, c["a"], 0.5 * c["dt"])

    def restore_device(self):
        for k in self.mut:
            self.devb[k].copy_(self.pristine_dev[k])

    def restore_host(self):
        if self.pristine_host is None:
            self.pristine_host = {k: self.host[k].clone() for k in self.mut}
        for k in self.mut:
            self.host[k].copy_(self.pristine_host[k])

    def device_leg(self, steps, warmup, sample_clocks=True):
        """-> dict(t_max seconds for `steps` steps (max over ranks), ms_steps, stats of the last step, clocks)"""
        torch, ctx = self.torch, self.ctx
        for _ in range(warmup):
            self.restore_device()
            self.step_device()
        ctx["barrier"]()
        sampler = ClockSampler(ctx["local_rank"]) if sample_clocks else None
        if sampler:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        stats = None
        drains = []
        wall0 = time.perf_counter()
        for s in range(steps):
            self.restore_device()      # untimed: every step integrates the same input (the path updates its FABs in place)
            torch.cuda.synchronize()
            ev[s][0].record(self.stream)
            stats = self.step_device()   # one persistent kernel + the 112-byte statistics read-back
            ev[s][1].record(self.stream)
            drains.append(ctx["hc"].last_launch_timing())
        ctx["barrier"]()
        wall = time.perf_counter() - wall0
        clocks = sampler.result() if sampler else None
        ms_steps = [e0.elapsed_time(e1) for e0, e1 in ev]
        t_local = sum(ms_steps) * 1e-3
        return {"t_local": t_local, "t_max": ctx["allmax"](t_local), "ms_steps": ms_steps, "stats": stats, "clocks": clocks, "wall": wall,
                "kernel_ms": float(np.mean([d[0] for d in drains])), "drain_ms": float(np.mean([d[1] for d in drains]))}

    def roofline(self, leg, steps):
        from nyx_b200 import sharded
        ctx, stats = self.ctx, leg["stats"]
        flops_local = sharded.algorithmic_flops(stats)
        t_kernel = leg["t_local"] / steps
        achieved = flops_local / t_kernel
        bytes_cell = 104 if self.path == "struct" else 56
        hbm_ach = bytes_cell * stats.n_cells / t_kernel / 1e9
        traffic, tsrc = traffic_from_capture(self.path, stats.n_cells)
        peaks = ctx["peaks"]
        return {"bound": "fp64", "achieved": achieved / 1e12, "peak": ctx["fp64_peak"] / 1e12, "unit": "TFLOP/s", "frac": achieved / ctx["fp64_peak"],
                "traffic": traffic, "traffic_source": tsrc, "algorithmic_bytes_per_launch": bytes_cell * stats.n_cells,
                "peak_source": "DFMA peak measured in this run by hc_measure_fp64_peak (MEASURED_PEAKS.json has no FP64 entry)",
                "flops_per_cell": flops_local / stats.n_cells,
                "kernel": "sorted::hc_sorted_kernel<%s, 384>" % ("PATH_STRUCT" if self.path == "struct" else "PATH_VEC"),
                "hbm": {"achieved": hbm_ach, "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s", "frac": hbm_ach / peaks.get("hbm_gbs", 6650.0),
                        "bytes_per_cell": bytes_cell, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"}}

    def e2e_leg(self, steps, ncell_global):
        torch, ctx = self.torch, self.ctx
        h2d = 8 * self.ncell * H2D_COMPS[self.path]
        d2h = 8 * self.ncell * D2H_COMPS[self.path]
        n_warm, n = 1, max(1, min(steps, 3))
        for _ in range(n_warm):
            self.restore_host(); self.step_host()
        ctx["barrier"]()
        t = 0.0
        st_h = None
        for _ in range(n):
            self.restore_host()
            ctx["barrier"]()
            t0 = time.perf_counter()
            st_h = self.step_host()            # returns after the D2H copies have completed (stream-synchronised inside)
            torch.cuda.synchronize()
            t += time.perf_counter() - t0
        t_loc = t
        t = ctx["allmax"](t)
        return {"value": ncell_global * n / t, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": n,
                "ms_per_step": 1e3 * t / n, "h2d_gbs_this_rank": h2d * n / t_loc / 1e9, "d2h_gbs_this_rank": d2h * n / t_loc / 1e9,
                "api": "hc_integrate_%s_host on pinned host FABs (H2D / kernel / D2H pipelined over groups of boxes -- 1/64, 1/32, 1/16, 1/8 ... of the cells, halving again at the end -- with the kernels of consecutive groups on alternating streams)" % self.path,
                "n_failed": st_h.n_failed}

    def free(self):
        self.host = self.devb = self.pristine_dev = self.pristine_host = self.dfab = self.hfab = None
        self.torch.cuda.empty_cache()


def host_mem_available_gb():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) / 1e6
    except OSError:
        pass
    return None


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from nyx_b200 import capi, sharded, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the HeatCool path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    hc = capi.NyxHC()
    hc.tables_upload(hc.tabulate_rates(TREECOOL, synth.mean_rhob()))
    dt = synth.step_dt(args.z)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    boxes = sharded.box_list(args.n, args.box)
    field_cache = {}

    def field(g, lo, hi):   # both paths (and the strong leg on rank 0) integrate the same boxes: generate each once
        if g not in field_cache:
            field_cache[g] = make_box_fields(args, g, lo, hi)
        return field_cache[g]

    ctx = {"args": args, "boxes": boxes, "dev": dev, "hc": hc, "a": 1.0 / (1.0 + args.z), "a_end": synth.a_after(args.z, dt), "dt": dt,
           "barrier": barrier, "allmax": allmax, "local_rank": local_rank, "peaks": peaks, "field": field}
    ctx["fp64_peak"] = hc.measure_fp64_peak()

    # ---- the rank's boxes: weak scaling = every rank owns a full n^3 domain's worth of boxes of a world-times larger field
    if args.scaling == "weak":
        mine = list(range(len(boxes)))
        gidx = [rank * len(boxes) + i for i in mine]
    else:
        mine = sharded.local_boxes(boxes, world, rank)
        gidx = list(mine)

    order = [args.path] + ([p for p in ("vec", "struct") if p != args.path] if args.paths == "both" else [])
    results, head = {}, None
    for path in order:
        pb = PathBench(ctx, path, mine, gidx)
        leg = pb.device_leg(args.steps if path == args.path else max(3, min(args.steps, 5)), args.warmup, sample_clocks=True)
        steps_p = len(leg["ms_steps"])
        gstats = sharded.allreduce_stats(leg["stats"], device=dev)   # the path's only collective: failure / iteration diagnostics
        value = gstats["n_cells"] * steps_p / leg["t_max"]
        roof = pb.roofline(leg, steps_p)
        e2e = None
        if not args.no_e2e:
            need_gb = 8e-9 * pb.ncell * (sum(COMPS[path].values()) + len(MUTATED[path]) * 3) * world
            avail = host_mem_available_gb()
            if avail is not None and need_gb > 0.6 * avail:
                e2e = {"skipped": "pinned host buffers of %d ranks (%.0f GB) against %.0f GB available" % (world, need_gb, avail)}
            else:
                e2e = pb.e2e_leg(args.steps, gstats["n_cells"])
        rec = {"value": value, "unit": UNIT, "ms_per_step": 1e3 * leg["t_max"] / steps_p, "steps": steps_p, "roofline": roof, "e2e": e2e,
               "kernel_ms_this_rank": leg["kernel_ms"], "drain_tail_ms_this_rank": leg["drain_ms"],
               "stats": gstats, "ms_steps": leg["ms_steps"], "clocks": leg["clocks"], "workload": workload_name(args, path)}
        results[path] = rec
        if path == args.path:
            head = dict(rec, pb_ncell=pb.ncell, nb=len(mine), wall=leg["wall"], t_gen=pb.t_gen, t_local=leg["t_local"])
            # ---- CPU baseline (rank 0, N = 1 only): the reference's OpenMP implementation on a bounded sample of the same boxes
        if path != order[-1] or (world > 1 and not args.no_strong and args.scaling == "weak"):
            pb.free()
            del pb

    # ---- strong scaling (N > 1): the SAME global n^3 field (the boxes rank 0 owns in the weak leg) dealt over the N ranks
    strong = None
    if world > 1 and not args.no_strong and args.scaling == "weak":
        strong = {}
        smine = sharded.local_boxes(boxes, world, rank)
        for path in order:
            pbs = PathBench(ctx, path, smine, list(smine))
            leg = pbs.device_leg(max(3, min(args.steps, 5)), args.warmup, sample_clocks=False)
            steps_s = len(leg["ms_steps"])
            gs = sharded.allreduce_stats(leg["stats"], device=dev)
            t_n1 = results[path]["ms_per_step"] if False else None
            # rank 0's weak-leg time IS the N = 1 time of this field (rank 0 owns global boxes 0 .. nboxes-1 there)
            t1 = torch.tensor([sum(results[path]["ms_steps"]) / len(results[path]["ms_steps"])], dtype=torch.float64, device=dev)
            dist.broadcast(t1, src=0)
            ms = 1e3 * leg["t_max"] / steps_s
            strong[path] = {"value": gs["n_cells"] * steps_s / leg["t_max"], "unit": UNIT, "ms_per_step": ms, "steps": steps_s,
                            "cells_total": gs["n_cells"], "boxes_per_gpu": len(smine), "ms_per_step_n1_rank0": float(t1.item()),
                            "efficiency_vs_n1": float(t1.item()) / (world * ms), "max_nst": gs["max_nst"],
                            "kernel_ms_this_rank": leg["kernel_ms"], "drain_tail_ms_this_rank": leg["drain_ms"],
                            "ms_steps_this_rank": leg["ms_steps"]}
            pbs.free()
            del pbs

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        os.environ["OMP_PROC_BIND"] = "close"
        v, info = cpu_reference_run(args, args.path, gidx, [boxes[i] for i in mine], args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]}

    if rank == 0:
        comps = COMPS[args.path]
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": config_of(args, args.path, world, head["pb_ncell"], head["nb"], extra={
                    "l2": "inputs (%.1f GB per GPU) are larger than L2" % (8e-9 * sum(comps.values()) * head["pb_ncell"]),
                    "restore": "mutated components are reset from a pristine device copy between steps, outside the event pairs",
                    "numa": numa}),
                "clocks": head["clocks"], "e2e": head["e2e"], "gpu_launches": 2 * args.steps,   # per step: hc_copy_words_kernel (tile descriptors) + hc_sorted_kernel
                "roofline": head["roofline"], "cpu_baseline": cpu,
                "paths": {p: {k: r[k] for k in ("value", "unit", "ms_per_step", "steps", "roofline", "e2e", "workload", "clocks", "kernel_ms_this_rank", "drain_tail_ms_this_rank")} for p, r in results.items()},
                "strong": strong,
                "stats": head["stats"], "ms_steps": head["ms_steps"], "wall_s_timed_loop": head["wall"], "gen_s": head["t_gen"]}
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
