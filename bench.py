#!/usr/bin/env python
"""bench.py -- HeatCool cell-updates/s on synthetic Lyman-alpha fields (BASELINE.json metric).

One "step" = one pass of the hot path over the rank's boxes: the Strang half-step
Nyx::integrate_state_vec (default) or the SDC step Nyx::integrate_state_struct (--path struct)
of a 512^3-per-GPU field (lognormal density, log-uniform 1e3..1e7 K, z = 3, TREECOOL_middle), through
the C-ABI of include/nyx_hc.h.  `value` = device-resident throughput (CUDA events on the launching
stream, max over ranks), `e2e` = the same call on pinned HOST FABs (H2D + kernel + D2H inside the timed
region), `roofline` = algorithmic FP64 flops / event time against the measured DFMA peak (plus HBM GB/s),
`cpu_baseline` = the reference's own OpenMP implementation (oracle/_ref, else the C port) on a bounded
sample of the same boxes.  `--impl reference` times only that CPU arm.
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
    ap.add_argument("--path", default="vec", choices=["vec", "struct"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workload
def workload_name(args):
    what = "integrate_state_vec Strang half-step dt/2" if args.path == "vec" else "integrate_state_struct SDC step dt"
    return f"Exec/LyA {args.n}^3 per GPU synthetic lognormal, z={args.z:g}, max_grid_size {args.box}, {what}"


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
def cpu_reference_run(args, boxes_idx, boxes, budget_s, repeat=1):
    """Times the reference's own CPU implementation (oracle/_ref OpenMP build; C port if it is absent) on as many
    boxes of the workload as fit `budget_s`.  Returns (cells_per_s, info dict)."""
    from nyx_b200 import synth
    from oracle import pyref
    a = 1.0 / (1.0 + args.z)
    dt = synth.step_dt(args.z)
    kind, cores, ref, port = "reference", 1, None, None
    try:
        ref = pyref.Reference("omp")
        # all the host threads the box offers (torchrun exports OMP_NUM_THREADS=1 to its children; NYX_REF_THREADS overrides)
        want = int(os.environ.get("NYX_REF_THREADS", "0")) or len(os.sched_getaffinity(0))
        ref.set("omp.num_threads", want)
        cores = ref.max_threads()
    except (FileNotFoundError, OSError):
        kind = "port"
        port = pyref.Port()

    def run_boxes(idx_list):
        fields = [make_box_fields(args, boxes_idx[i], *boxes[i]) for i in idx_list]
        cells = sum(f[0].shape[1] * f[0].shape[2] * f[0].shape[3] for f in fields)
        t0 = time.perf_counter()
        if args.path == "vec":
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

    # calibrate on one box, then size the sample to the budget
    c1, t1 = run_boxes([0])
    nb = int(max(1, min(len(boxes), budget_s / max(t1, 1e-3))))
    if nb > 1:
        cells, t = run_boxes(list(range(nb)))
    else:
        cells, t = c1, t1
    info = {"kind": kind, "cores": cores, "sample": f"{nb} box(es) of the workload ({cells} cells) in {t:.2f} s; first box alone {t1:.2f} s",
            "seconds": t, "cells": cells}
    return cells / t, info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from nyx_b200 import sharded
    boxes = sharded.box_list(args.n, args.box)
    idx = list(range(len(boxes)))
    per_step = max(2.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    vals, info = [], None
    for s in range(args.warmup + args.steps):
        v, info = cpu_reference_run(args, idx, boxes, per_step)
        if s >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals)) if vals else 0.0
    cells_per_step = info["cells"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * cells_per_step / value if value else None, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "note": "reference CPU/OpenMP implementation (oracle/_ref) on a bounded sample of the workload's boxes"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
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
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    hc = capi.NyxHC()
    hc.tables_upload(hc.tabulate_rates(TREECOOL, synth.mean_rhob()))
    a = 1.0 / (1.0 + args.z)
    dt = synth.step_dt(args.z)
    a_end = synth.a_after(args.z, dt)

    # ---- the rank's boxes: weak scaling = every rank owns a full n^3 domain's worth of boxes of a world-times larger field
    boxes = sharded.box_list(args.n, args.box)
    if args.scaling == "weak":
        mine = list(range(len(boxes)))
        gidx = [rank * len(boxes) + i for i in mine]
    else:
        mine = sharded.local_boxes(boxes, world, rank)
        gidx = list(mine)
    nb = len(mine)
    shapes = [tuple(h - l + 1 for l, h in zip(*boxes[i])) for i in mine]
    ncell_local = sum(s[0] * s[1] * s[2] for s in shapes)

    def alloc(ncomp, pinned=False):
        tot = sum(ncomp * s[0] * s[1] * s[2] for s in shapes)
        if pinned:
            return torch.empty(tot, dtype=torch.float64, pin_memory=True)
        return torch.empty(tot, dtype=torch.float64, device=dev)

    def views(buf, ncomp):
        out, off = [], 0
        for s in shapes:
            n = ncomp * s[0] * s[1] * s[2]
            out.append(buf[off:off + n].view(ncomp, s[2], s[1], s[0]))
            off += n
        return out

    struct = args.path == "struct"
    comps = {"state": 6, "diag": 2}
    if struct:
        comps.update({"s_new": 6, "hydro_src": 6, "reset_src": 1, "ir": 1})
    host = {k: alloc(c, pinned=True) for k, c in comps.items()}
    host_v = {k: views(host[k], c) for k, c in comps.items()}
    t_gen = time.perf_counter()
    for b, i in enumerate(mine):
        st, dg = make_box_fields(args, gidx[b], *boxes[i])
        host_v["state"][b].copy_(torch.from_numpy(st))
        host_v["diag"][b].copy_(torch.from_numpy(dg))
        if struct:
            host_v["s_new"][b].copy_(torch.from_numpy(st))
    if struct:
        host["hydro_src"].zero_(); host["reset_src"].zero_(); host["ir"].zero_()
    t_gen = time.perf_counter() - t_gen
    pristine_host = {k: host[k].clone() for k in (("state", "diag") if not struct else ("s_new", "diag", "ir"))}

    devb = {k: alloc(c) for k, c in comps.items()}
    for k in comps:
        devb[k].copy_(host[k], non_blocking=True)
    torch.cuda.synchronize()
    dev_v = {k: views(devb[k], c) for k, c in comps.items()}
    mutated = ("state", "diag") if not struct else ("s_new", "diag", "ir")
    pristine_dev = {k: devb[k].clone() for k in mutated}

    los = [boxes[i][0] for i in mine]
    tiles = [capi.make_box(*boxes[i]) for i in mine]
    dfab = {k: [capi.fab_of_torch(v, lo) for v, lo in zip(dev_v[k], los)] for k in comps}
    hfab = {k: [capi.make_fab(v.data_ptr(), lo, (s[0], s[1], s[2]), comps[k]) for v, lo, s in zip(host_v[k], los, shapes)] for k in comps}
    stream = torch.cuda.current_stream()

    def step_device():
        if struct:
            return hc.integrate_struct_batch(dfab["state"], dfab["diag"], dfab["s_new"], dfab["hydro_src"], dfab["reset_src"], dfab["ir"], tiles,
                                             a, a_end, dt, 0, stream=stream.cuda_stream)
        return hc.integrate_vec_batch(dfab["state"], dfab["diag"], tiles, a, 0.5 * dt, stream=stream.cuda_stream)

    def step_host():
        if struct:
            return hc.integrate_struct_host(hfab["state"], hfab["diag"], hfab["s_new"], hfab["hydro_src"], hfab["reset_src"], hfab["ir"], tiles,
                                            a, a_end, dt, 0)
        return hc.integrate_vec_host(hfab["state"], hfab["diag"], tiles, a, 0.5 * dt)

    def restore_device():
        for k in mutated:
            devb[k].copy_(pristine_dev[k])

    def restore_host():
        for k in mutated:
            host[k].copy_(pristine_host[k])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg
    for _ in range(args.warmup):
        restore_device()
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stats = None
    wall0 = time.perf_counter()
    for s in range(args.steps):
        restore_device()          # untimed: every step integrates the same input (the path updates its FABs in place)
        torch.cuda.synchronize()
        ev[s][0].record(stream)
        stats = step_device()     # one persistent kernel + the 112-byte statistics read-back
        ev[s][1].record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.result()
    ms_steps = [e0.elapsed_time(e1) for e0, e1 in ev]
    t_local = sum(ms_steps) * 1e-3
    t_max = t_local
    if world > 1:
        tt = torch.tensor([t_local], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_max = float(tt.item())
    gstats = sharded.allreduce_stats(stats, device=dev)   # the path's only collective: failure / iteration diagnostics
    ncell_global = gstats["n_cells"]
    value = ncell_global * args.steps / t_max

    # ---- roofline of the dominant kernel (sorted::hc_sorted_kernel): algorithmic flops / event time vs measured DFMA peak
    fp64_peak = hc.measure_fp64_peak()
    flops_local = sharded.algorithmic_flops(stats)
    t_kernel = t_local / args.steps
    achieved = flops_local / t_kernel
    bytes_cell = 104 if struct else 56
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_ach = bytes_cell * stats.n_cells / t_kernel / 1e9
    traffic = None
    try:   # DRAM bytes per cell of the dominant kernel from the committed ncu capture, scaled to this launch's cell count
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            traffic = json.load(f)["dram_bytes_per_cell"] * stats.n_cells
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"bound": "fp64", "achieved": achieved / 1e12, "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                "traffic": traffic, "traffic_note": "dram read+write bytes per launch: B/cell of the ncu --set full capture in profiles/r1_traffic.json, scaled by cells (algorithmic: %d B/cell)" % (104 if struct else 56), "peak_source": "DFMA peak measured in this run by hc_measure_fp64_peak (MEASURED_PEAKS.json has no FP64 entry)",
                "flops_per_cell": flops_local / stats.n_cells, "kernel": "sorted::hc_sorted_kernel<%s>" % ("PATH_STRUCT, 320" if struct else "PATH_VEC, 384"),
                "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak, "bytes_per_cell": bytes_cell,
                        "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"}}

    # ---- end-to-end leg: host FABs in pinned memory through the *_host entry points (H2D + kernel + D2H timed)
    e2e = None
    if not args.no_e2e:
        h2d_comps = {"state": 3, "diag": 2} if not struct else {"state": 3, "diag": 2, "s_new": 3, "hydro_src": 2, "reset_src": 1}
        d2h_comps = {"state": 2, "diag": 2} if not struct else {"s_new": 2, "ir": 1, "diag": 2}
        h2d = 8 * ncell_local * sum(h2d_comps.values())
        d2h = 8 * ncell_local * sum(d2h_comps.values())
        n_e2e_warm, n_e2e = 1, max(1, min(args.steps, 3))
        for _ in range(n_e2e_warm):
            restore_host(); step_host()
        barrier()
        t_e2e = 0.0
        for _ in range(n_e2e):
            restore_host()
            barrier()
            t0 = time.perf_counter()
            st_h = step_host()            # returns after the D2H copies have completed (stream-synchronised inside)
            torch.cuda.synchronize()
            t_e2e += time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_e2e = float(tt.item())
        e2e = {"value": ncell_global * n_e2e / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": n_e2e,
               "ms_per_step": 1e3 * t_e2e / n_e2e, "api": "hc_integrate_%s_host on pinned host FABs (H2D / kernel / D2H pipelined over 8 groups of boxes)" % args.path, "n_failed": st_h.n_failed}

    # ---- CPU baseline (rank 0, N = 1 only): the reference's OpenMP implementation on a bounded sample of the same boxes
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, info = cpu_reference_run(args, gidx, [boxes[i] for i in mine], args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args), "cells_per_gpu": ncell_local, "boxes_per_gpu": nb, "path": args.path, "z": args.z,
                           "rtol": 1e-4, "atol_factor": 1e-4, "l2": "inputs (%.1f GB per GPU) are larger than L2" % (8e-9 * sum(comps.values()) * ncell_local),
                           "restore": "mutated components are reset from a pristine device copy between steps, outside the event pairs",
                           "parallelism": "boxes sharded over %d GPU(s), no data-path collective, one scalar all-reduce of diagnostics" % world},
                "clocks": clocks, "e2e": e2e, "gpu_launches": 2 * args.steps,   # per step: hc_copy_words_kernel (tile descriptors) + hc_sorted_kernel
                "roofline": roofline, "cpu_baseline": cpu,
                "stats": gstats, "ms_steps": ms_steps, "wall_s_timed_loop": wall, "gen_s": t_gen}
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
