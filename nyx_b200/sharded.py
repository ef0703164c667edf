"""Box-sharded driver of the HeatCool path over N GPUs of one node.

The reference distributes boxes over MPI ranks with an AMReX DistributionMapping and does no
communication on this path except `ParallelDescriptor::ReduceLongMax(new_max_sundials_steps)` in the
callers (Source/HeatCool/strang_reactions.cpp:32,87, sdc_reactions.cpp:31).  Here: one process per GPU,
boxes dealt round-robin after sorting by cell count (what AMReX's knapsack does for equal boxes), NO
data-path collective, and one scalar all-reduce (SUM of the counters, MAX of max_nst) of HcStats through
torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np

from . import capi


def box_list(domain_n, max_grid_size):
    """BoxArray(domain).maxSize(max_grid_size) for a cubic domain: list of (lo, hi), x fastest."""
    n = int(domain_n)
    m = int(max_grid_size)
    edges = list(range(0, n, m))
    out = []
    for k0 in edges:
        for j0 in edges:
            for i0 in edges:
                lo = (i0, j0, k0)
                hi = (min(i0 + m, n) - 1, min(j0 + m, n) - 1, min(k0 + m, n) - 1)
                out.append((lo, hi))
    return out


def box_cells(b):
    lo, hi = b
    return (hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1)


def distribution_map(boxes, world_size):
    """owner rank of each box: sort by cell count (descending, stable), deal round-robin."""
    order = sorted(range(len(boxes)), key=lambda i: -box_cells(boxes[i]))
    owner = [0] * len(boxes)
    for pos, i in enumerate(order):
        owner[i] = pos % world_size
    return owner


def local_boxes(boxes, world_size, rank):
    owner = distribution_map(boxes, world_size)
    return [i for i, o in enumerate(owner) if o == rank]


SUM_FIELDS = tuple(f for f in capi.STATS_FIELDS if f != "max_nst")


def stats_to_array(st):
    d = st if isinstance(st, dict) else st.as_dict()
    return np.array([d[f] for f in capi.STATS_FIELDS], dtype=np.int64)


def allreduce_stats(st, device=None):
    """Global HcStats as a dict: counters summed, max_nst maximised over ranks (one tiny collective each)."""
    import torch
    import torch.distributed as dist
    arr = stats_to_array(st)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(zip(capi.STATS_FIELDS, (int(x) for x in arr)))
    t = torch.from_numpy(arr.copy())
    if device is not None:
        t = t.to(device)
    imax = capi.STATS_FIELDS.index("max_nst")
    mx = t[imax:imax + 1].clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    t[imax] = mx[0]
    return dict(zip(capi.STATS_FIELDS, (int(x) for x in t.cpu().tolist())))


def algorithmic_flops(stats):
    """SURVEY.md section 8d hand count: 186 per iterate_ne Newton iteration (146 flops + 2 transcendentals at weight 20),
    234 per RHS evaluation (174 + 3 x 20), 60 per step attempt, 99 per finalize EOS solve."""
    d = stats if isinstance(stats, dict) else stats.as_dict()
    return 186.0 * d["sum_ne_iters"] + 234.0 * (d["sum_nfe"] + d["sum_nfe_ls"]) + 60.0 * d["sum_attempts"] + 99.0 * d["sum_eos"]


def allreduce_min(value, device=None):
    """min over ranks of one double: the global S_new.min(Density_comp) of Nyx::enforce_minimum_density
    (Source/TimeStep/Nyx_enforce_minimum_density.cpp:22) -- the one real exchange step of the rank-2 row"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t.item())


def update_state_with_sources_sharded(update_local, enforce_local, small_dens, device=None):
    """Nyx::update_state_with_sources over the boxes of this rank, with the reference's GLOBAL density-floor decision.
    update_local() runs hc_update_state_with_sources_batch on the local boxes and returns their minimum new density (it has already
    applied the floor if that minimum is below small_dens); enforce_local() runs hc_enforce_minimum_density_batch on them.
    Returns (local_min, global_min, enforced_here_afterwards)."""
    local_min = update_local()
    global_min = allreduce_min(local_min, device)
    late = (global_min < small_dens) and not (local_min < small_dens)
    if late:
        enforce_local()
    return local_min, global_min, late
