"""Batch enumeration and sharding across GPUs.

The reference farms the variants of a task template over a thread pool or one SLURM job each
(newcheb.py:641-675, input_batch/launch_runT.sh).  Here the variants are enumerated the same way,
grouped by plan shape, split over the ranks of one node (one process per GPU) by
longest-processing-time-first, run as batched plans, and the per-run results are gathered once at
the end -- the only collective of the whole path.
"""
from __future__ import annotations

import copy
import math
from typing import Dict, List, Sequence

import numpy as np

SUBST = "@SUBST:LIST"


# ----------------------------------------------------------------------------- template expansion
def _find_substitutions(node, found):
    """Document-order walk collecting every dict that carries a "@SUBST:LIST" key
    (the matches of '$..*[?(@["@SUBST:LIST"])]', json_substitutions.py:54)."""
    if isinstance(node, dict):
        for key in node:
            child = node[key]
            if isinstance(child, dict) and SUBST in child:
                found.append((node, key))
            else:
                _find_substitutions(child, found)
    elif isinstance(node, list):
        for i, child in enumerate(node):
            if isinstance(child, dict) and SUBST in child:
                found.append((node, i))
            else:
                _find_substitutions(child, found)


def expand_substitutions(template: dict) -> List[dict]:
    """Cartesian product of all substitution lists, the FIRST list varying fastest
    (json_substitutions.py:70-99).  A template without lists yields itself once."""
    root = copy.deepcopy(template)
    slots = []
    _find_substitutions(root, slots)
    lists = [holder[key][SUBST] for holder, key in slots]
    if any(len(l) == 0 for l in lists):
        return []
    out = []
    idx = [0] * len(lists)
    while True:
        for (holder, key), values, i in zip(slots, lists, idx):
            holder[key] = values[i]
        out.append(copy.deepcopy(root))
        k = 0
        while k < len(idx):
            idx[k] += 1
            if idx[k] < len(lists[k]):
                break
            idx[k] = 0
            k += 1
        if k == len(idx):
            break
    return out


# ----------------------------------------------------------------------------- per-run derived parameters
def apply_auto_rules(conf, run_index: int = 0, seed_base: int = 1000):
    """The *_auto rules of newcheb.py:905-936 on a loaded TaskRootConfiguration.  The reference draws
    w_list from an UNSEEDED generator (newcheb.py:905-909); for reproducible parity the draw is pinned
    to default_rng(seed_base + run_index)."""
    from .config import TaskRootConfiguration as T
    G = T.FitterConfiguration.InitGuess
    nw = int(2 * conf.fitter.pcos - 1)
    if not conf.fitter.w_list:
        rng = np.random.default_rng(seed_base + run_index)
        conf.fitter.w_list = [x for x in rng.uniform(conf.fitter.w_min, conf.fitter.w_max, nw)]
    else:
        assert len(conf.fitter.w_list) == nw
    cp = conf.fitter.propagation
    if conf.sigma_auto:
        if conf.init_guess == G.SQRSIN:
            cp.sigma = np.float64(2.0 * conf.T)
        elif conf.init_guess == G.MAXWELL:
            cp.sigma = np.float64(conf.T / 5.0)
        elif conf.init_guess == G.GAUSS:
            cp.sigma = np.float64(conf.T / 8.0)
    if conf.t0_auto:
        if conf.init_guess in (G.SQRSIN, G.MAXWELL):
            cp.t0 = np.float64(0.0)
        elif conf.init_guess == G.GAUSS:
            cp.t0 = np.float64(conf.T / 2.0)
    if conf.nt_auto:
        cp.nt = math.floor(conf.T / 2.0E-15)
    return conf


def estimate_cost(conf) -> float:
    """Relative cost of one run: time steps x basis vectors x levels x grid work x iteration budget."""
    cp = conf.fitter.propagation
    n = int(conf.np)
    grid = n * max(1.0, math.log2(n)) if n > 1 else 1.0
    it = conf.fitter.iter_max if conf.fitter.iter_max > 0 else 1000
    return float(cp.nt) * conf.nb * conf.nlevs * grid * (cp.nch - 1) * it


# ----------------------------------------------------------------------------- sharding
def lpt_shards(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of run indices to `world` ranks.  Deterministic: ties
    are broken by run index, so every rank computes the same partition without communicating."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += costs[i]
    for s in shards:
        s.sort()
    return shards


def plan_groups(shapes: Sequence[tuple], indices: Sequence[int]) -> Dict[tuple, List[int]]:
    """Runs that can share one plan (same hamil, np, nlevs, nb, nch, hf_hide)."""
    groups: Dict[tuple, List[int]] = {}
    for i in indices:
        groups.setdefault(shapes[i], []).append(i)
    return groups


def gather_results(local: Dict[int, np.ndarray], n_runs: int, width: int):
    """Final collective: every rank contributes rows (one float64 vector of `width` per run it owns);
    all ranks receive the full [n_runs, width] table.  Uses torch.distributed all_gather on padded
    tensors (NCCL over NVLink on GPUs, gloo in the CPU tests); a no-op without a process group."""
    table = np.full((n_runs, width), np.nan)
    for i, row in local.items():
        table[i, :len(row)] = row
    try:
        import torch
        import torch.distributed as dist
    except Exception:
        return table
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return table
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(np.nan_to_num(table, nan=0.0)).to(dev)
    mask = torch.from_numpy((~np.isnan(table)).astype(np.float64)).to(dev)
    parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    masks = [torch.empty_like(mask) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, mine)
    dist.all_gather(masks, mask)
    out = np.full((n_runs, width), np.nan)
    for p, m in zip(parts, masks):
        p, m = p.cpu().numpy(), m.cpu().numpy() > 0
        out[m] = p[m]
    return out
