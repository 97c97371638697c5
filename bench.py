#!/usr/bin/env python
"""Benchmark of the propagation hot path (BASELINE.json metric: wavefunction grid-pt*timesteps/s).

Workload (config.workload = "krotov_batch_8192x2048"): BASELINE config 5 -- per GPU ``--runs`` independent
two-level Krotov problems on a 2048-point FFT grid, nch = 64 interpolation points, ``--nt`` time steps.
One bench "step" is one Krotov iteration of the whole batch: forward sweep with the in-kernel field
update from the stored backward trajectory, goal overlaps, chi(T) construction, backward sweep
(fitter.py:441-489), all device-resident.  grid-pt*timesteps = runs * np * (2*nt) per step.

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference --steps K --warmup W    # the CPU path (numpy oracle port) on host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions of value / e2e /
roofline / cpu_baseline.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "wavefunction grid-pt*timesteps/s"
UNIT = "grid-pt*timesteps/s"
NP_, NCH, NLEVS = 2048, 64, 2


def flops_per_step(np_=NP_, nlevs=NLEVS, nch=NCH, overlap=True):
    """F_step = (nch-1)*nlevs*np*(10*log2 np + 18) + 18*nlevs*np (+ 8*np Krotov overlap)."""
    return (nch - 1) * nlevs * np_ * (10 * math.log2(np_) + 18) + 18 * nlevs * np_ + (8 * np_ if overlap else 0)


# ------------------------------------------------------------------------------------------------
# CPU arm: the numpy port of the reference path (oracle/), one process per core
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    nsteps, seed = args
    from oracle import qc_oracle as qo
    from qcontrol_b200.synthetic import krotov_batch  # plain numpy set-up code; loads no CUDA library symbols it calls
    run = krotov_batch(1, NP_, nsteps, NCH, seed=seed)[0]
    pb = qo.Problem(task_type="trans_wo_control", hamil="ntriv", ntriv=1, np_=NP_, nlevs=2, nb=1, L=run.L, T=run.T,
                    dx=run.dx, x=run.x, v=run.v, vmin=run.vmin, akx2=run.akx2, psi0=run.psi0, psif=run.psif, m=run.m,
                    nch=NCH, nt=run.nt, E0=run.E0, t0=run.t0, sigma=run.sigma, nu_L=run.nu_L, nu=run.nu,
                    envelope="gauss", carrier="exp", hf_hide=True, t_step=run.t_step, t_list=run.t_list)
    st = qo.Stepper(pb, lambda s: pb.laser_field(pb.E0, s.t, pb.t0, pb.sigma), instrument=False)
    st._coeffs = lambda t_sc: qo.points(pb.nch, t_sc)     # the reference rebuilds the table every step (phys_base.py:155)
    st.start(pb.psi0[0], pb.psif[0], qo.FORWARD)
    t0 = time.perf_counter()
    for _ in range(nsteps):
        st.step(np.float64(0.0))
    return time.perf_counter() - t0


def cpu_rate(nsteps_per_proc, procs):
    """grid-pt*timesteps/s of the oracle port over `procs` processes, each propagating one run."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_worker, [(nsteps_per_proc, 100 + i) for i in range(procs)])
    wall = time.perf_counter() - t0
    return procs * NP_ * nsteps_per_proc / wall, wall


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nsteps = a.cpu_steps
    for _ in range(a.warmup if a.warmup < 2 else 1):
        cpu_rate(max(2, nsteps // 8), cores)
    vals = []
    t0 = time.perf_counter()
    for _ in range(a.steps):
        r, _w = cpu_rate(nsteps, cores)
        vals.append(r)
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    sample = "%d processes x 1 run x %d time steps (np=%d, nch=%d, 2 levels, hf_hide) per bench step; pool start-up included" % (
        cores, nsteps, NP_, NCH)
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * wall / max(a.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "krotov_batch_8192x2048", "np": NP_, "nch": NCH, "nlevs": NLEVS,
                       "impl_note": "numpy port of the reference path (oracle/qc_oracle.py); the reference is pure Python and /root/reference does not travel"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(a):
    import torch
    import torch.distributed as dist
    from qcontrol_b200 import control, engine
    from qcontrol_b200.synthetic import krotov_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL writes its version banner to stdout when the communicator is created: send it to stderr, so that
        # stdout carries the one JSON line only
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    # One explicit stream for everything: the library launches on it and the timing events are recorded on it.
    # (torch's default stream has handle 0, which qc_create reads as "create your own stream": events recorded
    # on torch's default stream would then not wait for the library's kernels.)
    tstream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(tstream)
    ctx = engine.Context(local, tstream.cuda_stream)
    B, nt = a.runs, a.nt

    pbs = krotov_batch(B, NP_, nt, NCH, seed=12345 + rank)
    gb = control.GridBatch(pbs, ctx=ctx)
    pl = gb.plan
    sqrt_hf_T = np.array([np.complex128(np.sqrt(p.laser_field_hf(np.float64(1.0), p.T, p.pcos, p.w_list))) for p in pbs])

    # iteration 0 (prescribed forward field) populates both trajectory buffers with real data
    pl.snapshot_load(0)
    pl.record_state(0, 0, 0)
    pl.set_E_base(gb._pad([gb._prescribed_field(p, control.FORWARD) for p in pbs]))
    pl.sweep(0, engine.FIELD_PRESCRIBED, 1, nt, True, -1, 0)
    pl.goal_overlap(1, fetch=False)
    pl.krotov_terminal(1, sqrt_hf_T, 1)
    pl.sweep(1, engine.FIELD_KROTOV_BWD, 1, nt, True, 0, 1)

    def krotov_iteration():
        pl.snapshot_load(0)
        pl.record_state(0, 0, 0)
        pl.sweep(0, engine.FIELD_KROTOV_FWD, 0, nt + 1, True, 1, 0)
        gc = pl.goal_overlap(1)                      # D2H of the per-run fidelities
        pl.krotov_terminal(1, sqrt_hf_T, 1)
        pl.sweep(1, engine.FIELD_KROTOV_BWD, 1, nt, True, 0, 1)
        return gc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        krotov_iteration()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(tstream)
    ctx.timer_start()
    for _ in range(a.steps):
        gc = krotov_iteration()
    ev1.record(tstream)
    lib_ms = ctx.timer_stop()                 # the library's own cudaEvent pair on the same stream (cross-check)
    barrier()
    launches = ctx.launches - launches0
    ms = ev0.elapsed_time(ev1)
    assert abs(ms - lib_ms) <= 0.02 * ms + 0.5, ("timing events disagree", ms, lib_ms)
    sampler.stop_flag = True
    sampler.join()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # the only collective of the path: gather the per-run fidelities of every rank (not timed per step)
        g = torch.from_numpy(np.ascontiguousarray(gc.view(np.float64))).cuda()
        allg = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(allg, g)
    units_per_step = world * B * NP_ * 2 * nt
    value = units_per_step * a.steps / (ms * 1e-3)
    iters_per_s = world * B * a.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with host buffers: every rank at the same time, max over ranks ----------
    hin = torch.empty(pl.state_shape, dtype=torch.complex128).pin_memory()
    hout = torch.empty(pl.state_shape, dtype=torch.complex128).pin_memory()
    hin.numpy()[...] = gb.psi0
    pl.set_E_base(gb._pad([gb._prescribed_field(p, control.FORWARD) for p in pbs]))
    e2e_t = []
    for i in range(3):
        barrier()
        t0 = time.perf_counter()
        pl.propagate_host(hin.numpy(), hout.numpy(), slot=0, l_first=1, nsteps=nt, renorm=True)
        e2e_t.append(time.perf_counter() - t0)
    e2e_s = float(np.min(e2e_t[1:]))
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": world * B * NP_ * nt / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(world * hin.numel() * 16),
           "d2h_bytes_per_step": int(world * hout.numel() * 16),
           "call": "qc_propagate_host on every GPU at once: H2D psi (pinned) + %d prescribed-field steps + D2H psi, "
                   "copies pipelined against the propagation in chunks of 592 runs; slowest rank" % nt}

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel: forward Krotov sweep, timed alone with CUDA events -----
        durs = []
        for _ in range(3):
            pl.snapshot_load(0)
            ctx.synchronize()
            ctx.timer_start()
            pl.sweep(0, engine.FIELD_KROTOV_FWD, 1, nt, True, 1, 0)
            durs.append(ctx.timer_stop())
        kdur = float(np.mean(durs)) * 1e-3
        flops_launch = flops_per_step() * B * nt
        achieved = flops_launch / kdur / 1e12
        peak_nominal = 148 * 64 * 2 * 1.965e9 / 1e12
        peak_meas = ctx.fp64_peak_tflops()
        traffic = traffic_src = None
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic_grid_sweep.json")
        if os.path.exists(tpath):         # DRAM bytes of this kernel from the committed ncu capture, scaled per (run, step)
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes_per_run_step"] * B * nt
            traffic_src = tj["source"]
        roof = {"bound": "fp64", "achieved": achieved, "peak": peak_meas, "unit": "TFLOP/s", "frac": achieved / peak_meas,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": "measured here: qc_measure_fp64_peak (DFMA chains, burst); MEASURED_PEAKS.json has no FP64 entry",
                "peak_nominal": peak_nominal, "kernel": "grid_sweep_kernel<11> (Krotov forward sweep)",
                "kernel_ms_per_launch": kdur * 1e3, "flops_per_launch": flops_launch,
                "hbm_algorithmic_bytes_per_launch": (32 * NP_ + 16) * B * nt,
                "hbm_gbs_achieved": (32 * NP_ + 16) * B * nt / kdur / 1e9}
        # ---- CPU baseline beside it ---------------------------------------------------------------------
        cpu = None
        if not a.no_cpu:
            cores = os.cpu_count() or 1
            v, wall = cpu_rate(a.cpu_steps, cores)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d processes x 1 run x %d time steps of the same workload (np=%d, nch=%d); %.1f s wall" % (
                       cores, a.cpu_steps, NP_, NCH, wall)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "krotov_batch_8192x2048", "runs_per_gpu": B, "np": NP_, "nlevs": NLEVS, "nch": NCH,
                           "nt": nt, "step": "one Krotov iteration (forward + backward sweep) of every run",
                           "l2": "working set (trajectories + state, %.1f GB) exceeds L2" % (
                               (2 * (nt + 1) * B * NP_ * 16 + 3 * B * 2 * NP_ * 16) / 1e9)},
                "krotov_iterations_per_s": iters_per_s, "gpu_launches": int(launches), "clocks": sampler.summary(),
                "roofline": roof, "e2e": e2e, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    gb.close()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--runs", type=int, default=8288,
                    help="independent runs per GPU: the 8192 of BASELINE config 5 rounded up to 56 x 148 SMs")
    ap.add_argument("--nt", type=int, default=16, help="time steps per sweep")
    ap.add_argument("--cpu-steps", type=int, default=96, help="time steps each host process propagates for the CPU figure")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_gpu_arm(a)


if __name__ == "__main__":
    main()
