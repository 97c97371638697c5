"""Batch entry point with the reference's command line (newcheb.py:612-979):

    python -m qcontrol_b200.newcheb --json_task TASK.json --json_rep REPORT.json [--seed_base N]
    torchrun --nproc-per-node 8 -m qcontrol_b200.newcheb --json_task ... --json_rep ...

The task template is expanded over its "@SUBST:LIST" entries exactly like the reference does; instead of
one thread-pool job (or SLURM job) per variant, variants that share a plan shape become ONE batched plan
per GPU, the variants are split over the ranks by longest-processing-time-first, and each rank writes the
output of its own runs under the report's ``out_path`` (with the reference's ``{input:$.json.path}``
templating) through ``qcontrol_b200.reporter``: ``table_inp_-1.txt`` (newcheb.py:526-573), tab_iter.csv /
tab_iter_E.csv, and -- for plotting_flag "tables" / "all" -- the per-step tables of every sweep
(reporter.py:73-164).  The HTML plots are presentation only and are not reproduced.
"""
from __future__ import annotations

import copy
import getopt
import json
import os
import re
import sys
from types import SimpleNamespace

import numpy

from . import batch as qbatch
from . import config as qconfig
from . import control, engine as _eng, math_base, phys_base, reporter, task_manager
from .config import ReportPlotRootConfiguration, ReportRootConfiguration, ReportTableRootConfiguration, TaskRootConfiguration

_TT = TaskRootConfiguration.TaskType
_NAMES = {_TT.SINGLE_POT: "single_pot", _TT.FILTERING: "filtering", _TT.TRANS_WO_CONTROL: "trans_wo_control",
          _TT.INTUITIVE_CONTROL: "intuitive_control", _TT.LOCAL_CONTROL_POPULATION: "local_control_population",
          _TT.LOCAL_CONTROL_PROJECTION: "local_control_projection", _TT.OPTIMAL_CONTROL_KROTOV: "optimal_control_krotov",
          _TT.OPTIMAL_CONTROL_UNIT_TRANSFORM: "optimal_control_unit_transform"}
_HAMIL = {_eng.HAMIL_NONTRIV: "ntriv", _eng.HAMIL_TWO_LEVELS: "two_levels", _eng.HAMIL_BH_X: "bh_x",
          _eng.HAMIL_BH_Z: "bh_z", _eng.HAMIL_QUBIT: "qubit"}


def usage():
    print("Usage: python -m qcontrol_b200.newcheb --json_task <task.json> --json_rep <report.json> [--json_create] "
          "[--seed_base N]")


def resolve_input_templates(data_task, node):
    """Replace every "{input:$.a.b}" inside the report configuration by the value found at that path of
    the task variant (newcheb.py:576-610; only plain dotted paths occur in the shipped files)."""
    if isinstance(node, dict):
        return {k: resolve_input_templates(data_task, v) for k, v in node.items()}
    if isinstance(node, list):
        return [resolve_input_templates(data_task, v) for v in node]
    if isinstance(node, str):
        def sub(m):
            cur = data_task
            for part in m.group(1).strip().lstrip("$").strip(".").split("."):
                if part:
                    cur = cur[part]
            return str(cur)
        return re.sub(r"\{input:([^\}]*)\}", sub, node)
    return node


def validate(conf):
    """The subset of newcheb.py:691-902 that guards the supported task types."""
    cp = conf.fitter.propagation
    if not (conf.np > 0 and (conf.np & (conf.np - 1)) == 0):
        raise ValueError("The number of collocation points 'np' must be a positive integer power of 2")
    if not (cp.nch > 0 and (cp.nch & (cp.nch - 1)) == 0):
        raise ValueError("The number of interpolation points 'nch' must be a positive integer power of 2")
    if not conf.T > 0.0 or not conf.L > 0.0:
        raise ValueError("'T' and 'L' must be positive")
    if conf.task_type == _TT.OPTIMAL_CONTROL_UNIT_TRANSFORM:
        if conf.np > 1:
            raise ValueError("For the 'task_type' = 'optimal_control_unit_transform' the number of collocation points, "
                             "'np', has to be equal to 1")
        if conf.init_guess == TaskRootConfiguration.FitterConfiguration.InitGuess.ZERO:
            raise ValueError("For the 'task_type' = 'optimal_control_unit_transform' the initial guess type for the "
                             "laser field envelope, 'init_guess', mustn't be 'zero'")
        if conf.L != 1.0:
            raise ValueError("For the 'task_type' = 'optimal_control_unit_transform' the spatial range of the problem, "
                             "'L', has to be equal to 1.0")
    if conf.task_type not in _NAMES:
        raise NotImplementedError("task type %s is not ported to the GPU path yet" % conf.task_type)


def problem_from(conf, tm):
    """Dense run description for the batched control loops from a task manager's products."""
    cf, cp = conf.fitter, conf.fitter.propagation
    h = tm.hamil2D
    F = TaskRootConfiguration.FitterConfiguration
    return SimpleNamespace(
        task_type=_NAMES[conf.task_type], hamil=_HAMIL[h.kind], ntriv=tm.ntriv, np_=conf.np, nlevs=conf.nlevs,
        nb=conf.nb, L=conf.L, T=conf.T, dx=tm.dx, x=tm.x, v=h.potentials(), vmin=h.minima(), akx2=tm.akx2,
        psi0=tm.psi0.as_array(), psif=tm.psif.as_array(), U=conf.U, W=conf.W, delta=conf.delta, m=cp.m, nch=cp.nch,
        nt=cp.nt, E0=cp.E0, t0=cp.t0, sigma=cp.sigma, nu_L=cp.nu_L, pcos=cf.pcos, w_list=cf.w_list,
        hf_hide=bool(cf.hf_hide), h_lambda=cf.h_lambda, h_lambda_mode=cf.h_lambda_mode.name.lower(), epsilon=cf.epsilon,
        iter_max=cf.iter_max, iter_mid_1=cf.iter_mid_1, iter_mid_2=cf.iter_mid_2, q=cf.q,
        impulses_number=cf.impulses_number, delay=cf.delay, k_E=cf.k_E, lamb=cf.lamb, pow_=cf.pow,
        F_type={F.FType.SM: "sm", F.FType.RE: "re", F.FType.SS: "ss"}[cf.F_type], F_goal=tm.F_goal, t_step=tm.t_step,
        t_list=tm.t_list, laser_field=tm.laser_field, laser_field_hf=tm.laser_field_hf,
        envelope=conf.init_guess.name.lower(), carrier=conf.init_guess_hf.name.lower(), nu=tm.nu)


def print_input(out_path, conf_task, file_name):
    """The parameter sheet post_processing.py reads (newcheb.py:526-573): same keys, tabs and number formats."""
    c, f, p = conf_task, conf_task.fitter, conf_task.fitter.propagation
    rows = [("task_type:\t\t", c.task_type), ("iter_max:\t\t", f.iter_max), ("iter_mid_1:\t\t", f.iter_mid_1),
            ("iter_mid_2:\t\t", f.iter_mid_2), ("q:\t\t\t", f.q), ("epsilon:\t\t", "%.1E" % f.epsilon),
            ("nb:\t\t\t", c.nb), ("nlevs:\t\t\t", c.nlevs), ("wf_type:\t\t", c.wf_type),
            ("impulses_number:\t", f.impulses_number), ("Em:\t\t\t", f.Em), ("E0:\t\t\t", p.E0), ("t0:\t\t\t", p.t0),
            ("t0_auto:\t\t", c.t0_auto), ("sigma:\t\t\t", "%.6E" % p.sigma), ("sigma_auto:\t\t", c.sigma_auto),
            ("nu_L:\t\t\t", "%.6E" % p.nu_L), ("nu_L_auto:\t\t", c.nu_L_auto), ("h_lambda:\t\t", f.h_lambda),
            ("h_lambda_mode:\t\t", f.h_lambda_mode), ("init_guess:\t\t", c.init_guess),
            ("init_guess_hf:\t\t", c.init_guess_hf), ("F_type:\t\t\t", f.F_type), ("pcos:\t\t\t", f.pcos),
            ("hf_hide:\t\t", f.hf_hide), ("w_list:\t\t\t", f.w_list), ("w_min:\t\t\t", f.w_min), ("w_max:\t\t\t", f.w_max),
            ("hamil_type:\t\t", c.hamil_type), ("U:\t\t\t", c.U), ("W:\t\t\t", c.W), ("delta:\t\t\t", c.delta),
            ("lf_aug_type:\t\t", c.lf_aug_type), ("pot_type:\t\t", c.pot_type), ("Du:\t\t\t", c.Du), ("np:\t\t\t", c.np),
            ("L:\t\t\t", c.L), ("nch:\t\t\t", p.nch), ("nt:\t\t\t", p.nt), ("nt_auto:\t\t", c.nt_auto),
            ("T:\t\t\t", "%.6E" % c.T)]
    os.makedirs(out_path, exist_ok=True)
    with open(os.path.join(out_path, file_name), "w") as finp:
        for label, val in rows:
            finp.write("%s%s\n" % (label, val))


def write_tables(out_path, rows, t_list, nt, imin=-1):
    """tab_iter.csv / tab_iter_E.csv in the reference's format (reporter.py:609-632)."""
    os.makedirs(out_path, exist_ok=True)
    with open(os.path.join(out_path, "tab_iter.csv"), "w") as f, open(os.path.join(out_path, "tab_iter_E.csv"), "w") as g:
        for r in rows:
            if r.it < imin:
                continue
            f.write("{:2d} {:.6f} {:.6f} {:.4f} {:.6f}\n".format(int(r.it), r.goal_close, r.Fsm, r.E_int, r.J))
            if r.E_tlist is None:          # plotting_flag "tables_iter": no field table (reporter.py:627-630)
                continue
            for i in range(nt + 1):
                g.write("{:2d} {:.6f} {:.6e}\n".format(r.it, t_list[i] * 1e+15, abs(r.E_tlist[i])))


def memory_chunks(confs, idxs, free_bytes, fraction=0.8):
    """Split the runs ``idxs`` of one plan shape into sub-batches whose plans fit the free device memory.

    The reference keeps psi_tlist / chi_tlist in host RAM, one run per job (fitter.py:322-329); here the two
    trajectory buffers of a Krotov / UT plan live in HBM: 2 * (nt + 1) * np * 16 B per run on the grid (7.5 GB for
    the shipped Krotov input), so one plan per shape group does not fit arbitrary batches.  Runs are taken longest
    first, so that a sub-batch holds runs of similar length (the plan is sized by its longest run), and a sub-batch
    is closed when qc_plan_footprint of one more run would exceed ``fraction`` of the free bytes.  Inside a sub-batch
    the runs stay ordered by length: neighbouring runs share warps / SMs, so the warps of the short runs of a ragged
    batch retire together (kernel B) instead of idling beside a long run."""
    if not idxs:
        return []
    c0 = confs[idxs[0]]
    grid = c0.hamil_type == TaskRootConfiguration.HamilType.NTRIV
    hamil = _eng.HAMIL_NONTRIV if grid else _eng.HAMIL_BH_X
    traj = _NAMES.get(c0.task_type) in control.TRAJ_TASKS
    cap = fraction * free_bytes

    def fits(n, nt_max):
        return _eng.plan_footprint(hamil, c0.np, c0.nlevs, c0.nb, c0.fitter.propagation.nch, n, nt_max,
                                   bool(c0.fitter.hf_hide) and grid, traj) <= cap

    order = sorted(idxs, key=lambda i: (-int(confs[i].fitter.propagation.nt), i))
    chunks, cur = [], []
    for i in order:
        nt_max = int(confs[(cur or [i])[0]].fitter.propagation.nt)
        if cur and not fits(len(cur) + 1, nt_max):
            chunks.append(cur)
            cur = []
            nt_max = int(confs[i].fitter.propagation.nt)
        if not cur and not fits(1, nt_max):
            raise MemoryError("run #%d alone (nt = %d, np = %d) needs more than %.0f%% of the free device memory (%.1f GB)"
                              % (i, nt_max, c0.np, 100 * fraction, free_bytes / 1e9))
        cur.append(i)
    if cur:
        chunks.append(cur)
    return chunks


def pipeline_pieces(chunk, np_, sm_count, grid):
    """Cut a memory-sized sub-batch of a GRID shape into pieces of whole waves so that the host set-up of the next piece
    (task_manager.create, step-scalar and Newton tables, plan allocation and upload) runs on a second thread and a
    second stream while the GPU propagates the current one.  A piece is two waves of CTAs of kernel A (one run per
    CTA at np = 2048, two at 1024, four below, a two-CTA cluster per run at 4096); N-level batches are not cut -- kernel
    B wants as many runs per launch as it can get (kernel B)."""
    if not grid:
        return [chunk]
    runs_per_wave = sm_count * (4 if np_ <= 512 else 2 if np_ == 1024 else 1) // (2 if np_ >= 4096 else 1)
    piece = 2 * max(runs_per_wave, 1)
    if os.environ.get("QC_PIPELINE_PIECE"):         # tests: any piece size; 0 = never cut
        piece = int(os.environ["QC_PIPELINE_PIECE"]) or len(chunk)
    if len(chunk) <= piece:
        return [chunk]
    return [chunk[k:k + piece] for k in range(0, len(chunk), piece)]


def run_variants(variants, data_rep, seed_base=1000, device=None, verbose=True, timings=None):
    """Expand -> configure -> shard -> batched GPU runs -> per-run tables.  Returns {run index: rows}.
    ``timings`` (a dict, optional) receives the wall-clock split of this rank: configuration, per-run task set-up
    (task_manager.create + report set-up), plan set-up (tables of step scalars, Newton tables, uploads), the control
    loops themselves, and reporting.  The set-up of later pieces overlaps the control loops of earlier ones (one
    producer thread, QC_PIPELINE_RUNNERS = 1 runner thread, a context = stream per piece in flight): task / plan /
    report times are sums over the pieces, ``setup_wait_s`` is what passed before the first sweep could start and
    ``run_s`` the span from there to the end of the last control loop."""
    import concurrent.futures
    import threading
    import time
    tm_ = {"config_s": 0.0, "reserve_s": 0.0, "task_setup_s": 0.0, "plan_setup_s": 0.0, "setup_wait_s": 0.0, "run_s": 0.0,
           "report_s": 0.0, "teardown_s": 0.0, "plans": 0}
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0")) if device is None else device
    t_mark = time.perf_counter()
    confs, costs, shapes = [], [], []
    for i, data_task in enumerate(variants):
        with qconfig.quiet():
            conf = TaskRootConfiguration()
            conf.load(data_task)
        validate(conf)
        qbatch.apply_auto_rules(conf, run_index=i, seed_base=seed_base)
        confs.append(conf)
        costs.append(qbatch.estimate_cost(conf))
        shapes.append((conf.hamil_type, conf.lf_aug_type, conf.np, conf.nlevs, conf.nb, conf.fitter.propagation.nch,
                       bool(conf.fitter.hf_hide), conf.task_type, float(conf.L), float(conf.fitter.propagation.m)))
    mine = qbatch.lpt_shards(costs, world)[rank]
    results = {}
    # Pipeline: ONE producer thread sets pieces up (tasks, reporters, plan upload), two runner threads drive the control
    # loops of consecutive pieces, each piece on a context (= stream) of its own; at most RUNNERS + 1 plans are alive.
    # The runners are GATED: piece k + 1 starts when piece k has launched its last sweep (GridBatch.on_last_sweep; for
    # control loops without that hook: when piece k is finished), so its first sweep queues behind the last one of
    # piece k and the host work that ends a piece (final overlaps, tables, report files) costs the GPU nothing.  Two
    # pieces in flight without the gate were measured 3 % SLOWER than one (1024 runs, np = 2048, nt = 1024: 5.33 s
    # against 5.17 s of control loops): the CTAs of two sweeps interleave on the SMs and each sweep takes longer.
    # QC_PIPELINE_RUNNERS=1: strictly one piece after the other.
    runners = max(1, min(2, int(os.environ.get("QC_PIPELINE_RUNNERS", "2"))))
    ctxs = [_eng.Context(local) for _ in range(runners + 2)]
    OT = ReportRootConfiguration.ReportFitterConfiguration.OutputType
    tm_["config_s"] = time.perf_counter() - t_mark
    work = []
    sm_count = ctxs[0].sm_count
    t_mark = time.perf_counter()
    free_bytes = ctxs[0].mem_info()[0]
    for shape, group in qbatch.plan_groups(shapes, mine).items():
        try:
            # whole waves first (longest runs first, like memory_chunks orders them), then the memory budget per piece:
            # runners + 1 plans are alive at a time, so it is the PIECE that has to fit its share of the free memory
            c0 = confs[group[0]]
            order = sorted(group, key=lambda i: (-int(confs[i].fitter.propagation.nt), i))
            for piece in pipeline_pieces(order, int(c0.np), sm_count, c0.hamil_type == TaskRootConfiguration.HamilType.NTRIV):
                work += memory_chunks(confs, piece, free_bytes, fraction=0.8 / (runners + 1))
        except MemoryError as e:
            print("An exception interrupted the jobs %s: %s" % (group, e), file=sys.stderr)
    if len(work) <= 1:
        runners = 1
    alive = threading.Semaphore(runners + 1)
    lock = threading.Lock()
    gates = [threading.Event() for _ in work]           # gates[k]: piece k has launched its last sweep (or is done)
    marks = {"first_ready": None, "first_run": None, "last_run": None}

    def prepare(k):
        """Host set-up of piece k (producer thread): per-run tasks and reporters, then the plan on this piece's context."""
        alive.acquire()
        try:
            idxs, ctx = work[k], ctxs[k % len(ctxs)]
            t0 = time.perf_counter()
            pbs, reps = [], []
            for i in idxs:
                with qconfig.quiet():
                    tm = task_manager.create(confs[i])
                    rep_json = resolve_input_templates(variants[i], copy.deepcopy(data_rep))
                    ct, cp = ReportTableRootConfiguration(), ReportPlotRootConfiguration()
                    ct.load(rep_json)
                    cp.load(rep_json)
                print_input(ct.fitter.out_path, confs[i], "table_inp_-1.txt")
                reps.append(reporter.MultipleFitterReporter(ct.fitter, cp.fitter).open())
                pbs.append(problem_from(confs[i], tm))
            t1 = time.perf_counter()
            grid = pbs[0].hamil == "ntriv"
            b = (control.GridBatch if grid else control.NLevelBatch)(pbs, ctx=ctx)
            if grid:
                b.prewarm()
            t2 = time.perf_counter()
            with lock:
                tm_["task_setup_s"] += t1 - t0
                tm_["plan_setup_s"] += t2 - t1
                tm_["plans"] += 1
                if marks["first_ready"] is None:
                    marks["first_ready"] = t2
            return idxs, pbs, reps, b, grid, ctx
        except BaseException:
            alive.release()
            raise

    def consume(k, fut):
        """Control loops and reports of one prepared piece (runner thread)."""
        try:
            idxs, pbs, reps, b, grid, ctx = fut.result()      # a failed set-up has released its slot itself
        except BaseException:
            gates[k].set()
            raise
        try:
            if k > 0:
                gates[k - 1].wait()
            b.on_last_sweep = gates[k].set
            t_run = time.perf_counter()
            with lock:
                if marks["first_run"] is None:
                    marks["first_run"] = t_run
            if any(r.conf_rep_table.plotting_flag in (OT.ALL, OT.TABLES) for r in reps):
                # per-step tables were asked for: instrument the resident states at the steps the tables keep
                from .fitter import BatchObserver
                akx = None
                if grid:
                    akx = numpy.ascontiguousarray((math_base.initak(pbs[0].np_, pbs[0].dx, 1, pbs[0].ntriv) *
                                                   (phys_base.hart_to_cm / (-1j) / phys_base.dalt_to_au)).real)
                b.plan.set_instr_grid(numpy.asarray(pbs[0].x, numpy.float64), akx)
                b.observer = BatchObserver(b, reps)
            tt = pbs[0].task_type
            if tt == "optimal_control_krotov":
                hist = b.run_krotov()
            elif tt == "optimal_control_unit_transform":
                tables_iter = str(data_rep.get("fitter", {}).get("plotting_flag", "")).lower() == "tables_iter"
                hist = b.run_unit_transform(keep_fields=not tables_iter)
            elif tt == "local_control_population":
                hist = b.run_local_population()
            elif tt == "local_control_projection":
                hist = b.run_local_projection()
            else:
                hist = b.run_prescribed()
            errors = getattr(b, "errors", [""] * len(idxs))
            ctx.synchronize()
            t_rep = time.perf_counter()
            if b.observer is not None:
                b.observer.close()
            for j, i in enumerate(idxs):
                results[i] = hist[j]
                for r in hist[j]:
                    reps[j].print_iter_point_fitter(r.it, r.goal_close, None if r.E_tlist is None else list(r.E_tlist),
                                                    pbs[j].t_list, numpy.complex128(r.Fsm), numpy.float64(r.E_int), r.J,
                                                    pbs[j].nt)
                reps[j].close()
                if errors[j] and verbose:
                    print("An exception interrupted the job #%d: %s" % (i, errors[j]), file=sys.stderr)
                elif verbose:
                    print("Job #%d is done" % i)
            t_end = time.perf_counter()
            with lock:
                tm_["report_s"] += t_end - t_rep
                marks["last_run"] = max(marks["last_run"] or t_end, t_rep)
        finally:
            gates[k].set()
            b.close()
            alive.release()

    # Fresh device memory is expensive beside a running sweep (an allocation by the driver returns only when the sweep
    # ends): while the GPU is still idle, reserve the working set of the pipeline -- runners + 1 plans are alive at a
    # time -- as an arena of the library (qc_pool_reserve); it goes back when the last piece is done.
    reserved = False
    if len(work) > 1:
        need = 0
        for idxs in work:
            c0 = confs[idxs[0]]
            grid = c0.hamil_type == TaskRootConfiguration.HamilType.NTRIV
            nt_max = max(int(confs[i].fitter.propagation.nt) for i in idxs)
            need = max(need, _eng.plan_footprint(_eng.HAMIL_NONTRIV if grid else _eng.HAMIL_BH_X, c0.np, c0.nlevs, c0.nb,
                                                 c0.fitter.propagation.nch, len(idxs), nt_max,
                                                 bool(c0.fitter.hf_hide) and grid, _NAMES.get(c0.task_type) in control.TRAJ_TASKS))
        want = int((runners + 1.1) * need) + (64 << 20)
        if want <= 0.9 * ctxs[0].mem_info()[0]:
            try:
                ctxs[0].pool_reserve(want)
                reserved = True
            except Exception:
                reserved = False
    tm_["reserve_s"] = time.perf_counter() - t_mark          # cached blocks back to the driver, pieces planned, arena reserved
    t_pipe = time.perf_counter()
    prep_pool = concurrent.futures.ThreadPoolExecutor(1)
    run_pool = concurrent.futures.ThreadPoolExecutor(runners)
    try:
        preps = [prep_pool.submit(prepare, k) for k in range(len(work))]
        runs = [run_pool.submit(consume, k, f) for k, f in enumerate(preps)]
        for f in runs:
            f.result()
    finally:
        prep_pool.shutdown(wait=True, cancel_futures=True)
        run_pool.shutdown(wait=True, cancel_futures=True)
    if work:
        tm_["setup_wait_s"] = (marks["first_run"] or t_pipe) - t_pipe       # what the GPU waited for before its first sweep
        tm_["run_s"] = (marks["last_run"] or t_pipe) - (marks["first_run"] or t_pipe)
    t_mark = time.perf_counter()
    if reserved:
        ctxs[0].pool_reserve(0)
    for ctx in ctxs:
        ctx.close()
    tm_["teardown_s"] = time.perf_counter() - t_mark          # arena and contexts released
    if timings is not None:
        timings.update(tm_)
    return results


def main(argv):
    try:
        opts, _ = getopt.getopt(argv, "h", ["help", "json_task=", "json_rep=", "json_create", "seed_base="])
    except getopt.GetoptError:
        usage()
        return 2
    file_task = file_rep = None
    seed_base = 1000
    for o, a in opts:
        if o in ("-h", "--help"):
            usage()
            return 0
        if o == "--json_task":
            file_task = a
        elif o == "--json_rep":
            file_rep = a
        elif o == "--seed_base":
            seed_base = int(a)
    if not file_task or not file_rep:
        usage()
        return 2
    with open(file_task) as f:
        data_task = json.load(f)
    with open(file_rep) as f:
        data_rep = json.load(f)
    variants = qbatch.expand_substitutions(data_task)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    results = run_variants(variants, data_rep, seed_base)
    # the one collective of the path: every rank learns the final row of every run
    local = {i: numpy.array([rows[-1].it, rows[-1].goal_close, rows[-1].Fsm, rows[-1].E_int, rows[-1].J])
             for i, rows in results.items() if rows}
    table = qbatch.gather_results(local, len(variants), 5)
    if int(os.environ.get("RANK", "0")) == 0:
        print("run  iter  goal_close        Fsm            E_int          J")
        for i, row in enumerate(table):
            print("%4d %5.0f %12.6f %14.6f %14.4f %14.6f" % ((i,) + tuple(row)))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
