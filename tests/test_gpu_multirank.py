"""Two ranks on two GPUs over NCCL: the batch entry point with the runs of a ragged template sharded by
batch.lpt_shards and the final table through batch.gather_results, against the single-rank run of the same template
(VERDICT r01, next-1d).  Needs two CUDA devices (``gpurun --gpus 2``); skipped on a one-GPU box, where the same host
logic runs over gloo (tests/test_sharding_gloo.py)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %(root)r)
    import numpy as np
    import torch
    import torch.distributed as dist
    from qcontrol_b200 import batch, newcheb
    from tests.golden.gen_golden import conf_ut_jx, conf_krotov
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    out_dir = sys.argv[1]
    tables = []
    # (1) the T-sweep family of input_batch/inp_Jx_*: 2 run ids x 4 durations, nt_auto -> four different lengths
    tpl = conf_ut_jx(1, 3)
    tpl["run_id"] = {"@SUBST:LIST": ["r1", "r2"]}
    tpl["T"] = {"@SUBST:LIST": [2.5e-13, 6.1e-13, 4.0e-13, 3.3e-13]}
    tpl["nt_auto"] = True
    tpl["fitter"]["w_list"] = []
    tpl["fitter"]["h_lambda"] = 5e-4
    # (2) Krotov runs on the grid (np = 256), two amplitudes x two lengths through the same entry point
    tpk = conf_krotov(40, 2, np_=256)
    tpk["run_id"] = {"@SUBST:LIST": ["k1", "k2"]}
    tpk["fitter"]["propagation"]["E0"] = {"@SUBST:LIST": [2000.0, 2861.6]}
    for name, template in (("ut", tpl), ("krotov", tpk)):
        rep = {"fitter": {"out_path": os.path.join(out_dir, "w%%d" %% world, name + "_{input:$.run_id}", "T={input:$.T}_E0={input:$.fitter.propagation.E0}"),
                          "table.imin": -1, "plotting_flag": "tables_iter"}}
        variants = batch.expand_substitutions(template)
        results = newcheb.run_variants(variants, rep, verbose=False)
        local = {i: np.array([rows[-1].it, rows[-1].goal_close, rows[-1].Fsm, rows[-1].E_int, rows[-1].J])
                 for i, rows in results.items()}
        assert 0 < len(local) < len(variants) or world == 1, (rank, len(local))       # every rank owns a share
        table = batch.gather_results(local, len(variants), 5)
        assert table.shape == (len(variants), 5) and not np.isnan(table).any()
        tables.append(table)
    if rank == 0:
        np.savez(os.path.join(out_dir, "tables_w%%d.npz" %% world), ut=tables[0], krotov=tables[1])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
''') % {"root": ROOT}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_nccl_batch_entry_point(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    one = subprocess.run([sys.executable, str(script), str(tmp_path)], env=env, capture_output=True, text=True, timeout=900)
    assert one.returncode == 0, one.stdout[-2000:] + one.stderr[-3000:]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script), str(tmp_path)],
                         env=env, capture_output=True, text=True, timeout=900)
    assert two.returncode == 0, two.stdout[-2000:] + two.stderr[-3000:]
    a, b = np.load(tmp_path / "tables_w1.npz"), np.load(tmp_path / "tables_w2.npz")
    for key in ("ut", "krotov"):
        assert a[key].shape == b[key].shape
        # a run does not see its neighbours: the sharded result equals the single-rank one to the last bits
        assert np.allclose(a[key], b[key], rtol=1e-12, atol=0), (key, a[key] - b[key])
    # the per-run report files of both layouts: every variant wrote its tab_iter.csv under its own rank
    for w in (1, 2):
        found = [os.path.join(dp, f) for dp, _, fs in os.walk(tmp_path / ("w%d" % w)) for f in fs if f == "tab_iter.csv"]
        assert len(found) == 8 + 4, (w, len(found))
