"""Multi-process host logic on CPU (gloo, world_size = 2): every rank derives the same LPT partition
without communicating, owns a disjoint share of the runs, and the single final collective
(batch.gather_results) hands every rank the complete per-run table.  No GPU involved: the per-run
"result" is a deterministic function of the run index."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np
    import torch.distributed as dist
    from qcontrol_b200 import batch
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    tpl = {"run_id": {"@SUBST:LIST": ["r%%d" %% i for i in range(5)]}, "T": {"@SUBST:LIST": [2.5e-13, 11.8e-12, 19.0e-12]}}
    variants = batch.expand_substitutions(tpl)
    costs = [v["T"] / 2e-15 for v in variants]
    shards = batch.lpt_shards(costs, world)
    mine = shards[rank]
    local = {i: np.array([i, costs[i], 7.0 * i + 1.0]) for i in mine}
    table = batch.gather_results(local, len(variants), 3)
    assert not np.isnan(table).any()
    assert np.array_equal(table[:, 0], np.arange(len(variants)))
    assert np.allclose(table[:, 2], 7.0 * np.arange(len(variants)) + 1.0)
    loads = [sum(costs[i] for i in s) for s in shards]
    print("rank", rank, "runs", len(mine), "balance", max(loads) / min(loads))
    assert max(loads) / min(loads) < 1.15
    dist.barrier()
    dist.destroy_process_group()
''') % ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_sharding_and_final_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=180)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for p, out in zip(procs, outs):
        assert p.returncode == 0, out
    assert "rank 0" in outs[0] and "rank 1" in outs[1]
