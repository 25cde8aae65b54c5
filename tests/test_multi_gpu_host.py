"""Host-side logic of the N>1 path, on CPU: (1) the group -> device partition the library uses
(ldsr_shard_groups: whole groups, balanced, deterministic), (2) bench.py's rank protocol with
world_size 2 over gloo: each rank owns its own batch (weak scaling), the timing is the max over
ranks, the aggregate counts all ranks' fits."""
import os
import socket
import subprocess
import sys

import numpy as np

from ldsr_b200 import _lib, workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_groups_keeps_groups_whole_and_balanced():
    w = W.np_cv(37, 9)
    for n in (1, 2, 3, 8):
        sh = _lib.shard_groups(w["series"], w["group_series"], w["held"], w["fit_group"], w["theta0"], n)
        assert sh.shape == (37,) and sh.min() >= 0 and sh.max() < n
        counts = np.bincount(sh, minlength=n)
        assert counts.max() - counts.min() <= 1           # equal-cost groups -> even split
        assert np.array_equal(sh, _lib.shard_groups(w["series"], w["group_series"], w["held"], w["fit_group"],
                                                    w["theta0"], n))  # deterministic


def test_shard_groups_weights_by_cost():
    # two series of very different length: the long series' groups must not pile on one shard
    rng = np.random.default_rng(0)
    s_long = dict(y=rng.standard_normal(800), u=rng.standard_normal((3, 800)), v=None, q=1)
    s_short = dict(y=rng.standard_normal(50), u=rng.standard_normal((3, 50)), v=None, q=1)
    gs = [0, 0, 1, 1, 1, 1, 1, 1]
    fg = np.repeat(np.arange(8), 4)
    th = np.tile(np.array([0.5, 0.1, 0.1, 0.1, 0.5, 0.0, 1, 1, 0, 1.0]), (32, 1))
    sh = _lib.shard_groups([s_long, s_short], gs, None, fg, th, 2)
    assert sh[0] != sh[1]


_WORKER = r"""
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
w = bench.build_workload("np_restarts", rank)
nf = int(w["fit_group"].size)
# every rank draws different initial values (its own batch), same series
chk = torch.tensor([float(w["theta0"].sum())], dtype=torch.float64)
allchk = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
dist.all_gather(allchk, chk)
ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)   # pretend device time of this rank
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
tot = torch.tensor([nf], dtype=torch.int64)
dist.all_reduce(tot)
if rank == 0:
    print(json.dumps({"chk": [float(c) for c in allchk], "max_ms": float(ms), "fits": int(tot),
                      "value": world * nf / (float(ms) * 1e-3)}))
dist.barrier()
dist.destroy_process_group()
"""


def test_bench_rank_protocol_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    import json
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["chk"][0] != res["chk"][1]          # ranks own different batches
    assert res["max_ms"] == 15.0                   # max over ranks, not rank 0's own
    assert res["fits"] == 200 and abs(res["value"] - 200 / 0.015) < 1e-6


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_tbrm_and_stored_cv_aggregate():
    """The stored cvLDS result aggregates metrics.dist with the plain mean (R/sysdata.rda::NPcv);
    api.tbrm (dplR's robust mean, unpinned) is at least a robust location: it ignores a gross outlier
    and equals the mean on symmetric data."""
    import json
    import os
    import numpy as np
    from ldsr_b200 import api
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "npcv.json")))
    for k in ("R2", "RE", "CE", "nRMSE", "KGE"):
        assert abs(np.mean(g["metrics_dist"][k]) - g["metrics"][k][0]) < 1e-12
    x = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    assert abs(api.tbrm(x) - 3.0) < 1e-12
    assert abs(api.tbrm(np.append(x, 1e6)) - api.tbrm(np.append(x, 3.5))) < 0.5
    assert np.isnan(api.tbrm([np.nan]))


def test_take_groups_partitions_a_job_by_shard():
    """bench.py's `strong` block: every rank takes the groups ldsr_shard_groups gives it; together the
    shards hold every fit exactly once, in order, with local group ids."""
    w = W.synthetic_stations(n_stations=3, T=120, p=2, n_folds=5, n_restarts=4)
    n = 4
    sh = _lib.shard_groups(w["series"], w["group_series"], w["held"], w["fit_group"], w["theta0"], n)
    seen = []
    for r in range(n):
        sub = W.take_groups(w, np.nonzero(sh == r)[0])
        assert np.all(np.diff(sub["fit_group"]) >= 0) and sub["fit_group"].max() == len(sub["group_series"]) - 1
        assert np.array_equal(sub["theta0"], w["theta0"][sub["fits"]])
        assert np.array_equal(sub["groups"][sub["fit_group"]], w["fit_group"][sub["fits"]])
        assert all(np.array_equal(a, w["held"][g]) for a, g in zip(sub["held"], sub["groups"]))
        seen.append(sub["fits"])
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(w["fit_group"].size))
