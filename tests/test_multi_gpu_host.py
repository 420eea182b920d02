"""N > 1 path on CPU: env sharding, the deterministic ctrl stream and the end-of-rollout metrics
gather with the gloo backend, world size 2.  No GPU needed."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nenv = 8                                  # envs per rank (weak scaling: rank r owns [r*nenv, (r+1)*nenv))
    lo, hi = np.array([-6.0, 0.0, -1.0]), np.array([6.0, 1.1, 1.0])
    mine = bench.ctrl_np(0, rank * nenv, nenv, 3, lo, hi)
    whole = bench.ctrl_np(0, 0, world * nenv, 3, lo, hi)
    assert np.array_equal(mine, whole[rank * nenv:(rank + 1) * nenv])      # sharding never changes an env's stream
    t = bench.ctrl_torch(0, rank * nenv, nenv, 3, torch.tensor(lo), torch.tensor(hi), "cpu").numpy()
    assert np.abs(t - mine.astype(np.float32)).max() == 0.0               # host and device generators agree bit for bit
    assert (mine >= lo).all() and (mine <= hi).all()
    metrics = torch.tensor([float(nenv * 50), 10.0 + rank, float(mine.sum()), 5.0 * nenv, 0.0])
    gathered = torch.empty(world * 5)
    dist.all_gather_into_tensor(gathered, metrics)
    gathered = gathered.view(world, 5)
    tmax = torch.tensor([10.0 + rank]); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put((gathered.numpy().copy(), float(tmax)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_rollout_metrics_gloo():
    world, port = 2, 29500 + (os.getpid() % 500)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert gathered.shape == (2, 5) and gathered[:, 0].sum() == 2 * 8 * 50
    assert tmax == 11.0                       # throughput uses the max over ranks
    assert gathered[0, 1] == 10.0 and gathered[1, 1] == 11.0


def test_reference_arm_runs_on_rank0_only(monkeypatch, capsys):
    import bench
    monkeypatch.setenv("RANK", "1")
    class A: gpus = 2; steps = 1; warmup = 3; cpu_nenv = 4; nenv = 4; config = "cfg2"; scaling = "weak"
    bench.run_reference(A, bench.CONFIGS["cfg2"])
    assert capsys.readouterr().out == ""      # other ranks exit without work


def test_reference_arm_line_and_strong_scaling_split(monkeypatch, capsys):
    """Rank 0 of the reference arm prints the contract line; strong scaling splits --nenv over the ranks."""
    import json
    import bench
    monkeypatch.setenv("RANK", "0")
    class A: gpus = 1; steps = 1; warmup = 3; cpu_nenv = 8; nenv = 8; config = "cfg2"; scaling = "strong"
    bench.run_reference(A, bench.CONFIGS["cfg2"])
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["scaling"] == "strong" and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
    assert set(bench.CONFIGS) == {"cfg2", "default", "cfg3", "cfg4", "cfg5"}
    for name, cfg in bench.CONFIGS.items():
        assert os.path.exists(os.path.join(bench.GOLDEN_DIR, cfg["blob"])), name
