"""CPU, world_size = 2 over gloo: the host-side logic of the multi-GPU path -- unique-id exchange, the
sharding arithmetic the CUDA kernels implement (gradient = all-reduced sum scaled by 1/(B_local*world); whiten from
all-reduced (Σx, Σx², n)), and bench.py's torchrun contract for the reference arm."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import crux_b200 as crux
        from crux_b200.device import exchange_unique_id, shard_seed
        from oracle import crux_oracle as o
        # 1. every rank ends up with rank 0's NCCL unique id
        uid = exchange_unique_id(rank)
        ids = [None] * world
        dist.all_gather_object(ids, uid)
        assert len(uid) == 128 and all(x == ids[0] for x in ids) and any(b != 0 for b in uid)
        assert shard_seed(7, 0) != shard_seed(7, 1)
        # 2. data-parallel PPO gradient: local loss scaled by 1/(B_local*world), summed over ranks == full-batch gradient
        rng = np.random.default_rng(0)  # same data on both ranks, each takes its shard
        n = 256
        mu = o.MLP([17, 64, 64, 6], [1, 1, 0], np.random.default_rng(1))
        pi = o.GaussianPolicy(mu, np.full(6, -0.5, np.float32))
        s = rng.standard_normal((n, 17)).astype(np.float32)
        a, lp = pi.exploration(s, rng.standard_normal((n, 6)).astype(np.float32))
        D = {"s": s, "a": a.detach().numpy(), "logprob": lp.detach().numpy()[:, 0] - np.float32(0.03),
             "advantage": rng.standard_normal(n).astype(np.float32), "return": rng.standard_normal(n).astype(np.float32)}
        P = {"eps": np.float32(0.2), "lp": np.float32(1), "le": np.float32(0.0)}
        for p_ in pi.params():
            p_.grad = None
        o.ppo_loss(pi, P, D).backward()
        full = o.flat_grads(pi.params())
        sh = slice(rank * n // world, (rank + 1) * n // world)
        Dl = {k: v[sh] for k, v in D.items()}
        for p_ in pi.params():
            p_.grad = None
        (o.ppo_loss(pi, P, Dl) / world).backward()   # mean over B_local, divided by world == sum / (B_local*world)
        g = torch.from_numpy(o.flat_grads(pi.params()))
        dist.all_reduce(g)
        assert np.allclose(g.numpy(), full, rtol=1e-4, atol=1e-7)
        # 3. global whitening from all-reduced moments (crux_whiten with NCCL): identical on every rank, equals whiten(all)
        x = rng.standard_normal(1000).astype(np.float32) * 3 + 1
        xl = x[rank::world].astype(np.float64)
        st = torch.tensor([xl.sum(), (xl * xl).sum(), float(len(xl))], dtype=torch.float64)
        dist.all_reduce(st)
        mean = st[0] / st[2]
        var = (st[1] - st[2] * mean * mean) / (st[2] - 1)
        got = (xl - float(mean)) / float(var.sqrt())
        assert np.allclose(got, o.whiten(x)[rank::world], rtol=1e-4, atol=1e-5)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_host_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_reference_arm_under_torchrun():
    """`torchrun --nproc-per-node 2 bench.py --impl reference --gpus 2`: rank 0 alone runs and prints; rank 1 exits 0."""
    env = dict(os.environ, CRUX_BENCH_TINY="1")
    port = 29600 + os.getpid() % 300
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["n_gpus"] == 2 and rec["value"] > 0
