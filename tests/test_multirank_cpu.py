"""World-size-2 gloo test of the N > 1 host logic (no GPU): chains are sharded by rank with
`chain_offset`, there is no data-path collective, and the only exchange is a sum all-reduce of the
observable block {count, sum, sumsq}.  The per-rank 'device' work is played by the CPU oracle here."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from oracle import model as OM, ref as OR
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, L, U, beta = 3, 4, -4.0, 0.5                      # chains per rank
    T = OM.hopping_matrix("square", (L, L)); M = OM.n_slices(beta)
    obs = np.zeros(1 + 2 * 2 * 16 * 16)
    for b in range(B):
        gidx = rank * B + b                              # chain_offset = rank * B
        g = np.random.default_rng(100 + gidx)
        c = OR.RefChain(T, U=U, beta=beta, seed=9, chain_id=gidx,
                        conf=np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(16, M))))
        c.init(); c.local_sweep()
        G = c.measured_greens().ravel(order="F")
        obs[0] += 1; obs[1:1 + G.size] += G; obs[1 + G.size:] += G * G
    t = torch.from_numpy(obs)
    dist.all_reduce(t)                                   # the path's single collective
    q.put((rank, t.numpy().copy()))
    dist.barrier(); dist.destroy_process_group()


def test_sharded_chains_and_observable_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert np.array_equal(res[0], res[1])                # both ranks hold the reduced block
    assert res[0][0] == 6                                # 2 ranks x 3 chains
    # single-process evaluation of the same 6 global chains gives the same sums
    sys.path.insert(0, str(ROOT))
    from oracle import model as OM, ref as OR
    T = OM.hopping_matrix("square", (4, 4)); M = OM.n_slices(0.5)
    tot = np.zeros_like(res[0])
    for gidx in range(6):
        g = np.random.default_rng(100 + gidx)
        c = OR.RefChain(T, U=-4.0, beta=0.5, seed=9, chain_id=gidx,
                        conf=np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(16, M))))
        c.init(); c.local_sweep()
        G = c.measured_greens().ravel(order="F")
        tot[0] += 1; tot[1:1 + G.size] += G; tot[1 + G.size:] += G * G
    assert np.allclose(res[0], tot, rtol=1e-13, atol=1e-13)


def test_bench_reference_arm_runs_on_cpu():
    """bench.py --impl reference times the oracle port on the host cores and prints the contract line."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--config", "cfg1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "sweeps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
