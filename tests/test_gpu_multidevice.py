"""Multi-GPU correctness ON HARDWARE (needs >= 2 B200s: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multidevice.py -m gpu`;
skipped on a one-GPU box).

  * two contexts on two devices inside ONE process (the header's "one context per task / GPU"): every kernel's
    per-device function attributes are configured on both devices and the results are identical to one device;
  * two ranks (torchrun-style, one process per GPU): the ABI's own NCCL path -- dqmc_comm_unique_id / dqmc_comm_init /
    dqmc_reduce_observables on the context's stream -- gives a reduced block equal to the sum of the per-rank blocks,
    and `chain_offset` gives every global chain index its own Philox stream (disjoint streams, reproducible sharding).
"""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _need_two(b200):
    n = b200._lib.load().dqmc_device_count()
    if n < 2:
        pytest.skip(f"needs 2 GPUs, found {n}")


def _mk(b200, device, chain_offset, B, Ls=(12, 12), U=-4.0, beta=0.5, seed=5):
    model = b200.HubbardModel(b200.SquareLattice(Ls[0]), U=U)
    mc = b200.DQMC(model, beta=beta, delta_tau=0.1, safe_mult=5, seed=seed, n_chains=B, device=device,
                   chain_offset=chain_offset)
    return mc


def test_two_contexts_on_two_devices_in_one_process(b200):
    _need_two(b200)
    B = 3
    g = np.random.default_rng(1)
    mcs = [_mk(b200, d, 0, B) for d in (0, 1)]
    N, M = mcs[0].ctx.N, mcs[0].ctx.M
    conf = np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(N, M, B)))
    out = []
    for mc in mcs:                       # n = 144: update3 (225 KB dynamic smem), cluster QR, form-Q, GEMM -- all opt-in smem
        mc.ctx.set_conf(conf)
        mc.ctx.build_stack()
    for _ in range(2):                   # interleave the two devices
        for mc in mcs:
            mc.ctx.sweep(1)
    for mc in mcs:
        out.append((mc.ctx.greens(), mc.ctx.get_conf(), mc.ctx.kernel_launches()))
    assert np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][0], out[1][0])          # same kernels, same inputs: bit-identical
    assert out[0][2] == out[1][2] > 0                    # per-context launch counters


_RANK_SCRIPT = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, os.environ["DQMC_ROOT"])
import torch, torch.distributed as dist
import _b200_loader
pkg = _b200_loader.load()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo")                           # only to ship the 128-byte NCCL id; the data path is the ABI's
B = 3
model = pkg.HubbardModel(pkg.SquareLattice(4), U=-4.0)
mc = pkg.DQMC(model, beta=1.0, delta_tau=0.1, safe_mult=5, seed=9, n_chains=B, device=rank, chain_offset=rank * B)
ctx = mc.ctx
N, M = ctx.N, ctx.M
conf = np.stack([np.random.default_rng(100 + rank * B + b).choice(np.array([-1, 1], dtype=np.int8), size=(N, M))
                 for b in range(B)], axis=2)
ctx.set_conf(np.asfortranarray(conf))
ctx.build_stack()
acc = ctx.sweep(1)
ctx.accumulate_greens()
cnt, s, s2 = ctx.observables()
local = np.concatenate([[cnt], s.ravel(order="F"), s2.ravel(order="F")])
ids = [pkg.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx.comm_init(world, rank, ids[0])
ctx.reduce_observables()                                  # ncclAllReduce on the context's stream
cnt, s, s2 = ctx.observables()
red = np.concatenate([[cnt], s.ravel(order="F"), s2.ravel(order="F")])
gathered = [None] * world
dist.all_gather_object(gathered, (local, red, acc.tolist(), ctx.get_conf().tolist()))
if rank == 0:
    json.dump({"local": [g[0].tolist() for g in gathered], "red": [g[1].tolist() for g in gathered],
               "acc": [g[2] for g in gathered], "conf": [g[3] for g in gathered]}, open(os.environ["DQMC_OUT"], "w"))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_ranks_abi_nccl_allreduce_and_disjoint_streams(b200, tmp_path):
    _need_two(b200)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT)
    out = tmp_path / "out.json"
    env = dict(os.environ, DQMC_ROOT=str(ROOT), DQMC_OUT=str(out))
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                    "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                   check=True, env=env, timeout=600)
    import json
    r = json.loads(out.read_text())
    local = np.array(r["local"]); red = np.array(r["red"])
    assert np.allclose(red[0], local.sum(axis=0), rtol=1e-14, atol=1e-14)     # reduced block == sum of the shards
    assert np.array_equal(red[0], red[1])                                     # every rank holds the same result
    assert red[0][0] == 6                                                     # count = all chains
    # chain_offset: global chain g = rank * B + b takes the decisions of an oracle chain with chain_id = g
    from oracle import model as OM, ref as OR
    T = OM.hopping_matrix("square", (4, 4))
    for rank in range(2):
        for b in range(3):
            gidx = rank * 3 + b
            conf0 = np.random.default_rng(100 + gidx).choice(np.array([-1, 1], dtype=np.int8), size=(16, 10))
            c = OR.RefChain(T, U=-4.0, beta=1.0, safe_mult=5, seed=9, chain_id=gidx, conf=np.asfortranarray(conf0))
            c.init()
            assert c.local_sweep() == r["acc"][rank][b]
            assert np.array_equal(np.array(r["conf"][rank], dtype=np.int8)[:, :, b], c.get_conf())
