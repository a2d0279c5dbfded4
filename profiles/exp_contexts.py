"""Experiment: K independent contexts (own stream each) on ONE GPU, driven from K host threads, against one context with all
chains.  Chains never interact, so a sweep of context A can overlap with context B's: latency-bound kernels (UDT steps, the
serial phase of the local update) leave issue slots and whole SMs idle that the other context's DMMA GEMMs can use.

    python profiles/exp_contexts.py [--config cfg4] [--chains 148] [--contexts 2] [--sweeps 3]
"""
import argparse
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _b200_loader  # noqa: E402
import bench  # noqa: E402  (CONFIGS only)

pkg = _b200_loader.load()


def make(kind, Ls, U, beta, B, offset):
    lattice = {"square": pkg.SquareLattice, "honeycomb": pkg.Honeycomb, "chain": pkg.Chain}[kind](*Ls[:1])
    mc = pkg.DQMC(pkg.HubbardModel(lattice, U=U), beta=beta, delta_tau=bench.DELTA_TAU, safe_mult=bench.SAFE_MULT,
                  seed=bench.SEED, n_chains=B, chain_offset=offset, device=0)
    g = np.random.default_rng(bench.SEED + offset)
    mc.ctx.set_conf(np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(mc.ctx.N, mc.ctx.M, B))))
    mc.ctx.build_stack()
    return mc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg4")
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--contexts", type=int, default=2)
    ap.add_argument("--sweeps", type=int, default=3)
    a = ap.parse_args()
    import torch
    kind, Ls, U, beta, Bdef, _ = bench.CONFIGS[a.config]
    B = a.chains or Bdef
    per = [B // a.contexts + (1 if i < B % a.contexts else 0) for i in range(a.contexts)]
    mcs, off = [], 0
    for b in per:
        mcs.append(make(kind, Ls, U, beta, b, off)); off += b

    def run(mc, n):
        for _ in range(n):
            mc.ctx.sweep(1)

    def timed(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=run, args=(mc, n)) for mc in mcs]
        [t.start() for t in th]; [t.join() for t in th]
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    timed(3)
    dt = timed(a.sweeps)
    print(f"{a.config}: {a.contexts} context(s) x {per} chains: {B * a.sweeps / dt:.2f} sweeps/s ({1e3 * dt / a.sweeps:.1f} ms per sweep)",
          flush=True)


if __name__ == "__main__":
    main()
