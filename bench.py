#!/usr/bin/env python
"""bench.py -- DQMC sweeps/sec (all chains) on N B200s, with roofline, CPU baseline and e2e.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg4] [--chains B] [--impl reference]

One "step" = one full local sweep (2M slice visits x N proposals, all wraps and stabilisations,
reference local_updates.jl:7-14) of every chain of the context.  Default workload = BASELINE.json
configs[3] ("cfg4": repulsive Hubbard 16x16, beta=16, dtau=0.1 -- the configuration the north_star
target is quoted on); the other configs are selectable and are parity-test cases.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CONFIGS = {
    # name: (lattice, Ls, U, beta, chains per GPU, description)
    "cfg1": ("square", (4, 4), 4.0, 5.0, 1, "attractive Hubbard 4x4, U=4, beta=5, single chain"),
    "cfg2": ("square", (8, 8), 4.0, 10.0, 256, "attractive Hubbard 8x8, U=4, beta=10, 256 chains/GPU"),
    "cfg3": ("square", (12, 12), -4.0, 8.0, 128, "repulsive Hubbard 12x12, U=-4, beta=8, 128 chains/GPU"),
    "cfg4": ("square", (16, 16), -4.0, 16.0, 148, "repulsive Hubbard 16x16, U=-4, beta=16, dtau=0.1 (N=256, M=160)"),
    "cfg5": ("honeycomb", (12, 12), 4.0, 10.0, 148, "attractive Hubbard honeycomb L=12 (N=288), U=4, beta=10"),
    # the only timings the reference's docs state (docs/src/DQMC/fields.md:49-56, ~1470 and ~588 sweeps/s on unstated
    # hardware): run with --impl reference to anchor the CPU port against them
    "anchor6a": ("square", (6, 6), 1.0, 1.0, 256, "anchor: attractive Hubbard 6x6, U=1, beta=1 (DensityHirschField)"),
    "anchor6r": ("square", (6, 6), -1.0, 1.0, 256, "anchor: repulsive Hubbard 6x6, U=-1, beta=1 (MagneticHirschField)"),
}
DELTA_TAU, SAFE_MULT, SEED = 0.1, 10, 1234


def flops_per_sweep(n, M, C, nb, acc_rate):
    """F_sweep = nb n^3 (12 M + 44 C + 4 a M) [+ 4 C n^3 nb for the check wrap]."""
    return nb * float(n) ** 3 * (12 * M + 44 * C + 4 * acc_rate * M + 4 * C)


def build_model(cfg):
    kind, Ls, U, beta, B, desc = CONFIGS[cfg]
    return kind, Ls, U, beta, B, desc


def config_dict(cfg, B, world):
    """The `config` object of the JSON line: identical for the GPU arm and the reference arm (it names the workload;
    what the CPU arm actually sampled is described in its cpu_baseline.sample)."""
    kind, Ls, U, beta, _, desc = CONFIGS[cfg]
    nbasis = 2 if kind == "honeycomb" else 1
    N = nbasis * int(np.prod(Ls if len(Ls) > 1 else Ls))
    M = int(round(beta / DELTA_TAU))
    nb = 2 if U < 0 else 1
    C = -(-M // SAFE_MULT)
    return {"workload": f"{cfg}: {desc}", "chains_per_gpu": B, "n_sites": N, "n_slices": M,
            "flavor_blocks": nb, "delta_tau": DELTA_TAU, "safe_mult": SAFE_MULT, "U": U,
            "l2": "inputs larger than L2 (per-GPU state %.1f GB)" % (ctx_bytes(N, M, C, nb, B) / 1e9),
            "parallelism": f"chains sharded over {world} GPU(s), no data-path collective"}


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------
def run_cpu(cfg, steps, warmup, max_chains=None):
    from oracle import model as OM, ref as OR
    kind, Ls, U, beta, _, desc = build_model(cfg)
    T = OM.hopping_matrix(kind, Ls)
    N, M = T.shape[0], OM.n_slices(beta, DELTA_TAU)
    cores = OR.lib().ref_max_threads()
    nch = cores if max_chains is None else min(cores, max_chains)
    g = np.random.default_rng(SEED)
    chains = [OR.RefChain(T, U=U, beta=beta, delta_tau=DELTA_TAU, safe_mult=SAFE_MULT, seed=SEED, chain_id=b,
                          conf=np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(N, M))))
              for b in range(nch)]
    dt, acc = OR.run_chains(chains, nthreads=nch, warm=warmup, nsweeps=steps)
    a = float(acc.sum()) / (steps * nch * 2 * N * M)
    return {"value": nch * steps / dt, "seconds": dt, "cores": nch, "chains": nch, "sweeps_per_chain": steps,
            "acceptance": a, "N": N, "M": M, "nb": chains[0].nb, "C": chains[0].C}


def run_cpu_measure(cfg):
    """The oracle's CPU time for one TimeIntegral measurement pass (CombinedGreensIterator + Wick kernels), one
    chain per host core."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import measure as OMS, model as OM, ref as OR
    kind, Ls, U, beta, _, desc = build_model(cfg)
    T = OM.hopping_matrix(kind, Ls)
    N, M = T.shape[0], OM.n_slices(beta, DELTA_TAU)
    cores = OR.lib().ref_max_threads()
    g = np.random.default_rng(SEED)
    chains = [OR.RefChain(T, U=U, beta=beta, delta_tau=DELTA_TAU, safe_mult=SAFE_MULT, seed=SEED, chain_id=b,
                          conf=np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(N, M))))
              for b in range(cores)]
    s2d, nbasis = OMS.bravais_srctrg2dir(Ls), OM.UNIT_CELLS[kind][0]

    def one(c):
        c.init()
        t0 = time.time()
        G00 = c.measured_greens()
        OMS.time_integral(G00, c.combined_greens_iterator(), c.delta_tau, c.M, s2d, nbasis)
        return time.time() - t0

    t0 = time.time()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        per = list(ex.map(one, chains))
    return {"time_integral_s_per_chain": float(np.mean(per)), "cores": cores, "chains": cores, "wall_s": time.time() - t0,
            "time_integral_passes_per_s": cores / float(np.max(per)), "kind": "port",
            "sample": f"{cores} chains (one per host core), one TimeIntegral pass each; oracle C iterator + numpy Wick kernels"}


def make_mc(pkg, cfg, B, dev, chain_offset=0, seed=SEED, rank=0):
    """The product's host mirror builds every input of a configuration (nothing under oracle/ is touched)."""
    kind, Ls, U, beta, _, _ = build_model(cfg)
    lattice = {"square": pkg.SquareLattice, "honeycomb": pkg.Honeycomb, "chain": pkg.Chain}[kind](*Ls[:1])
    model = pkg.HubbardModel(lattice, U=U)
    mc = pkg.DQMC(model, beta=beta, delta_tau=DELTA_TAU, safe_mult=SAFE_MULT, seed=seed, n_chains=B,
                  chain_offset=chain_offset, device=dev)
    g = np.random.default_rng(seed + rank)
    conf = np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(mc.ctx.N, mc.ctx.M, B)))
    mc.ctx.set_conf(conf)
    return mc, lattice, conf


def quick_config(pkg, torch, cfg, dev, rank, world, nsweeps, max_over_ranks):
    """Device-resident sweeps/s of one of the other BASELINE configurations (same timing rules, fewer sweeps)."""
    B = CONFIGS[cfg][4]
    mc, _, _ = make_mc(pkg, cfg, B, dev, chain_offset=rank * B, rank=rank)
    ctx = mc.ctx
    ctx.build_stack()
    for _ in range(3):
        ctx.sweep(1)
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.kernel_launches()
    e0.record(stream)
    acc = 0
    for _ in range(nsweeps):
        acc += int(ctx.sweep(1).sum())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1))
    out = {"workload": f"{cfg}: {CONFIGS[cfg][5]}", "chains_per_gpu": B, "chains_total": B * world, "steps": nsweeps,
           "value": world * B * nsweeps / (ms * 1e-3), "unit": "sweeps/s", "ms_per_step": ms / nsweeps,
           "acceptance": acc / (B * nsweeps * 2.0 * ctx.N * ctx.M), "gpu_launches_per_step": (ctx.kernel_launches() - l0) / nsweeps}
    ctx.close()
    return out


def parity_leg(pkg, cfg, dev):
    """cpu_baseline leg only (the one place bench.py may execute oracle/): ONE oracle chain x ONE full sweep of the bench
    workload against the device under the shared counter RNG -- decisions, G at sweep end (vs the oracle and vs the
    extended-precision arbiter oracle/truth_ld.c), propagation-error counts.  The oracle here is the checker."""
    from oracle import model as OM, ref as OR, truth as TR
    kind, Ls, U, beta, _, _ = build_model(cfg)
    mc, _, conf = make_mc(pkg, cfg, 2, dev, seed=SEED + 1)
    ctx = mc.ctx
    ctx.build_stack()
    c = OR.RefChain(OM.hopping_matrix(kind, Ls), U=U, beta=beta, delta_tau=DELTA_TAU, safe_mult=SAFE_MULT,
                    seed=SEED + 1, chain_id=0, conf=conf[:, :, 0])
    c.init()
    t0 = time.time()
    a_ref, p_ref, d_ref = c.local_sweep(trace=True)
    t_oracle = time.time() - t0
    acc, probs, dec = ctx.sweep_traced()
    G = ctx.greens()[:, :, :, 0]
    truth = TR.greens_truth_chain(c, chunk=10)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    st = ctx.stats()[0]
    out = {"config": cfg, "decisions": int(dec[0].size), "decisions_identical": bool(np.array_equal(dec[0], d_ref)),
           "decision_mismatches": int((dec[0] != d_ref).sum()), "accepted_dev": int(acc[0]), "accepted_oracle": int(a_ref),
           "conf_identical": bool(np.array_equal(ctx.get_conf()[:, :, 0], c.get_conf())),
           "G_dev_vs_oracle": rel(G, c.greens), "G_dev_vs_truth": rel(G, truth), "G_oracle_vs_truth": rel(c.greens, truth),
           "tolerance": 1e-10, "prop_count_dev": int(st["prop_count"]), "prop_count_oracle": int(c.stats["prop_count"]),
           "oracle_sweep_s": t_oracle,
           "truth": "oracle/truth_ld.c: x87 long double, independent stabilisation (role of BigFloat in the reference's tests)"}
    ctx.close()
    return out


# ------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.device)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.split(",") for l in Path(self.f.name).read_text().splitlines() if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].strip().lower() == "active":
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default=os.environ.get("DQMC_BENCH_CONFIG", "cfg4"), choices=sorted(CONFIGS))
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (0 = config default)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--delay-block", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-pass", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short device-resident runs of the other BASELINE configurations")
    ap.add_argument("--measure", action="store_true",
                    help="also time one equal-time + one TimeIntegral measurement pass (device Wick kernels and the "
                         "CombinedGreensIterator) and the oracle's CPU time for the same; default for cfg5")
    args = ap.parse_args()

    # stdout carries exactly ONE line, the JSON record: everything a library prints on fd 1 on the way (NCCL announces its
    # version there when torch.distributed or the library's own communicator comes up) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    kind, Ls, U, beta, Bdef, desc = build_model(args.config)
    B = args.chains or Bdef

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, args.steps), max(0, args.warmup)
        r = run_cpu(args.config, steps, warm)
        line = {"impl": "reference", "metric": "DQMC sweeps/sec (all chains)", "value": r["value"], "unit": "sweeps/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * r["seconds"] / steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(args.config, B, args.gpus),
                "cpu_baseline": {"value": r["value"], "unit": "sweeps/s", "cores": r["cores"], "kind": "port",
                                 "sample": f"{r['chains']} chains (one per host core) x {steps} sweep(s) after {warm} warm-up "
                                           "sweep(s); a step = one sweep of every sampled chain; C port of the "
                                           "reference's loops (oracle/dqmc_ref.c) -- Julia is not in the image"},
                "e2e": {"value": r["value"], "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "acceptance": r["acceptance"], "gpu_launches": 0}
        emit(line)
        return

    import torch
    import _b200_loader
    pkg = _b200_loader.load()

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the dqmc_b200 path has no CPU fallback")
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))

    # the product's own host mirror builds every input (lattice, hopping matrix, exponentials, ranges, alpha):
    # nothing under oracle/ is touched on this arm outside the cpu_baseline leg
    lattice = {"square": pkg.SquareLattice, "honeycomb": pkg.Honeycomb, "chain": pkg.Chain}[kind](*Ls[:1])
    model = pkg.HubbardModel(lattice, U=U)
    mc = pkg.DQMC(model, beta=beta, delta_tau=DELTA_TAU, safe_mult=SAFE_MULT, seed=SEED, n_chains=B,
                  chain_offset=rank * B, device=dev, delay_block=args.delay_block)
    ctx = mc.ctx
    T = mc.hopping_matrix
    N, M, nb, C = ctx.N, ctx.M, ctx.nb, ctx.C
    g = np.random.default_rng(SEED + rank)
    conf = np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(N, M, B)))
    ctx.set_conf(conf)
    ctx.build_stack()
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=dev)
    if world > 1:
        # the path's only collective (final observable reduction) goes through the library's own NCCL communicator
        # on the context's stream; torch.distributed only ships the 128-byte id and does the timing barriers
        ids = [pkg.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(world, rank, ids[0])
        ctx.reduce_observables()             # first collective of a communicator sets up its connections: not timed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def sum_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t.item())
        return x

    W, K = max(3, args.warmup), max(1, args.steps)
    for _ in range(W):
        ctx.sweep(1)

    # ---------------- timed region 1: device-resident sweeps ---------------------------------
    clocks = Clocks(dev)
    barrier()
    if rank == 0:
        clocks.start()
    l0 = ctx.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    acc_total = 0
    for _ in range(K):
        acc_total += int(ctx.sweep(1).sum())
    ctx.accumulate_greens()
    if world > 1:      # the only collective of the path: final observable reduction over NVLink (dqmc_reduce_observables)
        ctx.reduce_observables()
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.kernel_launches() - l0
    clk = clocks.stop() if rank == 0 else None
    acc_rate = sum_over_ranks(acc_total) / (world * B * K * 2.0 * N * M)
    value = world * B * K / (ms * 1e-3)

    # ---------------- timed region 2: end to end through the C ABI with host buffers ---------
    u_host = torch.empty((B, 2 * M, N), dtype=torch.float64).pin_memory()
    g_host = torch.empty((B, nb, N, N), dtype=torch.float64).pin_memory()
    c_host = torch.empty((B, M, N), dtype=torch.int8).pin_memory()
    rng_t = torch.Generator().manual_seed(SEED + 17 * rank)
    u_host.uniform_(generator=rng_t)     # the host's own RNG stream (drop-in mode: Julia supplies the uniforms)
    for _ in range(min(W, 2)):           # untimed: the table-RNG mode has its own captured graph (first sweep eager, second captures)
        ctx.sweep(1, uniforms=u_host.numpy())
        ctx.greens(out=g_host.numpy())
        ctx.get_conf(out=c_host.numpy())
    barrier()
    e2e0, e2e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e0.record(stream)
    for _ in range(K):
        ctx.sweep(1, uniforms=u_host.numpy())                    # H2D of this sweep's uniforms inside the call
        ctx.greens(out=g_host.numpy())                           # D2H: mc.stack.greens of every chain
        ctx.get_conf(out=c_host.numpy())                         # D2H: field.conf of every chain
    e2e1.record(stream)
    barrier()
    ms_e2e = max_over_ranks(e2e0.elapsed_time(e2e1))
    e2e_value = world * B * K / (ms_e2e * 1e-3)
    h2d = u_host.numel() * 8
    d2h = g_host.numel() * 8 + c_host.numel() + B * 8

    # ---------------- roofline pass: per-launch CUDA events on the launching stream -----------
    roof = None
    if not args.no_kernel_pass:
        ctx.profile(True)
        for _ in range(K):
            ctx.sweep(1)
        prof = ctx.profile_report()
        ctx.profile(False)
        p64 = measure_fp64_peak(torch) if rank == 0 else None
        same_shape = measure_cublas_batched(torch, N, B * nb) if rank == 0 else None
        if rank == 0:
            gm = prof["gemm"]
            flops_per_launch = 2.0 * N ** 3 * B * nb        # every GEMM launch of the sweep is n x n x n over all matrices
            avg_ms = gm["ms"] / max(gm["count"], 1)
            achieved = flops_per_launch / (avg_ms * 1e-3) / 1e12
            total_ms = sum(v["ms"] for v in prof.values())
            traffic, traffic_src = None, None
            tf = ROOT / "profiles" / "traffic.json"      # ncu --set full captures (dram__bytes_read.sum + dram__bytes_write.sum per launch)
            tj = json.loads(tf.read_text()) if tf.exists() else {}
            shape_key = f"{args.config}:{B}"
            tk = tj.get(shape_key, {})
            if "gemm" in tk:
                traffic, traffic_src = tk["gemm"]["dram_bytes_per_launch"], tk["gemm"]["source"]
            a_rate = acc_rate
            udt, upd = prof["udt"], prof["update"]
            udt_flops = (10.0 / 3.0) * N ** 3 * B * nb                    # QR 4/3 + norms 2/3 + form-Q 4/3
            upd_flops = 2.0 * N * N * nb * (a_rate * N) * B               # 2 n^2 per accepted flip and flavor block
            upd_bytes = 2.0 * B * nb * N * N * 8                           # one read + one write of every G
            hbm_peak = None
            mp = ROOT / "MEASURED_PEAKS.json"
            if mp.exists():
                hbm_peak = json.loads(mp.read_text()).get("hbm_gbs")
            hbm_peak = hbm_peak or 6552.0
            def per_class(c, flops, extra=None):
                avg = c["ms"] / max(c["count"], 1)
                tfl = flops / (avg * 1e-3) / 1e12 if avg > 0 else None
                d = {"launches_or_calls": c["count"], "avg_ms": avg, "ms_per_step": c["ms"] / K,
                     "share_of_step": c["ms"] / total_ms if total_ms else None,
                     "achieved_tflops": tfl, "frac_of_fp64_peak": (tfl / p64) if (tfl and p64) else None,
                     "algorithmic_flops": flops}
                if extra:
                    d.update(extra)
                return d
            upd_avg = upd["ms"] / max(upd["count"], 1)
            classes = {
                "gemm": per_class(gm, flops_per_launch, {"dram_bytes_per_launch": traffic, "dram_source": traffic_src}),
                "udt": per_class(udt, udt_flops, {"dram_bytes_per_call": tk.get("udt", {}).get("dram_bytes_per_launch"),
                                                  "dram_source": tk.get("udt", {}).get("source")}),
                "update": per_class(upd, upd_flops, {
                    "algorithmic_bytes": upd_bytes,
                    "achieved_gbs_algorithmic": upd_bytes / (upd_avg * 1e-3) / 1e9 if upd_avg > 0 else None,
                    "frac_of_hbm_peak_algorithmic": upd_bytes / (upd_avg * 1e-3) / 1e9 / hbm_peak if upd_avg > 0 else None,
                    "dram_bytes_per_launch": tk.get("update", {}).get("dram_bytes_per_launch"),
                    "dram_source": tk.get("update", {}).get("source")}),
            }
            roof = {"bound": "tensor", "kernel": "gemm_kernel (FP64 DMMA batched GEMM)", "achieved": achieved,
                    "peak": p64, "unit": "TFLOP/s", "frac": achieved / p64 if p64 else None, "traffic": traffic,
                    "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, "
                                    f"{traffic_src or 'no capture of this launch shape committed'}; algorithmic "
                                    f"{3 * B * nb * N * N * 8 / 1e6:.0f} MB = two operands read + one result written)",
                    "peak_source": "measured in this run: cuBLAS DGEMM (torch.matmul f64 8192^3, best of 5); "
                                   "MEASURED_PEAKS.json has no FP64 entry",
                    "cublas_batched_same_shape": same_shape,   # torch.bmm f64 on (chains x blocks) n x n x n, for context
                    "launches": gm["count"], "avg_launch_ms": avg_ms,
                    "share_of_step": gm["ms"] / total_ms if total_ms else None,
                    "kernel_ms_per_step": {k: v["ms"] / K for k, v in prof.items()},
                    "classes": classes,
                    "whole_sweep_frac_of_fp64_peak": (flops_per_sweep(N, M, C, nb, acc_rate) * value / world / 1e12) / p64 if p64 else None}

    # ---------------- measurement pass (cfg5's equal-time + unequal-time clause) ----------------
    meas = None
    if args.measure or args.config == "cfg5":
        ctx.set_lattice(np.array(lattice.bravais_srctrg2dir(), dtype=np.int32), len(lattice.unitcell.sites), T, U)
        ctx.measure_equal_time(); ctx.measure_time_integral(SAFE_MULT, DELTA_TAU)          # warm-up
        l1 = ctx.kernel_launches()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        barrier()
        ev[0].record(stream)
        ctx.measure_equal_time()
        ev[1].record(stream)
        ctx.measure_time_integral(SAFE_MULT, DELTA_TAU)
        ev[2].record(stream)
        barrier()
        t_et, t_ti = max_over_ranks(ev[0].elapsed_time(ev[1])), max_over_ranks(ev[1].elapsed_time(ev[2]))
        rate = 10                                                                         # DQMC.jl:37 measure_rate
        meas = {"equal_time_ms": t_et, "time_integral_ms": t_ti, "chains_per_gpu": B, "gpu_launches": int(ctx.kernel_launches() - l1),
                "observables": "occupation, kinetic/interaction/total energy, CDC, SDC x/y/z (equal time); CDS, SDS x/y/z "
                               "(TimeIntegral over the CombinedGreensIterator, recalculate = 2 safe_mult)",
                "sweeps_per_s_with_measurements": world * B * rate / ((rate * ms / K + t_et + t_ti) * 1e-3),
                "measure_rate": rate}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            meas["cpu"] = run_cpu_measure(args.config)

    line = None
    if rank == 0:
        line = {"metric": "DQMC sweeps/sec (all chains)", "value": value, "unit": "sweeps/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(args.config, B, world),
                "acceptance": acc_rate, "gpu_launches": int(launches), "clocks": clk,
                "e2e": {"value": e2e_value, "unit": "sweeps/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / K},
                "roofline": roof,
                "flops_per_sweep_per_chain": flops_per_sweep(N, M, C, nb, acc_rate)}
        if meas is not None:
            line["measurement"] = meas
    # ---------------- the other BASELINE configurations, device-resident, same timing rules (short) ----------
    others = None
    if args.config == "cfg4" and not args.no_other_configs:
        others = {}
        for cfg2, nsw in (("cfg1", 100), ("cfg2", 20), ("cfg3", 5), ("cfg5", 3)):
            others[cfg2] = quick_config(pkg, torch, cfg2, dev, rank, world, nsw, max_over_ranks)
    if rank == 0:
        if others is not None:
            if world == 1 and not args.no_cpu_baseline:      # rows "configs 1-2 faster than the reference": CPU port beside them
                for cfg2 in ("cfg1", "cfg2"):
                    r1 = run_cpu(cfg2, 20 if cfg2 == "cfg2" else 400, 1, max_chains=1 if cfg2 == "cfg1" else None)
                    others[cfg2]["cpu_port"] = {"value": r1["value"], "cores": r1["cores"], "chains": r1["chains"],
                                                "sweeps_per_chain": r1["sweeps_per_chain"], "seconds": r1["seconds"]}
            line["other_configs"] = others
        if world == 1 and not args.no_cpu_baseline:
            r = run_cpu(args.config, 1, 0)
            if r["seconds"] < 4.0:      # bounded sample of about 10 s: size it from the first sweep
                r = run_cpu(args.config, int(min(2000, max(2, round(10.0 / max(r["seconds"], 1e-3))))), 0)
            line["cpu_baseline"] = {"value": r["value"], "unit": "sweeps/s", "cores": r["cores"], "kind": "port",
                                    "sample": f"{r['chains']} chains (one per host core) x {r['sweeps_per_chain']} sweep(s) "
                                              f"each of the same workload, {r['seconds']:.1f} s; oracle/dqmc_ref.c "
                                              "(C port, Julia absent)"}
            line["parity"] = parity_leg(pkg, args.config, dev)
        emit(line)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ctx_bytes(N, M, C, nb, B):
    ld = (N + 1) & ~1
    mat = B * nb * ld * N * 8
    return mat * (2 * (C + 1) + 11) + B * M * N


def measure_cublas_batched(torch, n, batch):
    """cuBLAS batched DGEMM on this launch shape (context only: what the vendor library gets on batch x n^3)."""
    a = torch.randn(batch, n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(batch, n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    torch.bmm(a, b, out=c)
    best = 1e30
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); torch.bmm(a, b, out=c); e.record(); e.synchronize()
        best = min(best, s.elapsed_time(e))
    del a, b, c
    return 2.0 * n ** 3 * batch / (best * 1e-3) / 1e12


def measure_fp64_peak(torch, n=8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = 1e30
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); torch.matmul(a, b); e.record(); e.synchronize()
        best = min(best, s.elapsed_time(e))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


if __name__ == "__main__":
    main()
