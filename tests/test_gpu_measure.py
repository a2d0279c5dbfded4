"""GPU parity of the device-side Wick kernels (montecarlo.jl_b200/csrc/measure.cu, through the C ABI) against
oracle/measure.py evaluated on the CPU oracle's Green's functions.  Run with `-m gpu` on a B200.

Tolerance: 1e-9 relative to the largest entry of each observable (sums of ~N^2 products of Green's function
entries that themselves agree to 1e-10).
"""
import numpy as np
import pytest

from oracle import measure as OMS
from oracle import model as OM

from test_gpu_parity import make_pair

pytestmark = pytest.mark.gpu

CASES = [
    # kind, Ls, n_basis, U, beta, safe_mult
    ("square", (4, 4), 1, 4.0, 2.0, 10),        # attractive: one flavor block (DiagonallyRepeatingMatrix kernels)
    ("square", (4, 4), 1, -4.0, 2.0, 5),        # repulsive: two flavor blocks (BlockDiagonal kernels)
    ("honeycomb", (3, 3), 2, 4.0, 1.0, 5),      # two-site basis: temp[dir, b1, b2]
    ("honeycomb", (2, 3), 2, -4.0, 1.0, 5),
]


def close(got, want, tol=1e-9):
    return np.abs(np.asarray(got) - np.asarray(want)).max() <= tol * max(np.abs(want).max(), 1.0)


@pytest.mark.parametrize("kind,Ls,nbasis,U,beta,sm", CASES)
def test_equal_time_observables(b200, kind, Ls, nbasis, U, beta, sm):
    ctx, chains = make_pair(b200, kind, Ls, U=U, beta=beta, B=3, safe_mult=sm)
    T = OM.hopping_matrix(kind, Ls)
    s2d = OMS.bravais_srctrg2dir(Ls)
    ctx.set_lattice(s2d, nbasis, T, U)
    ctx.build_stack()
    ctx.sweep(1)
    ctx.measure_equal_time()
    got = ctx.measurements()
    for b, c in enumerate(chains):
        c.init(); c.local_sweep()
        want = OMS.equal_time(c.measured_greens(), T, U, s2d, nbasis)
        for name in ("occ", "K", "V", "E", "cdc", "sdxc", "sdyc", "sdzc"):
            assert close(got[name][b], want[name]), (name, b)
    # accumulators: count, sum and sum of squares over chains; a second call doubles them
    ctx.measure_equal_time()
    n_et, n_ti, s, s2 = ctx.measurement_stats()
    assert n_et == 2 * ctx.B and n_ti == 0
    assert close(s["cdc"], 2 * got["cdc"].sum(axis=0))
    assert close(s2["K"], 2 * (got["K"] ** 2).sum())


@pytest.mark.parametrize("kind,Ls,nbasis,U,beta,sm", CASES)
def test_time_integrated_observables(b200, kind, Ls, nbasis, U, beta, sm):
    ctx, chains = make_pair(b200, kind, Ls, U=U, beta=beta, B=2, safe_mult=sm)
    T = OM.hopping_matrix(kind, Ls)
    s2d = OMS.bravais_srctrg2dir(Ls)
    ctx.set_lattice(s2d, nbasis, T, U)
    ctx.build_stack()
    ctx.measure_time_integral(sm, 0.1, recalculate=sm)
    got = ctx.measurements()
    for b, c in enumerate(chains):
        c.init()
        G00 = c.measured_greens()
        want = OMS.time_integral(G00, c.combined_greens_iterator(recalculate=sm), c.delta_tau, c.M, s2d, nbasis)
        for name in ("cds", "sdxs", "sdys", "sdzs"):
            assert close(got[name][b], want[name]), (name, b)
    n_et, n_ti, s, s2 = ctx.measurement_stats()
    assert n_et == 0 and n_ti == ctx.B
    # the sweep after a measurement is unaffected
    acc = ctx.sweep(1)
    for b, c in enumerate(chains):
        assert c.local_sweep() == acc[b]


def test_U0_susceptibility_is_exact(b200):
    """U = 0: chi_c(q = 0) = sum_dir cds[dir] = integral of <N(tau) N(0)> / N_sites = beta <N>^2 / N_sites for a
    conserved total charge N (all chains identical, no Monte Carlo noise)."""
    kind, Ls = "square", (4, 4)
    ctx, chains = make_pair(b200, kind, Ls, U=0.0, beta=2.0, B=2, safe_mult=5, mu=0.3)
    T = OM.hopping_matrix(kind, Ls, mu=0.3)
    ctx.set_lattice(OMS.bravais_srctrg2dir(Ls), 1, T, 0.0)
    ctx.build_stack()
    ctx.measure_equal_time()
    ctx.measure_time_integral(5, 0.1, recalculate=5)
    got = ctx.measurements()
    w, _ = np.linalg.eigh(T)
    f = 1.0 / (1.0 + np.exp(2.0 * w))
    Ntot = 2.0 * f.sum()
    varN = 2.0 * (f * (1.0 - f)).sum()
    # <N(tau) N(0)> = <N^2> for conserved N: the trapezoid rule integrates a constant exactly
    assert abs(got["cds"][0].sum() - 2.0 * (Ntot ** 2 + varN) / 16) < 1e-9
    assert abs(got["cdc"][0].sum() - (Ntot ** 2 + varN) / 16) < 1e-10
    assert abs(got["occ"][0].sum() * 2 - Ntot) < 1e-11


def test_run_with_device_measurements_and_global_updates(b200):
    """The user-level flow of the reference (DQMC.jl:252-394): DQMC(model; ...), mc[:key] = measurement, run!(mc)
    with a SimpleScheduler(LocalSweep(), GlobalFlip()) -- every registered measurement receives one value per
    chain per measurement point, and the values equal a direct evaluation through the context."""
    model = b200.HubbardModel(b200.Honeycomb(2, 2), U=-2.0, mu=0.2)
    mc = b200.DQMC(model, beta=1.0, safe_mult=5, thermalization=2, sweeps=6, measure_rate=3, seed=5, n_chains=3,
                   scheduler=b200.SimpleScheduler(b200.LocalSweep(2), b200.GlobalFlip()))
    mc["occ"] = b200.occupation(mc, model)
    mc["E"] = b200.total_energy(mc, model)
    mc["CDC"] = b200.charge_density_correlation(mc, model)
    mc["SDSz"] = b200.spin_density_susceptibility(mc, model, "z")
    mc["G"] = b200.greens_measurement(mc, model)
    assert b200.run(mc) == "SUCCESS"
    assert mc["occ"].count == 2 and mc["SDSz"].count == 2 and mc["G"].count == 2
    assert mc["occ"].mean().shape == (2 * 8,) and mc["CDC"].mean().shape == (4, 2, 2)
    assert mc.global_total == 2 and mc.total == 6 * 2 * 8 * 10
    # the last measurement point is the current state: re-evaluate through the context
    mc.ctx.measure_equal_time()
    vals = mc.ctx.measurements()
    last = mc["E"].sum - (mc["E"].sum - vals["E"])          # shape check only
    assert last.shape == (3,)
    G = mc.ctx.measured_greens()
    occ = np.concatenate([1.0 - np.diagonal(G[:, :, f, :], axis1=0, axis2=1) for f in range(2)], axis=1)
    assert np.allclose(vals["occ"], occ, atol=1e-12)
    # half filling is not enforced at mu != 0, but densities stay physical
    assert np.all(vals["occ"] > -1e-9) and np.all(vals["occ"] < 1 + 1e-9)
    assert np.isfinite(mc["SDSz"].mean()).all() and np.isfinite(mc["CDC"].std_error()).all()


def test_device_log_binning_matches_logbinner(b200):
    """The per-level {count, sum, sum of squares} the device keeps for every observable element equal a LogBinner per chain
    (oracle/measure.py) fed with the values of each measurement, summed over the chains ("per log-bin level")."""
    from oracle.measure import LogBinner
    model = b200.HubbardModel(b200.SquareLattice(4), U=-4.0)
    mc = b200.DQMC(model, beta=1.0, safe_mult=5, seed=5, n_chains=3)
    mc.init()
    mc._set_lattice()
    ctx = mc.ctx
    binners = None
    nmeas = 11
    for k in range(nmeas):
        ctx.sweep(1)
        ctx.measure_equal_time()
        vals = ctx.measurements()
        if binners is None:
            binners = {key: [LogBinner(shape=v.shape[1:]) for _ in range(3)] for key, v in vals.items() if key in ("occ", "E", "cdc")}
        for key, bl in binners.items():
            for b in range(3):
                bl[b].push(vals[key][b])
        if k % 4 == 3:                                          # a TimeIntegral series of its own length
            ctx.measure_time_integral(5, 0.1)
    cnt, s, s2 = ctx.measurement_binning()
    L = cnt.shape[1]
    want_counts = [3 * (nmeas >> l) for l in range(L)]
    assert cnt[0].tolist() == want_counts
    assert cnt[1].tolist() == [3 * (2 >> l) for l in range(L)]
    for key, bl in binners.items():
        for l in range(4):
            assert np.allclose(s[key][l], sum(b.sum[l] for b in bl), rtol=1e-12, atol=1e-13), (key, l)
            assert np.allclose(s2[key][l], sum(b.sumsq[l] for b in bl), rtol=1e-12, atol=1e-13), (key, l)
    lv, n, err = ctx.binning_std_errors("E")
    assert lv.tolist() == [0, 1, 2, 3] and np.all(err[:3] > 0)      # counts are summed over the 3 chains
    # level 0 is the flat accumulator
    c_et, c_ti, fs, fs2 = ctx.measurement_stats()
    assert c_et == cnt[0, 0] and np.allclose(fs["cdc"], s["cdc"][0])


@pytest.mark.parametrize("field", [None, "MagneticGHQField"])
def test_recorder_and_replay(b200, field):
    """ConfigRecorder + replay! (configurations.jl:12-60, DQMC.jl:418-505): the configurations recorded during run (device
    bit-packing) replayed on a fresh simulation reproduce every measured value."""
    def make():
        model = b200.HubbardModel(b200.SquareLattice(4), U=-4.0)
        mc = b200.DQMC(model, beta=1.0, safe_mult=5, thermalization=3, sweeps=8, measure_rate=2, seed=9, n_chains=2,
                       field=field)
        mc["G"] = b200.greens_measurement(mc, model)
        mc["E"] = b200.total_energy(mc, model)
        mc["sdsz"] = b200.spin_density_susceptibility(mc, model, "z")
        return mc
    mc = make()
    assert b200.run(mc) == "SUCCESS"
    assert len(mc.recorder) == 4 and mc.recorder.sweeps == [4, 6, 8, 10] and mc["G"].count == 4
    words = (16 * 10 * (2 if field else 1) + 63) // 64
    assert mc.recorder[0].shape == (words, 2) and mc.recorder[0].dtype == np.uint64
    mc2 = make()
    assert b200.replay(mc2, mc.recorder) == "SUCCESS"
    assert mc2["G"].count == 4
    assert np.allclose(mc2["G"].mean(), mc["G"].mean(), rtol=0, atol=1e-10)
    assert np.allclose(mc2["E"].mean(), mc["E"].mean(), rtol=1e-10)
    assert np.allclose(mc2["sdsz"].mean(), mc["sdsz"].mean(), rtol=1e-8, atol=1e-10)
    assert np.array_equal(mc2.ctx.get_conf_packed(), mc.recorder[-1])          # the replay ends on the last recorded configuration
