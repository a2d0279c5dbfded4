"""End-to-end physics check of the GPU path against exact diagonalisation at finite U -- the reference's own
statistical test (test/ED/ED_tests.jl:402-560: 2 x 2 Hubbard, beta = 1, dtau = 0.1, safe_mult = 5; tolerance
atol = 3.05 dtau^2, rtol = 2 dtau^2 for the Trotter error) run through the user-level API: DQMC(model; ...),
device measurements, run.  Many chains replace the reference's 5k + 5k sweeps of one chain.
"""
import numpy as np
import pytest

from oracle import measure as OMS
from oracle import model as M
from oracle.ed import HubbardED

pytestmark = pytest.mark.gpu

ATOL, RTOL = 3.05 * 0.1 ** 2, 2 * 0.1 ** 2


def check(a, b, extra=0.0):
    return np.all(np.abs(np.asarray(a) - np.asarray(b)) <= ATOL + RTOL * np.abs(b) + extra)


@pytest.mark.parametrize("U,mu", [(-1.0, 0.0), (1.0, 1.0)])
def test_gpu_dqmc_against_exact_diagonalisation(b200, U, mu):
    model = b200.HubbardModel(b200.SquareLattice(2), U=U, mu=mu)
    mc = b200.DQMC(model, beta=1.0, delta_tau=0.1, safe_mult=5, thermalization=100, sweeps=400, measure_rate=2,
                   seed=17, n_chains=256)
    for key, m in (("occ", b200.occupation), ("K", b200.kinetic_energy), ("V", b200.interaction_energy),
                   ("E", b200.total_energy), ("cdc", b200.charge_density_correlation),
                   ("cds", b200.charge_density_susceptibility)):
        mc[key] = m(mc, model)
    mc["sdzc"] = b200.spin_density_correlation(mc, model, "z")
    mc["sdxc"] = b200.spin_density_correlation(mc, model, "x")
    mc["sdzs"] = b200.spin_density_susceptibility(mc, model, "z")
    assert b200.run(mc) == "SUCCESS"
    assert mc["E"].count == 200

    T = M.hopping_matrix("square", (2, 2), mu=mu)
    assert np.array_equal(T, mc.hopping_matrix)
    s2d = OMS.bravais_srctrg2dir((2, 2))
    ed = HubbardED(T, 1.0, U)
    dens = [ed.density(i) for i in range(4)]
    mz = [ed.spin_ops(i)[2] for i in range(4)]
    mx = [ed.spin_ops(i)[0] for i in range(4)]
    err = {k: 4 * mc[k].std_error() * np.sqrt(10) for k in mc.measurements}       # 4 sigma, tau_int <= 5 measurements
    nflv = mc.ctx.nb
    occ = mc["occ"].mean()
    for f in range(nflv):
        assert check(occ[4 * f:4 * f + 4], [ed.expect(ed.n(i, f)) for i in range(4)], np.max(err["occ"]))
    assert check(mc["K"].mean(), ed.expect(ed.kinetic()), err["K"])
    assert check(mc["V"].mean(), ed.expect(ed.interaction()), err["V"])
    assert check(mc["E"].mean(), ed.expect(ed.H), err["E"])
    assert check(mc["cdc"].mean()[:, 0, 0], ed.pair_by_distance(dens, s2d), np.max(err["cdc"]))
    assert check(mc["sdzc"].mean()[:, 0, 0], ed.pair_by_distance(mz, s2d), np.max(err["sdzc"]))
    assert check(mc["sdxc"].mean()[:, 0, 0], ed.pair_by_distance(mx, s2d), np.max(err["sdxc"]))
    # susceptibilities: the same trapezoid rule over the exact <O(tau) O(0)>
    M_, dt = mc.ctx.M, 0.1
    chi_c = np.zeros(4); chi_z = np.zeros(4)
    for l in range(M_ + 1):
        w = (0.5 if l in (0, M_) else 1.0) * dt
        chi_c += w * ed.pair_by_distance(dens, s2d, l * dt)
        chi_z += w * ed.pair_by_distance(mz, s2d, l * dt)
    assert check(mc["cds"].mean()[:, 0, 0], chi_c, np.max(err["cds"]))
    assert check(mc["sdzs"].mean()[:, 0, 0], chi_z, np.max(err["sdzs"]))
