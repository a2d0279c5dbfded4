"""Statistical pin of the C oracle against exact diagonalisation at finite U: the reference's own end-to-end
check (test/ED/ED_tests.jl:402-560: 2 x 2 Hubbard, beta = 1, dtau = 0.1, safe_mult = 5, atol = 3.05 dtau^2,
rtol = 2 dtau^2 for the Trotter error).  Sweeps + measured Green's function + Wick kernels.  No GPU.
"""
import numpy as np
import pytest

from oracle import measure as OMS
from oracle import model as M
from oracle import ref as R
from oracle.ed import HubbardED

ATOL, RTOL = 3.05 * 0.1 ** 2, 2 * 0.1 ** 2


def check(a, b, extra=0.0):
    return np.all(np.abs(np.asarray(a) - np.asarray(b)) <= ATOL + RTOL * np.abs(b) + extra)


def ed_reference(T, beta, U, s2d):
    ed = HubbardED(T, beta, U)
    N = T.shape[0]
    dens = [ed.density(i) for i in range(N)]
    mz = [ed.spin_ops(i)[2] for i in range(N)]
    mx = [ed.spin_ops(i)[0] for i in range(N)]
    return {"occ": np.array([ed.expect(ed.n(i, 0)) for i in range(N)]),
            "K": ed.expect(ed.kinetic()), "V": ed.expect(ed.interaction()),
            "cdc": ed.pair_by_distance(dens, s2d), "sdzc": ed.pair_by_distance(mz, s2d),
            "sdxc": ed.pair_by_distance(mx, s2d)}


@pytest.mark.parametrize("U,mu", [(-1.0, 0.0), (1.0, 1.0)])
def test_oracle_sweeps_against_exact_diagonalisation(U, mu):
    T = M.hopping_matrix("square", (2, 2), mu=mu)
    s2d = OMS.bravais_srctrg2dir((2, 2))
    want = ed_reference(T, 1.0, U, s2d)
    g = np.random.default_rng(7)
    nchains, therm, sweeps = 8, 200, 1500
    acc = {k: [] for k in ("occ", "K", "V", "cdc", "sdzc", "sdxc")}
    for b in range(nchains):
        c = R.RefChain(T, U=U, beta=1.0, safe_mult=5, seed=99, chain_id=b,
                       conf=np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(4, 10))))
        c.init()
        for s in range(therm + sweeps):
            c.local_sweep()
            if s >= therm:
                et = OMS.equal_time(c.measured_greens(), T, U, s2d, 1)
                for k in acc:
                    acc[k].append(et[k])
    mean = {k: np.mean(np.array(v), axis=0) for k, v in acc.items()}
    err = {k: 4 * np.std(np.array(v), axis=0) / np.sqrt(len(v) / 20) for k, v in acc.items()}   # 4 sigma, tau_int <= 10
    occ = mean["occ"][:4]
    assert check(occ, want["occ"], err["occ"][:4].max())
    assert check(mean["K"], want["K"], err["K"])
    assert check(mean["V"], want["V"], err["V"])
    assert check(mean["cdc"][:, 0, 0], want["cdc"], err["cdc"].max())
    assert check(mean["sdzc"][:, 0, 0], want["sdzc"], err["sdzc"].max())
    assert check(mean["sdxc"][:, 0, 0], want["sdxc"], err["sdxc"].max())
