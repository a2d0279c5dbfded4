"""The extended-precision arbiter (oracle/truth_ld.c) against brute force, the analytic U = 0 result and the
double-precision oracle.  CPU only."""
import numpy as np
import pytest

from oracle import model as OM
from oracle import ref as OR
from oracle import truth as TR


def rand_conf(seed, N, M):
    return np.asfortranarray(np.random.default_rng(seed).choice(np.array([-1, 1], dtype=np.int8), size=(N, M)))


@pytest.mark.parametrize("U", [4.0, -4.0])
def test_truth_equals_brute_force_at_short_beta(U):
    """beta = 0.5: the chain product is harmless in double, so inv(1 + B_M ... B_1) is exact to ~1e-14."""
    T = OM.hopping_matrix("square", (4, 4))
    c = OR.RefChain(T, U=U, beta=0.5, conf=rand_conf(1, 16, 5))
    G = TR.greens_truth_chain(c)
    for b in range(c.nb):
        P = np.eye(16)
        for l in range(c.M):
            P = c.eT2 @ (TR.slice_diagonals(c.get_conf(), c.alpha, c.kind, b)[l][:, None] * P)
        assert np.abs(G[:, :, b] - np.linalg.inv(np.eye(16) + P)).max() < 1e-13


def test_truth_U0_analytic():
    """test/flavortests_DQMC.jl:355-386: U = 0 => G = V diag(1 / (1 + e^{-beta eps})) V^T."""
    T = OM.hopping_matrix("square", (8, 8), mu=0.3)
    c = OR.RefChain(T, U=0.0, beta=10.0, conf=rand_conf(2, 64, 100))
    G = TR.greens_truth_chain(c)
    assert np.abs(G[:, :, 0] - OM.analytic_greens(T, 10.0)).max() < 2e-13   # eigh/expm of the input is double


@pytest.mark.parametrize("U,beta", [(4.0, 10.0), (-4.0, 8.0)])
def test_oracle_within_1e11_of_truth_at_long_beta(U, beta):
    """cfg 2 family (8x8): the double-precision oracle's stack G vs the arbiter, at the init point and for any slice;
    the arbiter does not depend on its own stabilisation interval."""
    T = OM.hopping_matrix("square", (8, 8))
    M = OM.n_slices(beta)
    c = OR.RefChain(T, U=U, beta=beta, conf=rand_conf(3, 64, M))
    c.init()
    G5, G10 = TR.greens_truth_chain(c, chunk=5), TR.greens_truth_chain(c, chunk=10)
    assert np.abs(G5 - G10).max() < 1e-14
    assert np.abs(c.greens - G10).max() / np.abs(G10).max() < 1e-11
    k = 37
    Gk = TR.greens_truth_chain(c, slice0=k)
    assert np.abs(c.calculate_greens_at(k) - Gk).max() / np.abs(Gk).max() < 1e-11
