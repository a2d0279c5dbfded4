"""Pins the unequal-time part of the C oracle (oracle/dqmc_ref_ut.inc.c) against the reference's own
tests for this path: test/DQMC/unequal_time_stack.jl and the U = 0 analytic G(k, l) used by
test/ED/ED_tests.jl.  numpy / mpmath are the independent arbiters.  No GPU.
"""
import mpmath as mp
import numpy as np
import pytest

from oracle import model as M
from oracle import ref as R


def rand_conf(seed, N, Ms):
    g = np.random.default_rng(seed)
    return np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(N, Ms)))


def chain6(beta, safe_mult, U=1.0, seed=11):
    """HubbardModel(6, 1); DQMC(m; beta, safe_mult) of test/DQMC/unequal_time_stack.jl:17-22."""
    T = M.hopping_matrix("chain", (6,))
    c = R.RefChain(T, U=U, beta=beta, safe_mult=safe_mult)
    c.set_conf(rand_conf(seed, c.N, c.M))
    return c


# ------------------------------------------------------------ unequal_time_stack.jl:1-15
def test_find_range_with_value():
    T = M.hopping_matrix("square", (2, 2))
    c = R.RefChain(T, U=1.0, beta=2.3, safe_mult=10)
    assert c.find_range_with_value(-81273) == 0
    assert c.find_range_with_value(0) == 0
    for i in range(1, 24):
        idx = c.find_range_with_value(i)
        assert c.ranges[idx - 1][0] <= i <= c.ranges[idx - 1][1]
    assert c.find_range_with_value(24) == c.C + 1
    assert c.find_range_with_value(1239874) == c.C + 1


# ------------------------------------------------------------ unequal_time_stack.jl:24-69
def test_lazy_builds_match_equal_time_stack():
    c = chain6(15.0, 5)
    c.build_stack()                                   # forward build: u_stack[i] = B(range i) ... B_1
    for upto in (4, 6):
        c.ut_lazy_build(forward_upto=upto)
        for i in range(upto):
            for w, uw in (("u_stack", "forward_u"), ("d_stack", "forward_d"), ("t_stack", "forward_t")):
                assert np.allclose(c.array(w, i), c.ut_array(uw, i), rtol=1e-12, atol=1e-14)
        for i in range(upto, c.C + 1):                # untouched slots are still zero
            assert np.all(c.ut_array("forward_u", i) == 0)
    while c.state[2] == -1:
        c.propagate()
    for downto in (8, 6):
        c.ut_lazy_build(backward_downto=downto)
        for i in range(c.C, downto - 2, -1):
            for w, uw in (("u_stack", "backward_u"), ("d_stack", "backward_d"), ("t_stack", "backward_t")):
                assert np.allclose(c.array(w, i), c.ut_array(uw, i), rtol=1e-12, atol=1e-14)
        for i in range(downto - 2, -1, -1):
            assert np.all(c.ut_array("backward_u", i) == 0)


# ------------------------------------------------------------ unequal_time_stack.jl:71-91
def test_build_stack_forward_backward():
    c = chain6(15.0, 5)
    c.build_stack()
    c.ut_build_stack()
    for i in range(c.C + 1):
        for w, uw in (("u_stack", "forward_u"), ("d_stack", "forward_d"), ("t_stack", "forward_t")):
            assert np.allclose(c.array(w, i), c.ut_array(uw, i), rtol=1e-12, atol=1e-14)
    while c.state[2] == -1:
        c.propagate()
    for i in range(1, c.C + 1):
        for w, uw in (("u_stack", "backward_u"), ("d_stack", "backward_d"), ("t_stack", "backward_t")):
            assert np.allclose(c.array(w, i), c.ut_array(uw, i), rtol=1e-12, atol=1e-14)


# ------------------------------------------------------------ unequal_time_stack.jl:97-113
@pytest.mark.parametrize("U", [1.0, -1.0])
def test_equal_time_from_unequal_time_stack(U):
    c = chain6(15.0, 5, U=U)
    c.init()
    for s in range(0, c.M + 1):
        G1 = c.calculate_greens_at(s)
        G2 = c.ut_calculate_greens(s, s)
        assert np.abs(G1 - G2).max() < 1e-13           # reference: 1e-14 on its own data
    for s in range(0, c.M):
        G1 = c.ut_greens(s, 0)
        G2 = c.ut_greens(s, c.M)
        assert np.allclose(G1, -G2, atol=1e-13, rtol=1e-10)


# ------------------------------------------------------------ unequal_time_stack.jl:116-172
@pytest.mark.parametrize("U", [1.0, -1.0])
def test_combined_greens_iterator_against_greens_kl(U):
    c = chain6(15.0, 5, U=U)
    c.init()                                          # current_slice 1, direction +1: greens == G(0, 0)
    Gk0 = [c.ut_greens(k, 0) for k in range(c.M + 1)]
    G0k = [c.ut_greens(0, k) for k in range(c.M + 1)]
    Gkk = []
    eTh, eThi = c.eThalf, c.eThalfinv
    for k in range(c.M + 1):
        g = c.calculate_greens_at(k)
        Gkk.append(np.stack([eThi @ g[:, :, b] @ eTh for b in range(c.nb)], axis=2))
    c.init()                                          # restore mc.stack.greens
    # high precision: recalculate = safe_mult
    seen = 0
    for (l, g0l, gl0, gll) in c.combined_greens_iterator(recalculate=c.safe_mult, start=0, stop=c.M):
        assert np.abs(gl0 - Gk0[l]).max() < 2e-14
        assert np.abs(g0l - G0k[l]).max() < 2e-14
        assert np.abs(gll - Gkk[l]).max() < 2e-14
        seen += 1
    assert seen == c.M + 1
    # low precision: recalculate = 4 safe_mult, default start/stop
    c.init()
    for (l, g0l, gl0, gll) in c.combined_greens_iterator(recalculate=4 * c.safe_mult):
        assert np.abs(gl0 - Gk0[l]).max() < 1e-10
        assert np.abs(g0l - G0k[l]).max() < 1e-10
        assert np.abs(gll - Gkk[l]).max() < 1e-10


@pytest.mark.parametrize("start", [1, 7])
def test_combined_greens_iterator_start_variants(start):
    """iterate(it) branches for start == 1 (:242-244) and start > 1 (:246-292)."""
    c = chain6(4.0, 5)
    c.init()
    Gk0 = [c.ut_greens(k, 0) for k in range(c.M + 1)]
    G0k = [c.ut_greens(0, k) for k in range(c.M + 1)]
    c.init()
    ls = []
    for (l, g0l, gl0, gll) in c.combined_greens_iterator(recalculate=10, start=start, stop=c.M - 3):
        assert np.abs(gl0 - Gk0[l]).max() < 1e-11
        assert np.abs(g0l - G0k[l]).max() < 1e-11
        ls.append(l)
    assert ls == list(range(start, c.M - 2))


# ------------------------------------------------------------ unequal_time_stack.jl:176-304 (BigFloat)
def _mp_B(c, conf, l, inverse=False):
    eV = mp.diag([mp.e ** (mp.mpf(c.alpha) * int(conf[i, l - 1]) * (-1 if inverse else 1)) for i in range(c.N)])
    if inverse:
        return eV * mp.matrix(c.eT2inv.tolist())
    return mp.matrix(c.eT2.tolist()) * eV


def test_time_displaced_greens_high_precision():
    mp.mp.prec = 128
    c = chain6(5.0, 10, seed=5)
    conf = c.get_conf()
    k, l = 37, 14
    inv_B = mp.eye(c.N)
    for s in range(k, l, -1):
        inv_B = _mp_B(c, conf, s, inverse=True) * inv_B
    fwd = mp.eye(c.N)
    for s in range(1, l + 1):
        fwd = _mp_B(c, conf, s) * fwd
    bwd = mp.eye(c.N)
    for s in range(c.M, k, -1):
        bwd = _mp_B(c, conf, s).T * bwd
    ref = mp.inverse(inv_B + fwd * bwd.T)
    ref = np.array(ref.tolist(), dtype=float)
    G = c.ut_calculate_greens(k, l)[:, :, 0]
    assert np.abs(G - ref).max() <= 1e-11 * np.abs(ref).max() + 1e-15
    # stack values (unequal_time_stack.jl:189-247): U D T of every slot vs the BigFloat chain product
    c.ut_build_stack()
    P = mp.eye(c.N)
    for idx, (a, b) in enumerate(c.ranges, start=1):
        for s in range(a, b + 1):
            P = _mp_B(c, conf, s) * P
        got = c.ut_array("forward_u", idx)[:, :, 0] @ np.diag(c.ut_array("forward_d", idx)[:, 0]) @ c.ut_array("forward_t", idx)[:, :, 0]
        want = np.array(P.tolist(), dtype=float)
        assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
    for idx, (a, b) in enumerate(c.ranges, start=1):
        Q = mp.eye(c.N)
        for s in range(b, a - 1, -1):
            Q = _mp_B(c, conf, s, inverse=True) * Q
        got = c.ut_array("inv_u", idx - 1)[:, :, 0] @ np.diag(c.ut_array("inv_d", idx - 1)[:, 0]) @ c.ut_array("inv_t", idx - 1)[:, :, 0]
        want = np.array(Q.tolist(), dtype=float)
        assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
    mp.mp.prec = 53


def test_full2_high_precision():
    """calculate_greens_full2! (k < l): G(k, l) = -[B_l ... B_{k+1} + (B_k ... B_1 B_M ... B_{l+1})^-1]^-1."""
    mp.mp.prec = 128
    c = chain6(5.0, 10, seed=6)
    conf = c.get_conf()
    k, l = 12, 41
    mid = mp.eye(c.N)
    for s in range(k + 1, l + 1):
        mid = _mp_B(c, conf, s) * mid
    outer = mp.eye(c.N)
    for s in list(range(l + 1, c.M + 1)) + list(range(1, k + 1)):
        outer = _mp_B(c, conf, s) * outer
    ref = -mp.inverse(mid + mp.inverse(outer))
    ref = np.array(ref.tolist(), dtype=float)
    G = c.ut_calculate_greens(k, l)[:, :, 0]
    assert np.abs(G - ref).max() <= 1e-11 * np.abs(ref).max() + 1e-15
    mp.mp.prec = 53


# ------------------------------------------------------------ test/ED/ED_tests.jl:102-183, 284-299 (U = 0)
@pytest.mark.parametrize("lat,Ls", [("honeycomb", (2, 1)), ("square", (4, 4))])
def test_U0_analytic_time_displaced(lat, Ls):
    T = M.hopping_matrix(lat, Ls)
    c = R.RefChain(T, U=0.0, beta=2.0, safe_mult=5)
    c.set_conf(rand_conf(3, c.N, c.M))
    c.init()
    w, V = np.linalg.eigh(T)
    f = 1.0 / (1.0 + np.exp(-c.beta * w))             # <c c^dagger>
    for (k, l) in [(0, 0), (7, 0), (c.M, 0), (13, 4), (c.M, c.M), (9, 9)]:
        tau = (k - l) * c.delta_tau
        want = (V * (np.exp(-tau * w) * f)) @ V.T
        assert np.abs(c.ut_greens(k, l)[:, :, 0] - want).max() < 1e-12
    for (k, l) in [(0, 5), (3, 17), (0, c.M)]:
        tau = (k - l) * c.delta_tau                   # negative
        want = -(V * (np.exp(-tau * w) * (1.0 - f))) @ V.T
        assert np.abs(c.ut_greens(k, l)[:, :, 0] - want).max() < 1e-12
