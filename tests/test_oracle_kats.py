"""Pins the C oracle (oracle/dqmc_ref.c) against the reference's known-answer tests.

Each test restates one test of /root/reference/test (cited), with numpy/scipy as the
independent arbiter.  No GPU.  These are what make the oracle a trustworthy checker.
"""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import model as M
from oracle import ref as R


def rng(seed=0):
    return np.random.default_rng(seed)


def rand_conf(g, N, Ms):
    return np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(N, Ms)))


# ------------------------------------------------------------ test/linalg.jl:16-93
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_vmul_variants(ta, tb):
    g = rng(1)
    A, B = g.random((8, 8)), g.random((8, 8))
    ref = (A.T if ta else A) @ (B.T if tb else B)
    assert np.allclose(R.vmul(A, B, ta, tb), ref, atol=100 * np.finfo(float).eps, rtol=0)


@pytest.mark.parametrize("n", [1, 3, 7, 15, 16, 17, 23, 36, 50, 64, 100])
def test_vmul_register_tiles_and_tails(n):
    """The oracle's products are register-tiled (16 x 4 row-vector tiles, 4 x 4 dot-product tiles, 8-wide k vectors): every
    tail combination must still be the plain product."""
    g = rng(100 + n)
    A, B = g.standard_normal((n, n)), g.standard_normal((n, n))
    for ta in (0, 1):
        for tb in (0, 1):
            ref = (A.T if ta else A) @ (B.T if tb else B)
            assert np.allclose(R.vmul(A, B, ta, tb), ref, atol=1e-13 * max(1.0, np.abs(ref).max()), rtol=0), (n, ta, tb)


# ------------------------------------------------------------ test/linalg.jl:97-131
@pytest.mark.parametrize("kind", ["random", "rank1"])
def test_udt_identities_and_rdivp(kind):
    g = rng(2)
    X = g.random((8, 8)) if kind == "random" else np.kron(g.random(8)[:, None], g.random(8)[None, :])
    U, D, T, piv = R.udt_pivot(X, apply_pivot=True)
    assert np.allclose(U @ np.diag(D) @ T, X)
    assert np.allclose(U.T @ U, np.eye(8), atol=1e-13)

    U, D, T, piv = R.udt_pivot(X, apply_pivot=False)
    P = np.zeros((8, 8))
    for i, j in enumerate(piv):
        P[i, j] = 1.0
    assert np.allclose(U @ np.diag(D) @ np.triu(T) @ P, X)

    if kind == "random":  # test/linalg.jl:126-131 (rank-1 T is singular -> skip like a sane person)
        u = R.rdivp(U, T, piv)
        assert np.allclose(u, U @ P.T @ np.linalg.inv(np.triu(T)))


def test_udt_matches_lapack_pivoted_qr():
    """Column pivoting by largest remaining norm is LAPACK dgeqp3's rule as well
    (test/linalg/old_linalg.jl:16-24): same D and same pivots for a generic matrix."""
    g = rng(3)
    X = g.random((12, 12)) * np.exp(g.normal(size=12) * 3)[None, :]
    U, D, T, piv = R.udt_pivot(X, apply_pivot=True)
    Q, Rm, p = sla.qr(X, pivoting=True)
    assert np.array_equal(piv, p)
    assert np.allclose(D, np.abs(np.diag(Rm)), rtol=1e-12)


def test_udt_zero_column_gives_unit_D():
    """UDT.jl:293-301 (issue #169): exact zeros on diag(R) are replaced by 1."""
    X = np.zeros((4, 4)); X[:, 0] = [1.0, 2.0, 3.0, 4.0]
    U, D, T, piv = R.udt_pivot(X)
    assert np.all(D[1:] == 1.0)
    assert np.allclose(U @ np.diag(D) @ T, X)


# ------------------------------------------------------------ test/slice_matrices.jl:13-42
def test_slice_matrix_products():
    g = rng(4)
    T = M.hopping_matrix("chain", (8,))
    c = R.RefChain(T, U=1.0, beta=5.0, conf=rand_conf(g, 8, 50))
    eT = c.eThalf
    for sl in (7, 33):
        eV = np.diag(np.exp(c.alpha * c.get_conf()[:, sl - 1].astype(float)))
        A = eT @ eT @ eV
        X = g.random((8, 8))
        as3 = lambda Y: Y[:, :, None]
        assert np.allclose(c.multiply_slice_matrix("left", sl, as3(X))[:, :, 0], A @ X)
        assert np.allclose(c.multiply_slice_matrix("right", sl, as3(X))[:, :, 0], X @ A)
        Ai = np.linalg.inv(A)
        assert np.allclose(c.multiply_slice_matrix("inv_left", sl, as3(X))[:, :, 0], Ai @ X)
        assert np.allclose(c.multiply_slice_matrix("inv_right", sl, as3(X))[:, :, 0], X @ Ai)
        assert np.allclose(c.multiply_slice_matrix("daggered_left", sl, as3(X))[:, :, 0], A.T @ X)


# ------------------------------------------------------------ test/fields.jl:133-222
@pytest.mark.parametrize("U", [1.0, -1.0])
def test_rank1_update_formula(U):
    g = rng(5)
    N, i = 4, 2
    T = M.hopping_matrix("square", (2, 2))
    c = R.RefChain(T, U=U, beta=1.0, conf=rand_conf(g, N, 10))
    c.set_state(3, 1, 1)
    G = np.asfortranarray(g.random((N, N, c.nb)))
    c.set_greens(G)
    x = float(c.get_conf()[i, 2])
    dE = -2.0 * c.alpha * x
    p = c.propose_local(i, accept=True)
    Gn = c.greens
    if c.kind == 0:
        D = [np.exp(dE) - 1.0]
    else:
        D = [np.exp(dE) - 1.0, np.exp(-dE) - 1.0]
    Rs = [1.0 + D[b] * (1.0 - G[i, i, b]) for b in range(c.nb)]
    pref = np.exp(-dE) * Rs[0] ** 2 if c.kind == 0 else Rs[0] * Rs[1]
    assert np.isclose(p, pref, rtol=1e-14)
    for b in range(c.nb):
        IG = (np.eye(N) - G[:, :, b])[:, i]
        Q = G[:, :, b] - np.outer(IG, (D[b] / Rs[b]) * G[i, :, b])
        assert np.allclose(Gn[:, :, b], Q, rtol=1e-14, atol=0)
    assert c.get_conf()[i, 2] == -x


from oracle.bruteforce import decompose_udt, greens_brute, greens_lapack, log_weight, slice_B  # noqa: E402,F401


# ------------------------------------------------------------ test/flavortests_DQMC.jl:282-301
def test_stack_greens_vs_lapack_qr_and_wraps():
    g = rng(6)
    T = M.hopping_matrix("chain", (8,))
    conf = rand_conf(g, 8, 50)
    c = R.RefChain(T, U=1.0, beta=5.0, safe_mult=1, conf=conf)
    c.build_stack()
    c.propagate()
    cs = c.state[0]
    assert cs == 50
    Gl = greens_lapack(c, conf, cs)
    Gw = c.wrap_greens(Gl[:, :, None], cs + 1, -1)[:, :, 0]
    assert np.allclose(Gw, c.greens[:, :, 0])
    Gl = greens_lapack(c, conf, cs - 1)
    assert np.abs(Gl - c.greens[:, :, 0]).max() < 1e-12
    G = c.greens
    for k in range(10):
        G = c.wrap_greens(G, cs - k, -1)
    Gl = greens_lapack(c, conf, cs - 11)
    assert np.abs(Gl - G[:, :, 0]).max() < 1e-9


# ------------------------------------------------------------ test/flavortests_DQMC.jl:303-311
@pytest.mark.parametrize("U", [1.0, -1.0])
def test_calculate_greens_at_every_slice(U):
    g = rng(7)
    T = M.hopping_matrix("chain", (8,))
    conf = rand_conf(g, 8, 50)
    c = R.RefChain(T, U=U, beta=5.0, safe_mult=5, conf=conf)
    for k in g.permutation(51):
        G2 = c.calculate_greens_at(int(k))
        for b in range(c.nb):
            assert np.allclose(greens_lapack(c, conf, int(k), b), G2[:, :, b])


# ------------------------------------------------------------ test/flavortests_DQMC.jl:313-341
@pytest.mark.parametrize("U", [1.0, -2.0])
def test_forward_build_plus_propagates_equals_reverse_build(U):
    g = rng(8)
    T = M.hopping_matrix("chain", (8,))
    conf = rand_conf(g, 8, 50)
    c1 = R.RefChain(T, U=U, beta=5.0, conf=conf)
    c1.build_stack(); c1.propagate()
    for _ in range(c1.M):
        c1.propagate()
    c2 = R.RefChain(T, U=U, beta=5.0, conf=conf)
    c2.reverse_build_stack(); c2.propagate()
    assert c1.state == c2.state == (1, 1, 1)
    assert np.allclose(c1.greens, c2.greens)
    for name in ("u_stack", "d_stack", "t_stack"):
        for i in range(c1.C + 1):
            assert np.allclose(c1.array(name, i), c2.array(name, i)), (name, i)
    for name in ("Ul", "Ur", "Dl", "Dr", "Tl", "Tr"):
        assert np.allclose(c1.array(name), c2.array(name)), name


# ------------------------------------------------------------ test/flavortests_DQMC.jl:355-386
@pytest.mark.parametrize("L,mu", [(7, 0.0), (8, 1.0)])
@pytest.mark.parametrize("beta", [1.0, 10.0])
def test_U0_analytic_greens(L, mu, beta):
    g = rng(9)
    T = M.hopping_matrix("square", (L, L), t=1.0, mu=mu)
    Ms = M.n_slices(beta)
    c = R.RefChain(T, U=0.0, beta=beta, delta_tau=0.1, safe_mult=5, conf=rand_conf(g, L * L, Ms))
    c.init()
    c.local_sweep()           # thermalization = 1
    acc = np.zeros((L * L, L * L))
    for _ in range(2):        # sweeps = 2, measure_rate = 1
        c.local_sweep()
        acc += c.measured_greens()[:, :, 0]
    Gan = M.analytic_greens(T, beta)
    assert np.allclose(acc / 2, Gan, atol=1e-12, rtol=1e-12)
    assert c.stats["prop_count"] == 0 and c.stats["neg_count"] == 0


# ------------------------------------------------------------ test/DQMC/measurements.jl:168-189
def test_measured_greens_transform():
    g = rng(10)
    T = M.hopping_matrix("square", (4, 4))
    c = R.RefChain(T, U=-2.0, beta=1.0, conf=rand_conf(g, 16, 10))
    c.init()
    Gm = c.measured_greens()
    for b in range(2):
        assert np.allclose(Gm[:, :, b], c.eThalfinv @ c.greens[:, :, b] @ c.eThalf, rtol=1e-13)


# ------------------------------------------------------------ test/updates.jl:186-245
@pytest.mark.parametrize("U,mu", [(1.0, 0.5), (-1.0, 0.0)])
def test_local_ratio_product_equals_global_ratio(U, mu):
    g = rng(11)
    T = M.hopping_matrix("square", (2, 2), mu=mu)
    for trial in range(5):
        old = rand_conf(g, 4, 20)
        new = np.asfortranarray(g.permutation(old.ravel()).reshape(4, 20).astype(np.int8))
        c = R.RefChain(T, U=U, beta=2.0, conf=old)
        c.init()
        logp = 0.0
        for t in range(c.M):
            sl = c.state[0]
            for i in range(4):
                if c.get_conf()[i, sl - 1] != new[i, sl - 1]:
                    p = c.propose_local(i, accept=True)
                    assert p > 0
                    logp += np.log(p)
            c.propagate()
        for t in range(c.M):
            c.propagate()
        assert np.array_equal(c.get_conf(), new)
        assert c.state == (1, 1, 1)
        assert np.isclose(logp, log_weight(c, new) - log_weight(c, old), rtol=1e-9, atol=1e-9)
        for b in range(c.nb):
            assert np.allclose(c.greens[:, :, b], greens_brute(c, new, 1, b), atol=1e-10)


# ------------------------------------------------------------ full sweep vs brute force
@pytest.mark.parametrize("U", [4.0, -4.0])
def test_full_sweep_against_brute_force(U):
    g = rng(12)
    T = M.hopping_matrix("square", (4, 4))
    conf = rand_conf(g, 16, 5)
    c = R.RefChain(T, U=U, beta=0.5, safe_mult=2, conf=conf, seed=77, chain_id=3)
    c.init()
    for b in range(c.nb):
        assert np.allclose(c.greens[:, :, b], greens_brute(c, conf, 1, b), atol=1e-13)
    lw0 = log_weight(c, conf)
    acc, probs, dec = c.local_sweep(trace=True)
    assert acc == int(dec.sum()) and 0 < acc < dec.size
    new = c.get_conf()
    assert c.state == (1, 1, 1)
    for b in range(c.nb):
        assert np.allclose(c.greens[:, :, b], greens_brute(c, new, 1, b), atol=1e-12)
    # sum of log p over accepted flips == log[W(new) / W(old)]
    assert np.isclose(np.log(probs[dec == 1]).sum(), log_weight(c, new) - lw0, atol=1e-8)
    # replay with forced decisions reproduces conf and G bit-for-bit
    c2 = R.RefChain(T, U=U, beta=0.5, safe_mult=2, conf=conf)
    c2.init()
    c2.local_sweep(forced=dec)
    assert np.array_equal(c2.get_conf(), new)
    assert np.array_equal(c2.greens, c.greens)
    # and an explicit uniform table gives the same decisions as the counter RNG it was drawn from
    from oracle.rng import uniforms_for_sweep
    u = uniforms_for_sweep(77, 3, 0, 2 * c.M, c.N)
    c3 = R.RefChain(T, U=U, beta=0.5, safe_mult=2, conf=conf)
    c3.init()
    c3.local_sweep(uniforms=u)
    assert np.array_equal(c3.get_conf(), new)


def test_sweep_schedule_counts():
    """slices visited 1..M then M..1 and the state returns to (1, 1, +1)."""
    g = rng(13)
    T = M.hopping_matrix("chain", (4,))
    for Ms, sm in ((50, 10), (23, 10), (5, 3), (10, 10), (8, 1)):
        c = R.RefChain(T, U=1.0, slices=Ms, safe_mult=sm, conf=rand_conf(g, 4, Ms))
        c.init()
        assert c.state == (1, 1, 1)
        seen = []
        for _ in range(2 * Ms):
            seen.append(c.state[0])
            c.propagate()
        assert seen == list(range(1, Ms + 1)) + list(range(Ms, 0, -1))
        assert c.state == (1, 1, 1)
