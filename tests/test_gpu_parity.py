"""GPU parity tests: the CUDA library (through the C ABI) against the CPU oracle, the
reference's known-answer tests, and size-independent properties.  Run with `-m gpu` on a B200.

Tolerances: G and every matrix result <= 1e-10 relative to max|G| (north_star), accept/reject
sequences identical (disagreement is only tolerated where |p - u| < 1e-9, and is reported).
"""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import model as OM
from oracle import ref as OR
from oracle.rng import uniforms_for_sweep

pytestmark = pytest.mark.gpu

GTOL = 1e-10


def rng(seed=0):
    return np.random.default_rng(seed)


def rand_confs(g, N, M, B, ghq=False):
    vals = np.array([1, 2, 3, 4] if ghq else [-1, 1], dtype=np.int8)
    return np.asfortranarray(g.choice(vals, size=(N, M, B)))


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def make_pair(b200, kind, Ls, *, U, beta, B=2, safe_mult=10, mu=0.0, seed=11, delta_tau=0.1, delay_block=0,
              check_prop=True, update_variant=0, field_kind=None):
    """-> (Context, [RefChain]) on the same model, conf and RNG key."""
    T = OM.hopping_matrix(kind, Ls, mu=mu)
    N = T.shape[0]
    M = OM.n_slices(beta, delta_tau)
    fk = OM.choose_field(U) if field_kind is None else field_kind
    alpha = OM.field_alpha(U, delta_tau, fk)
    e2, e2i, eh, ehi = OM.hopping_exponentials(T, delta_tau)
    confs = rand_confs(rng(seed), N, M, B, ghq=fk >= 2)
    ctx = b200.Context(n_sites=N, n_slices=M, field_kind=fk, n_chains=B,
                       ranges=OM.generate_chunks(M, safe_mult), alpha=alpha, hopping_exp_squared=e2,
                       hopping_exp_inv_squared=e2i, hopping_exp=eh, hopping_exp_inv=ehi, seed=seed,
                       delay_block=delay_block, check_propagation_error=check_prop, update_variant=update_variant)
    ctx.set_conf(confs)
    chains = [OR.RefChain(T, U=U, beta=beta, delta_tau=delta_tau, safe_mult=safe_mult, seed=seed, chain_id=b,
                          conf=confs[:, :, b], check_propagation_error=check_prop, field_kind=fk) for b in range(B)]
    return ctx, chains


# ===================================================================== operator level
@pytest.mark.parametrize("n", [8, 16, 49, 64, 100, 144])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_vmul(b200, n, ta, tb):
    """test/linalg.jl:16-93"""
    g = rng(n)
    A, B = g.random((n, n, 3)), g.random((n, n, 3))
    C = b200.vmul(A, B, ta, tb)
    for i in range(3):
        ref = (A[:, :, i].T if ta else A[:, :, i]) @ (B[:, :, i].T if tb else B[:, :, i])
        assert np.allclose(C[:, :, i], ref, atol=100 * n * np.finfo(float).eps, rtol=0)


@pytest.mark.parametrize("n", [4, 8, 16, 36, 49, 64, 100, 144, 256, 288])
def test_udt_identities(b200, n):
    """test/linalg.jl:97-124 at every kernel geometry (cluster sizes 1, 2, 4)."""
    g = rng(n)
    X = g.random((n, n, 3))
    X[:, :, 1] *= np.exp(3 * g.normal(size=n))[None, :]            # graded columns like B-chains
    X[:, :, 2] = np.kron(g.random(n)[:, None], g.random(n)[None, :])  # rank 1 (kron input of the reference test)
    U, D, T, piv = b200.udt_AVX_pivot(X, apply_pivot=True)
    for i in range(3):
        assert relerr(U[:, :, i] @ np.diag(D[:, i]) @ T[:, :, i], X[:, :, i]) < 1e-12
        assert np.abs(U[:, :, i].T @ U[:, :, i] - np.eye(n)).max() < 1e-12
        assert sorted(piv[:, i]) == list(range(1, n + 1))
    U2, D2, T2, piv2 = b200.udt_AVX_pivot(X, apply_pivot=False)
    for i in range(3):
        P = np.zeros((n, n))
        P[np.arange(n), piv2[:, i] - 1] = 1.0
        assert np.allclose(np.tril(T2[:, :, i], -1), 0.0)
        assert relerr(U2[:, :, i] @ np.diag(D2[:, i]) @ np.triu(T2[:, :, i]) @ P, X[:, :, i]) < 1e-12
    # same D and pivots as the oracle (and LAPACK dgeqp3) on the generic matrices
    for i in range(2):
        _, Do, _, po = OR.udt_pivot(X[:, :, i])
        assert np.array_equal(piv[:, i] - 1, po)
        assert np.allclose(D[:, i], Do, rtol=1e-10)


def test_udt_zero_column(b200):
    """UDT.jl:293-301: exact zeros on diag(R) become D = 1."""
    X = np.zeros((8, 8)); X[:, 0] = np.arange(1.0, 9.0)
    U, D, T, piv = b200.udt_AVX_pivot(X)
    assert np.all(D[1:] == 1.0)
    assert np.allclose(U @ np.diag(D) @ T, X)


@pytest.mark.parametrize("n", [8, 33, 64, 144, 256])
def test_rdivp(b200, n):
    """test/linalg.jl:126-131"""
    g = rng(n)
    X = g.random((n, n, 2))
    U, D, T, piv = b200.udt_AVX_pivot(X, apply_pivot=False)
    A = g.random((n, n, 2))
    out = b200.rdivp(A, T, piv)
    for i in range(2):
        ref = OR.rdivp(A[:, :, i], T[:, :, i], piv[:, i] - 1)
        assert relerr(out[:, :, i], ref) < 1e-9 * max(1.0, np.linalg.cond(np.triu(T[:, :, i])) * 1e-6)
        P = np.zeros((n, n)); P[np.arange(n), piv[:, i] - 1] = 1.0
        assert np.allclose(out[:, :, i] @ np.triu(T[:, :, i]), A[:, :, i] @ P.T, atol=1e-9)


@pytest.mark.parametrize("n", [16, 64, 144])
def test_calculate_greens_AVX(b200, n):
    """test/updates.jl:153-180: G from two UDTs; vs the oracle and vs the direct inverse."""
    g = rng(n)
    Bl, Br = g.random((n, n, 2)), g.random((n, n, 2))
    Ul, Dl, Tl, _ = b200.udt_AVX_pivot(Bl)
    Ur, Dr, Tr, _ = b200.udt_AVX_pivot(Br)
    G = b200.calculate_greens_AVX(Ul, Dl, Tl, Ur, Dr, Tr)
    for i in range(2):
        Go = OR.calculate_greens_udt(Ul[:, :, i], Dl[:, i], Tl[:, :, i], Ur[:, :, i], Dr[:, i], Tr[:, :, i])
        assert relerr(G[:, :, i], Go) < 1e-9
        direct = np.linalg.inv(np.eye(n) + Bl[:, :, i] @ Br[:, :, i].T)
        assert relerr(G[:, :, i], direct) < 1e-8


@pytest.mark.parametrize("U", [1.0, -1.0])
def test_slice_matrices(b200, U):
    """test/slice_matrices.jl:13-42"""
    ctx, chains = make_pair(b200, "chain", (8,), U=U, beta=5.0, B=2)
    g = rng(3)
    for sl in (7, 33):
        X = np.asfortranarray(g.random((8, 8, ctx.nb, 2)))
        for which in ("left", "right", "inv_left", "inv_right", "daggered_left"):
            Y = ctx.multiply_slice_matrix(which, sl, X)
            for b, c in enumerate(chains):
                assert relerr(Y[:, :, :, b], c.multiply_slice_matrix(which, sl, X[:, :, :, b])) < 1e-13
        for d in (1, -1):
            cs = sl if d == 1 else sl + 1
            Y = ctx.wrap_greens(X, cs, d)
            for b, c in enumerate(chains):
                assert relerr(Y[:, :, :, b], c.wrap_greens(X[:, :, :, b], cs, d)) < 1e-12


# ===================================================================== stack
def udt_product(ctx_or_chain, slot, chain=None, oracle=False):
    if oracle:
        c = ctx_or_chain
        U, D, T = c.array("u_stack", slot - 1), c.array("d_stack", slot - 1), c.array("t_stack", slot - 1)
    else:
        U = ctx_or_chain.stack_array("u_stack", slot, chain)
        D = ctx_or_chain.stack_array("d_stack", slot, chain)
        T = ctx_or_chain.stack_array("t_stack", slot, chain)
    return np.stack([U[:, :, b] @ np.diag(D[:, b]) @ T[:, :, b] for b in range(U.shape[2])], axis=2), D


@pytest.mark.parametrize("U", [1.0, -2.0])
def test_build_stack_matches_oracle(b200, U):
    """test/flavortests_DQMC.jl:282-341: G after init, every stack slot (as U*D*T -- UDT is not
    unique), forward build + M propagates == reverse build + 1 propagate."""
    ctx, chains = make_pair(b200, "chain", (8,), U=U, beta=5.0, B=3)
    ctx.build_stack()
    assert ctx.state == (1, 1, 1)
    G = ctx.greens()
    for b, c in enumerate(chains):
        c.init()
        assert relerr(G[:, :, :, b], c.greens) < GTOL
        for slot in range(2, ctx.C + 2):
            P, D = udt_product(ctx, slot, b)
            Po, Do = udt_product(c, slot, oracle=True)
            assert relerr(P, Po) < 1e-9
            assert np.allclose(D, Do, rtol=1e-8)
    ctx.forward_build_stack()
    assert ctx.state == (51, 5, -1)
    ctx.propagate(1 + ctx.M)
    assert ctx.state == (1, 1, 1)
    G2 = ctx.greens()
    assert relerr(G2, G) < 1e-9
    for b, c in enumerate(chains):
        assert relerr(G2[:, :, :, b], c.greens) < 1e-9


def test_calculate_greens_at_every_slice(b200):
    """test/flavortests_DQMC.jl:303-311"""
    ctx, chains = make_pair(b200, "chain", (8,), U=-1.0, beta=3.0, B=2, safe_mult=5)
    for k in (0, 1, 4, 5, 17, 29, 30):
        G = ctx.calculate_greens_at(k, 5)
        for b, c in enumerate(chains):
            assert relerr(G[:, :, :, b], c.calculate_greens_at(k, 5)) < GTOL


@pytest.mark.parametrize("L,mu", [(7, 0.0), (8, 1.0)])
@pytest.mark.parametrize("beta", [1.0, 10.0])
def test_U0_analytic_greens(b200, L, mu, beta):
    """test/flavortests_DQMC.jl:355-386 through the user API: mean(mc[:G]) == analytic G at 1e-12."""
    model = b200.HubbardModel(b200.SquareLattice(L), U=0.0, mu=mu)
    mc = b200.DQMC(model, beta=beta, delta_tau=0.1, safe_mult=5, thermalization=1, sweeps=2, measure_rate=1,
                   seed=5, n_chains=2)
    mc["G"] = b200.greens_measurement(mc, model)
    assert b200.run(mc) == "SUCCESS"
    Gan = OM.analytic_greens(OM.hopping_matrix("square", (L, L), mu=mu), beta)
    assert np.allclose(mc["G"].mean(), Gan, atol=1e-12, rtol=1e-12)
    assert mc["G"].count == 2
    assert all(s["prop_count"] == 0 and s["neg_count"] == 0 for s in mc.analysis())


def test_measured_greens(b200):
    """test/DQMC/measurements.jl:168-189"""
    ctx, chains = make_pair(b200, "square", (4, 4), U=-2.0, beta=1.0, B=2)
    ctx.build_stack()
    Gm = ctx.measured_greens()
    for b, c in enumerate(chains):
        c.init()
        assert relerr(Gm[:, :, :, b], c.measured_greens()) < GTOL


# ===================================================================== the sweep
def check_sweeps(ctx, chains, nsweeps, gtol=GTOL, ptol=1e-7):
    """ptol: tolerance on the acceptance probabilities *between* stabilisations.  G is only 1e-10-exact
    right after a stabilisation; in between, wraps amplify round-off (the reference itself reports
    propagation errors of 1e-7..1e-6 as normal, docs/src/examples/ALF1.md:103-104), so long-beta cases
    compare p at that scale.  Decisions must still be identical."""
    ctx.build_stack()
    for c in chains:
        c.init()
    for s in range(nsweeps):
        acc, probs, dec = ctx.sweep_traced()
        G, conf = ctx.greens(), ctx.get_conf()
        for b, c in enumerate(chains):
            a, po, do = c.local_sweep(trace=True)
            mism = np.argwhere(dec[b] != do)
            if len(mism):
                st, si = mism[0]
                u = uniforms_for_sweep(c_seed(ctx, c), b, s, 2 * ctx.M, ctx.N)[st, si]
                raise AssertionError(f"decision mismatch chain {b} sweep {s} step {st} site {si}: "
                                     f"p_gpu={probs[b, st, si]!r} p_ref={po[st, si]!r} u={u!r}")
            assert np.allclose(probs[b], po, rtol=ptol, atol=1e-12), np.abs(probs[b] / po - 1).max()
            assert a == acc[b]
            assert np.array_equal(conf[:, :, b], c.get_conf())
            assert relerr(G[:, :, :, b], c.greens) < gtol
    assert ctx.state == (1, 1, 1)


def c_seed(ctx, c):
    return 11


@pytest.mark.parametrize("U", [4.0, -4.0])
@pytest.mark.parametrize("beta,safe_mult", [(0.5, 2), (2.0, 10), (5.0, 10)])
def test_sweep_parity_4x4(b200, U, beta, safe_mult):
    """config 1 family: free-running sweeps with the shared counter RNG take identical decisions."""
    ctx, chains = make_pair(b200, "square", (4, 4), U=U, beta=beta, B=3, safe_mult=safe_mult)
    check_sweeps(ctx, chains, 2)


@pytest.mark.parametrize("U", [4.0, -4.0])
def test_sweep_parity_8x8(b200, U):
    ctx, chains = make_pair(b200, "square", (8, 8), U=U, beta=2.0, B=2)
    check_sweeps(ctx, chains, 1)


def test_sweep_parity_odd_sizes_and_ragged_blocks(b200):
    """N = 49 (odd leading dimension), delay block 12 (49 = 4 * 12 + 1), M = 23 (ragged ranges)."""
    ctx, chains = make_pair(b200, "square", (7, 7), U=-3.0, beta=2.3, B=2, delay_block=12)
    check_sweeps(ctx, chains, 1)
    ctx, chains = make_pair(b200, "honeycomb", (3, 3), U=2.0, beta=0.7, B=1, safe_mult=3, delay_block=4)
    check_sweeps(ctx, chains, 2)


def test_sweep_single_range_and_single_slice(b200):
    ctx, chains = make_pair(b200, "square", (4, 4), U=-4.0, beta=0.3, B=2, safe_mult=10)   # M = 3, C = 1
    check_sweeps(ctx, chains, 2)
    ctx, chains = make_pair(b200, "chain", (6,), U=2.0, beta=0.1, B=1)                     # M = 1
    check_sweeps(ctx, chains, 2)


def test_sweep_12x12_one_sweep(b200):
    """config 3 geometry (n = 144, two flavor blocks, cluster-of-2 QR) at short beta."""
    ctx, chains = make_pair(b200, "square", (12, 12), U=-4.0, beta=0.6, B=2, safe_mult=3)
    check_sweeps(ctx, chains, 1)


def test_uniform_table_equals_counter_rng_and_forcing(b200):
    ctx, chains = make_pair(b200, "square", (4, 4), U=-4.0, beta=1.0, B=2)
    ctx.build_stack()
    acc0, p0, d0 = ctx.sweep_traced()
    G0, c0 = ctx.greens(), ctx.get_conf()
    # same sweep from an explicit table of the same uniforms
    ctx2, _ = make_pair(b200, "square", (4, 4), U=-4.0, beta=1.0, B=2)
    ctx2.build_stack()
    u = np.stack([uniforms_for_sweep(11, b, 0, 2 * ctx.M, ctx.N) for b in range(2)])
    acc1, p1, d1 = ctx2.sweep_traced(uniforms=u)
    assert np.array_equal(d0, d1) and np.array_equal(c0, ctx2.get_conf())
    assert np.array_equal(G0, ctx2.greens())
    # teacher forcing with the recorded decisions
    ctx3, _ = make_pair(b200, "square", (4, 4), U=-4.0, beta=1.0, B=2)
    ctx3.build_stack()
    acc2, p2, d2 = ctx3.sweep_traced(forced=d0)
    assert np.array_equal(d2, d0) and np.array_equal(ctx3.get_conf(), c0)
    assert np.array_equal(G0, ctx3.greens())
    # dqmc_sweep (no traces) is the same path
    ctx4, _ = make_pair(b200, "square", (4, 4), U=-4.0, beta=1.0, B=2)
    ctx4.build_stack()
    assert np.array_equal(ctx4.sweep(1), acc0)
    assert np.array_equal(ctx4.get_conf(), c0)


def test_delay_block_size_does_not_change_decisions(b200):
    res = []
    for kb in (4, 8, 16, 0):
        ctx, _ = make_pair(b200, "square", (6, 6), U=-4.0, beta=1.0, B=2, delay_block=kb)
        ctx.build_stack()
        acc, p, d = ctx.sweep_traced()
        res.append((d, ctx.greens()))
    for d, G in res[1:]:
        assert np.array_equal(d, res[0][0])
        assert relerr(G, res[0][1]) < 1e-11


@pytest.mark.parametrize("U,mu", [(1.0, 0.5), (-1.0, 0.0)])
def test_local_ratio_product_equals_global_ratio(b200, U, mu):
    """test/updates.jl:186-245 on the GPU: force exactly the flips that turn conf into a shuffled
    conf; sum log p == log W(new) - log W(old) (brute force), final G == G(new conf)."""
    g = rng(21)
    ctx, chains = make_pair(b200, "square", (2, 2), U=U, beta=2.0, B=2, mu=mu)
    old = ctx.get_conf()
    new = np.asfortranarray(np.stack([g.permutation(old[:, :, b].ravel()).reshape(4, 20) for b in range(2)], axis=2))
    ctx.build_stack()
    M, N = ctx.M, ctx.N
    forced = np.zeros((2, 2 * M, N), dtype=np.uint8)
    for b in range(2):
        forced[b, :M, :] = (old[:, :, b] != new[:, :, b]).T        # up direction visits slices 1..M
    acc, probs, dec = ctx.sweep_traced(forced=forced)
    assert np.array_equal(ctx.get_conf(), new)
    from oracle.bruteforce import greens_brute, log_weight
    G = ctx.greens()
    for b, c in enumerate(chains):
        lw = log_weight(c, new[:, :, b]) - log_weight(c, old[:, :, b])
        assert np.isclose(np.log(probs[b][forced[b] == 1]).sum(), lw, rtol=1e-9, atol=1e-9)
        for f in range(ctx.nb):
            assert np.allclose(G[:, :, f, b], greens_brute(c, new[:, :, b], 1, f), atol=1e-10)


def test_sweep_spatial_single_slice(b200):
    """test/fields.jl:133-222: one slice of proposals, G vs the dense rank-1 formula via the oracle."""
    ctx, chains = make_pair(b200, "square", (4, 4), U=-4.0, beta=1.0, B=2)
    ctx.build_stack()
    acc, probs, dec = ctx.sweep_spatial()
    G = ctx.greens()
    for b, c in enumerate(chains):
        c.init()
        a, po, do = c.sweep_spatial()
        assert np.array_equal(dec[b], do)
        assert np.allclose(probs[b], po, rtol=1e-12)
        assert relerr(G[:, :, :, b], c.greens) < 1e-13


def test_sign_problem_statistics(b200):
    """local_updates.jl:40-46 + statistics.jl:9-38: repulsive model off half filling has p < 0."""
    ctx, chains = make_pair(b200, "square", (4, 4), U=-6.0, beta=3.0, B=3, mu=1.5)
    check_sweeps(ctx, chains, 2)
    st = ctx.stats()
    for b, c in enumerate(chains):
        so = c.stats
        assert st[b]["neg_count"] == so["neg_count"]
        if so["neg_count"]:
            assert np.isclose(st[b]["neg_sumlog10"], so["neg_sumlog"], rtol=1e-9)
            assert np.isclose(st[b]["neg_min"], so["neg_min"], rtol=1e-7)
    assert sum(s["neg_count"] for s in st) > 0


# ===================================================================== full-size properties
def test_16x16_sweep_properties(b200):
    """config 4 geometry (n = 256, 2 blocks, cluster-of-4 QR): after a sweep the propagated G equals
    the from-scratch G of the new configuration, and one oracle chain agrees at short beta."""
    ctx, chains = make_pair(b200, "square", (16, 16), U=-4.0, beta=0.4, B=2, safe_mult=2, seed=3)
    ctx.build_stack()
    for c in chains[:1]:
        c.init()
    acc, probs, dec = ctx.sweep_traced()
    G = ctx.greens()
    a, po, do = chains[0].local_sweep(trace=True)
    assert np.array_equal(dec[0], do)
    assert relerr(G[:, :, :, 0], chains[0].greens) < GTOL
    # the stack's G at current_slice = 1 is calculate_greens(mc, 0) = [I + B_M ... B_1]^-1: against the library's own
    # from-scratch evaluation and against the extended-precision arbiter (oracle/truth_ld.c) for both chains
    Gs = ctx.calculate_greens_at(0, 2)
    assert relerr(G, Gs) < GTOL
    from oracle import truth as TR
    conf = ctx.get_conf()
    for b, c in enumerate(chains):
        assert relerr(G[:, :, :, b], TR.greens_truth_chain(c, conf=conf[:, :, b])) < GTOL


# ===================================================================== errors
def test_error_behaviour(b200):
    ctx, _ = make_pair(b200, "square", (4, 4), U=1.0, beta=1.0, B=1)
    with pytest.raises(b200.DQMCError):
        ctx.sweep(1)                       # stack not built
    bad = np.zeros((16, 10, 1), dtype=np.int8)
    with pytest.raises(b200.DQMCError):
        ctx.set_conf(bad)
    with pytest.raises(b200.DQMCError):
        b200.Context(n_sites=4, n_slices=10, field_kind=0, n_chains=1, ranges=[(1, 4)], alpha=0.1,
                     hopping_exp_squared=np.eye(4), hopping_exp_inv_squared=np.eye(4), hopping_exp=np.eye(4),
                     hopping_exp_inv=np.eye(4))


def test_observable_accumulators(b200):
    ctx, chains = make_pair(b200, "square", (4, 4), U=-4.0, beta=1.0, B=4)
    ctx.build_stack()
    ctx.accumulate_greens()
    ctx.sweep(1)
    ctx.accumulate_greens()
    cnt, s, s2 = ctx.observables()
    assert cnt == 8
    ptr, n = ctx.observable_buffer()
    assert ptr and n == 1 + 2 * ctx.nb * 16 * 16
    Gm = ctx.measured_greens()
    assert s.shape == (16, 16, 2) and np.all(s2 >= 0)
    assert np.abs(s).max() > 0 and np.isfinite(Gm).all()


# ===================================================================== long imaginary time
@pytest.mark.parametrize("U", [4.0, -4.0])
def test_sweep_parity_8x8_long_beta(b200, U):
    """config 2 geometry (8x8, attractive, beta = 10) and its repulsive sibling at beta = 8: D spans
    ~40 orders of magnitude, every stabilisation matters.  Decisions identical, G <= 1e-10 at sweep end,
    and the propagation-error statistics (stack.jl:644-654) agree with the oracle's."""
    beta = 10.0 if U > 0 else 8.0
    ctx, chains = make_pair(b200, "square", (8, 8), U=U, beta=beta, B=2, seed=5)
    check_sweeps(ctx, chains, 1, ptol=2e-5)
    st = ctx.stats()
    for b, c in enumerate(chains):
        assert st[b]["prop_count"] == c.stats["prop_count"]
        if c.stats["prop_count"]:
            assert np.isclose(st[b]["prop_max"], c.stats["prop_max"], rtol=0.5)


def test_honeycomb_L4_sweep(b200):
    """config 5 family (honeycomb, attractive): N = 32 sites, two-site basis ordering."""
    ctx, chains = make_pair(b200, "honeycomb", (4, 4), U=4.0, beta=2.0, B=2)
    check_sweeps(ctx, chains, 2)


def test_check_propagation_error_off_is_identical(b200):
    """parameters.check_propagation_error = false only skips the check (stack.jl:636-654)."""
    out = []
    for chk in (True, False):
        ctx, _ = make_pair(b200, "square", (4, 4), U=-4.0, beta=2.0, B=2, check_prop=chk)
        ctx.build_stack()
        acc, p, d = ctx.sweep_traced()
        out.append((d, ctx.greens()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


# ===================================================================== every local-update kernel variant
@pytest.mark.parametrize("version", [1, 3])
def test_every_update_kernel_variant(b200, version):
    """The library picks update.cu (delayed rank-kb factors) below n = 96 and update3.cu (submatrix form) above
    (dqmc_desc.update_variant forces one).  Both must take the oracle's decisions on every geometry: one / two
    flavor blocks, odd n, ragged delay blocks, single range, and n > kb (several blocks per slice)."""
    cases = [("square", (4, 4), 4.0, 1.0, 3, 10, 0), ("square", (6, 6), -4.0, 1.0, 2, 5, 0),
             ("square", (7, 7), -3.0, 1.3, 2, 10, 12), ("honeycomb", (3, 3), 2.0, 0.7, 1, 3, 4),
             ("square", (10, 10), -4.0, 0.4, 2, 10, 0), ("square", (12, 12), 4.0, 0.3, 2, 10, 0)]
    for kind, Ls, U, beta, B, sm, db in cases:
        ctx, chains = make_pair(b200, kind, Ls, U=U, beta=beta, B=B, safe_mult=sm, delay_block=db, update_variant=version)
        check_sweeps(ctx, chains, 1)
    # traces, forced decisions and the uniform table go through the same kernel
    ctx, chains = make_pair(b200, "square", (6, 6), U=-4.0, beta=0.5, B=2, update_variant=version)
    ctx.build_stack()
    for c in chains:
        c.init()
    acc, probs, dec = ctx.sweep_traced()
    for b, c in enumerate(chains):
        a_ref, p_ref, d_ref = c.local_sweep(trace=True)
        assert a_ref == acc[b] and np.array_equal(dec[b], d_ref)
        assert np.allclose(probs[b], p_ref, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("kind,Ls,U", [("square", (16, 16), -4.0), ("honeycomb", (12, 12), 4.0)])
def test_sweep_parity_at_bench_sizes(b200, kind, Ls, U):
    """The kernel geometries of cfg 4 (n = 256, two flavor blocks: update3 with kb = 44, QR levels 256/192/128/64) and
    cfg 5 (n = 288, one block: kb = 36, cluster-of-8 QR) at short beta: decisions identical, G <= 1e-10 after a sweep (the stated beta of both configurations
    is covered by tests/test_gpu_parity_configs.py)."""
    ctx, chains = make_pair(b200, kind, Ls, U=U, beta=0.3, B=2, safe_mult=2)
    check_sweeps(ctx, chains, 1)


def test_conf_packed_on_device(b200):
    """dqmc_get_conf_packed / dqmc_set_conf_packed: BitArray(conf .== 1) chunks (fields.jl:331-334) packed on the
    device equal the host-side compress, for a bit count that is not a multiple of 64."""
    ctx, chains = make_pair(b200, "square", (7, 7), U=-3.0, beta=1.3, B=3)          # 49 x 13 = 637 bits
    conf = ctx.get_conf()
    packed = ctx.get_conf_packed()
    assert packed.shape == ((49 * 13 + 63) // 64, 3)
    for b in range(3):
        flat = conf[:, :, b].ravel(order="F")
        by = np.packbits(flat == 1, bitorder="little")
        by = np.concatenate([by, np.zeros((-len(by)) % 8, dtype=np.uint8)])
        assert np.array_equal(packed[:, b], by.view("<u8"))
    ctx.set_conf(-conf)
    ctx.set_conf_packed(packed[:, 1:], chain0=1)
    now = ctx.get_conf()
    assert np.array_equal(now[:, :, 0], -conf[:, :, 0]) and np.array_equal(now[:, :, 1:], conf[:, :, 1:])


@pytest.mark.parametrize("kind,Ls,U,fk", [("square", (4, 4), 4.0, None), ("square", (6, 6), -4.0, None),
                                          ("square", (10, 10), -4.0, None), ("square", (4, 4), -4.0, 3)])
def test_graph_replayed_sweeps_match_oracle(b200, kind, Ls, U, fk):
    """dqmc_sweep replays a captured CUDA graph from the second sweep of a context on (the launch schedule of a sweep is
    data independent; the sweep index is read from device memory).  Five sweeps -- eager, capture, three replays --
    with the counter RNG and with explicit uniform tables must follow the oracle exactly, and a traced (eager) sweep in
    between must not disturb the captured one."""
    for use_table in (False, True):
        ctx, chains = make_pair(b200, kind, Ls, U=U, beta=1.0, B=2, safe_mult=5, field_kind=fk)
        ctx.build_stack()
        for c in chains:
            c.init()
        for s in range(5):
            if s == 3:                                           # an eager traced sweep between replays
                acc, probs, dec = ctx.sweep_traced()
                refs = [c.local_sweep(trace=True) for c in chains]
                for b in range(2):
                    assert np.array_equal(dec[b], refs[b][2])
            else:
                u = None
                if use_table:
                    u = np.stack([uniforms_for_sweep(11, b, s, 2 * ctx.M, ctx.N, ghq=ctx.ghq) for b in range(2)])[None]
                acc = ctx.sweep(1, uniforms=u)
                refs = [(c.local_sweep(),) for c in chains]
            G, conf = ctx.greens(), ctx.get_conf()
            for b, c in enumerate(chains):
                assert refs[b][0] == acc[b], (s, b)
                assert np.array_equal(conf[:, :, b], c.get_conf())
                assert relerr(G[:, :, :, b], c.greens) < GTOL
        assert ctx.state == (1, 1, 1)
        assert ctx.kernel_launches() > 0
