"""GPU parity of the device-side global updates (montecarlo.jl_b200/csrc/global.cu, through the C ABI) against
the CPU oracle (oracle/dqmc_ref_global.inc.c) and the reference's backbone check (test/updates.jl:186-245).
Run with `-m gpu` on a B200.
"""
import numpy as np
import pytest

from oracle.bruteforce import log_weight
from oracle.rng import philox_uniform

from test_gpu_parity import GTOL, make_pair, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,Ls,U,beta,sm", [("square", (2, 2), 1.0, 2.0, 10), ("square", (2, 2), -1.0, 2.0, 10),
                                               ("square", (4, 4), 4.0, 1.0, 5), ("square", (6, 6), -4.0, 2.0, 10),
                                               ("honeycomb", (3, 3), -4.0, 1.0, 5)])
def test_global_update_matches_oracle(b200, kind, Ls, U, beta, sm):
    B = 4
    ctx, chains = make_pair(b200, kind, Ls, U=U, beta=beta, B=B, safe_mult=sm, mu=0.3)
    ctx.build_stack()
    g = np.random.default_rng(5)
    old = ctx.get_conf()
    prop = np.asfortranarray(np.stack([g.permutation(old[:, :, b].ravel()).reshape(old.shape[:2]) for b in range(B)],
                                      axis=2).astype(np.int8))                        # GlobalShuffle
    # chain 0 and 1: forced accept (u = 0), chain 2: forced reject unless p > 1 (u = 2), chain 3: u = 0.5
    u = np.array([0.0, 0.0, 2.0, 0.5])
    acc, p = ctx.global_update(sm, proposed=prop, uniforms=u)
    G = ctx.greens()
    conf = ctx.get_conf()
    assert ctx.state == (1, 1, 1)
    for b, c in enumerate(chains):
        c.init()
        a_ref, p_ref = c.global_update(prop[:, :, b], uniform=u[b])
        assert np.isclose(p[b], p_ref, rtol=1e-9), (b, p[b], p_ref)
        assert acc[b] == a_ref
        assert np.array_equal(conf[:, :, b], c.get_conf())
        assert relerr(G[:, :, :, b], c.greens) < GTOL
    assert acc[0] == 1 and acc[1] == 1 and (acc[2] == 0 or p[2] > 1.0)
    # the Markov chain continues identically after the update
    a2 = ctx.sweep(1)
    G = ctx.greens()
    for b, c in enumerate(chains):
        assert c.local_sweep() == a2[b]
        assert relerr(G[:, :, :, b], c.greens) < GTOL


def test_global_flip_probability_is_weight_ratio(b200):
    """GlobalFlip with the counter RNG: p = W(-conf) / W(conf) by brute force; decisions follow the ABI's uniform
    (step = 2M, site = 0)."""
    B = 6
    ctx, chains = make_pair(b200, "square", (2, 2), U=1.0, beta=2.0, B=B, safe_mult=10, mu=0.5, seed=31)
    ctx.build_stack()
    old = ctx.get_conf()
    acc, p = ctx.global_update(10)
    conf = ctx.get_conf()
    for b, c in enumerate(chains):
        lw = log_weight(c, -old[:, :, b]) - log_weight(c, old[:, :, b])
        assert np.isclose(np.log(p[b]), lw, rtol=1e-8, atol=1e-8)
        uu = float(philox_uniform(31, b, 0, 2 * c.M, 0))
        assert acc[b] == int(p[b] > 1.0 or uu < p[b])
        assert np.array_equal(conf[:, :, b], -old[:, :, b] if acc[b] else old[:, :, b])


def test_global_update_error_behaviour(b200):
    ctx, _ = make_pair(b200, "square", (2, 2), U=1.0, beta=1.0, B=1, safe_mult=5)
    with pytest.raises(b200.DQMCError):
        ctx.global_update(5)               # stack not built: not at (slice 1, direction +1)


def test_consecutive_global_updates_draw_different_uniforms(b200):
    """SimpleScheduler(LocalSweep(), GlobalFlip(), GlobalShuffle()) performs two global updates without a local sweep in
    between: their counter-RNG uniforms must differ (site word = running index of the global update), otherwise the
    second decision is correlated with the first.  Decisions are checked against the ABI's definition of u."""
    B = 8
    ctx, chains = make_pair(b200, "square", (2, 2), U=1.0, beta=2.0, B=B, safe_mult=10, mu=0.5, seed=77)
    ctx.build_stack()
    M = chains[0].M
    u0 = np.array([float(philox_uniform(77, b, 0, 2 * M, 0)) for b in range(B)])
    u1 = np.array([float(philox_uniform(77, b, 0, 2 * M, 1)) for b in range(B)])
    assert not np.any(u0 == u1)
    acc0, p0 = ctx.global_update(10)
    acc1, p1 = ctx.global_update(10)
    assert np.array_equal(acc0, (p0 > 1.0) | (u0 < p0))
    assert np.array_equal(acc1, (p1 > 1.0) | (u1 < p1))
    # a chain whose first flip was rejected sees the same p again; with a shared uniform it could never accept
    # now -- with independent uniforms it accepts iff u1 < p
    same = (acc0 == 0)
    assert np.allclose(p1[same], p0[same], rtol=1e-9)
    # resume: the running index is restorable like the sweep index
    ctx.set_global_update_index(0)
    acc2, p2 = ctx.global_update(10)
    assert np.array_equal(acc2, (p2 > 1.0) | (u0 < p2))
