"""Pins the global-update part of the C oracle (oracle/dqmc_ref_global.inc.c) against the reference's own
backbone check for global updates (test/updates.jl:186-245) and the brute-force determinant ratio.  No GPU.
"""
import numpy as np
import pytest

from oracle import model as M
from oracle import ref as R
from oracle.bruteforce import greens_brute, log_weight


def rand_conf(g, N, Ms):
    return np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(N, Ms)))


# models of the reference test: HubbardModel(2, 2, mu = 0.5) (U = 1) and HubbardModel(2, 2, U = -1), beta = 2
@pytest.mark.parametrize("U,mu", [(1.0, 0.5), (-1.0, 0.0), (4.0, 0.0), (-4.0, 0.3)])
def test_global_probability_equals_product_of_local_probabilities(U, mu):
    g = np.random.default_rng(21)
    T = M.hopping_matrix("square", (2, 2), mu=mu)
    for trial in range(6):
        old = rand_conf(g, 4, 20)
        new = np.asfortranarray(g.permutation(old.ravel()).reshape(4, 20).astype(np.int8))   # shuffle!(conf)
        c1 = R.RefChain(T, U=U, beta=2.0, conf=old)
        c2 = R.RefChain(T, U=U, beta=2.0, conf=old)
        c1.init(); c2.init()
        acc, global_p = c1.global_update(new, uniform=0.0)      # uniform 0 < p: always accepted (accept_global!)
        assert acc == 1
        local_p = 1.0
        for t in range(c2.M):
            sl = c2.state[0]
            for i in range(4):
                if c2.get_conf()[i, sl - 1] != new[i, sl - 1]:
                    local_p *= c2.propose_local(i, accept=True)
            c2.propagate()
        for t in range(c2.M):
            c2.propagate()
        assert np.isclose(local_p, global_p, rtol=1e-8)
        assert np.array_equal(c1.get_conf(), c2.get_conf())
        assert c1.state == c2.state == (1, 1, 1)
        assert np.allclose(c1.greens, c2.greens, atol=1e-10)
        # and against the definition: p = W(new) / W(old)
        assert np.isclose(np.log(global_p), log_weight(c1, new) - log_weight(c1, old), rtol=1e-8, atol=1e-8)
        for b in range(c1.nb):
            assert np.allclose(c1.greens[:, :, b], greens_brute(c1, new, 1, b), atol=1e-10)


@pytest.mark.parametrize("U", [4.0, -4.0])
def test_global_flip_accept_and_reject(U):
    """GlobalFlip (global_updates.jl:237-248): conf -> -conf; a rejected proposal restores the field and
    leaves the stack untouched."""
    g = np.random.default_rng(22)
    T = M.hopping_matrix("square", (4, 4), mu=0.4)
    conf = rand_conf(g, 16, 10)
    c = R.RefChain(T, U=U, beta=1.0, safe_mult=5, conf=conf)
    c.init()
    G0 = c.greens.copy()
    acc, p = c.global_update(-conf, uniform=2.0)                # uniform > 1 >= min(p, 1): rejected unless p > 1
    if p <= 1.0:
        assert acc == 0
        assert np.array_equal(c.get_conf(), conf)
        assert np.array_equal(c.greens, G0)
    assert np.isclose(np.log(p), log_weight(c, -conf) - log_weight(c, conf), rtol=1e-8, atol=1e-8)
    c2 = R.RefChain(T, U=U, beta=1.0, safe_mult=5, conf=conf)
    c2.init()
    acc, p2 = c2.global_update(-conf, uniform=0.0)
    assert acc == 1 and np.isclose(p, p2, rtol=1e-12)
    assert np.array_equal(c2.get_conf(), -conf)
    for b in range(c2.nb):
        assert np.allclose(c2.greens[:, :, b], greens_brute(c2, -conf, 1, b), atol=1e-10)
