"""The 4-state Gauss-Hermite fields (DensityGHQField / MagneticGHQField, fields.jl:464-637) in the oracle: the reference's
own known-answer tests (test/fields.jl:70-92 lookup tables, :95-130 compress round trip, test/updates.jl:186-245 local
vs global ratio with a shuffled configuration) and the definition-level weight ratio.  CPU only."""
import math

import mpmath as mp
import numpy as np
import pytest

from oracle import model as OM
from oracle import ref as OR
from oracle.bruteforce import greens_brute, log_weight
from oracle.rng import ghq_choice, philox_uniform, uniforms_for_sweep


def rand_ghq(seed, N, M):
    return np.asfortranarray(np.random.default_rng(seed).integers(1, 5, size=(N, M)).astype(np.int8))


def test_lookup_tables():
    """test/fields.jl:70-92: gamma, eta against BigFloat at rtol 1e-15, the choices matrix, alpha."""
    eta, gam, choices = OM.ghq_tables()
    mp.mp.prec = 256
    s6 = mp.sqrt(6)
    want_g = [1 - s6 / 3, 1 + s6 / 3, 1 + s6 / 3, 1 - s6 / 3]
    want_e = [-mp.sqrt(2 * (3 + s6)), -mp.sqrt(2 * (3 - s6)), mp.sqrt(2 * (3 - s6)), mp.sqrt(2 * (3 + s6))]
    for k in range(4):
        assert abs(gam[k] / float(want_g[k]) - 1) < 1e-15
        assert abs(eta[k] / float(want_e[k]) - 1) < 1e-15
    assert choices.tolist() == [[2, 3, 4], [1, 3, 4], [1, 2, 4], [1, 2, 3]]
    for x in range(1, 5):                                  # dqmc_ghq_choice reproduces choices[x, r]
        for r in range(3):
            assert ghq_choice(x, (r + 0.5) / 3) == choices[x - 1, r]
    assert OM.ghq_alpha(-1.0, 0.1, 3) == math.sqrt(0.05) and OM.ghq_alpha(1.0, 0.1, 2) == math.sqrt(0.05)
    with pytest.raises(ValueError):
        OM.ghq_alpha(1.0, 0.1, 3)                          # complex coupling: out of scope


def test_compress_roundtrip():
    """test/fields.jl:117-124 + fields.jl:476-489: (1,2,3,4) -> (00,01,10,11), high bit first."""
    conf = rand_ghq(1, 7, 13)
    chunks = OM.ghq_compress(conf)
    assert len(chunks) == (2 * 7 * 13 + 63) // 64
    assert np.array_equal(OM.ghq_decompress(chunks, (7, 13)), conf)
    one = np.array([[1, 2, 3, 4]], dtype=np.int8)          # bits 00 01 10 11 in BitArray order
    assert int(OM.ghq_compress(one)[0]) == 0b11_01_10_00    # little-endian bit positions 0..7 = 0,0, 0,1, 1,0, 1,1


@pytest.mark.parametrize("kind,U", [(2, 1.0), (3, -1.0), (2, 3.0), (3, -3.0)])
def test_local_probabilities_are_weight_ratios(kind, U):
    """prod of local p over a sequence of accepted proposals == W(new) / W(old) from the definition
    W = prod gamma(x) exp(-alpha eta(x) [density]) prod_flavors det(1 + B_M ... B_1), and G follows."""
    g = np.random.default_rng(5)
    T = OM.hopping_matrix("square", (2, 2), mu=0.3)
    c = OR.RefChain(T, U=U, beta=1.0, field_kind=kind, conf=rand_ghq(2, 4, 10))
    c.init()
    old = c.get_conf().copy()
    lp = 0.0
    for i in range(4):
        lp += math.log(abs(c.propose_local(i, accept=True, u_choice=g.random())))
    new = c.get_conf()
    assert np.all(new[:, 0] != old[:, 0]) and np.array_equal(new[:, 1:], old[:, 1:])
    assert abs(lp - (log_weight(c, new) - log_weight(c, old))) < 1e-9
    for b in range(c.nb):
        assert np.abs(c.greens[:, :, b] - greens_brute(c, new, 1, b)).max() < 1e-10


@pytest.mark.parametrize("kind,U", [(3, -1.0), (2, 1.0)])
def test_local_vs_global_with_shuffle(kind, U):
    """test/updates.jl:186-245: a shuffled configuration reached by forced local updates gives the same probability as
    the global update (gamma factors cancel for a permutation) and the same G."""
    g = np.random.default_rng(9)
    T = OM.hopping_matrix("square", (2, 2), mu=0.5)
    conf = rand_ghq(4, 4, 20)
    new = np.asfortranarray(g.permutation(conf.ravel()).reshape(4, 20))
    c1 = OR.RefChain(T, U=U, beta=2.0, field_kind=kind, conf=conf); c1.init()
    c2 = OR.RefChain(T, U=U, beta=2.0, field_kind=kind, conf=conf); c2.init()
    acc, p_global = c1.global_update(new, uniform=0.0)
    assert acc == 1
    lp = 0.0
    choices = OM.ghq_tables()[2]
    for t in range(c2.M):
        sl = c2.state[0]
        for i in range(4):
            xo, xn = c2.get_conf()[i, sl - 1], new[i, sl - 1]
            if xo != xn:
                r = list(choices[xo - 1]).index(xn)
                lp += math.log(c2.propose_local(i, accept=True, u_choice=(r + 0.5) / 3))
        c2.propagate()
    for t in range(c2.M):
        c2.propagate()
    assert np.array_equal(c1.get_conf(), c2.get_conf())
    assert abs(lp - math.log(p_global)) < 1e-8
    assert np.abs(c1.greens - c2.greens).max() < 1e-9


@pytest.mark.parametrize("kind,U", [(2, 4.0), (3, -4.0)])
def test_table_and_counter_rng_agree(kind, U):
    """The explicit [2M][2][N] uniform table reproduces the counter-RNG sweep (Metropolis + choice uniforms)."""
    T = OM.hopping_matrix("square", (4, 4))
    conf = rand_ghq(6, 16, 10)
    a = OR.RefChain(T, U=U, beta=1.0, field_kind=kind, conf=conf, seed=21, chain_id=3); a.init()
    b = OR.RefChain(T, U=U, beta=1.0, field_kind=kind, conf=conf, seed=21, chain_id=3); b.init()
    acc_a, pa, da = a.local_sweep(trace=True)
    acc_b, pb, db = b.local_sweep(uniforms=uniforms_for_sweep(21, 3, 0, 20, 16, ghq=True), trace=True)
    assert acc_a == acc_b and np.array_equal(da, db) and np.array_equal(a.get_conf(), b.get_conf())
    assert 0 < acc_a < 2 * 10 * 16
    assert set(np.unique(a.get_conf())) <= {1, 2, 3, 4}
