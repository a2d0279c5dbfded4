"""GPU parity of the 4-state Gauss-Hermite fields (DensityGHQField / MagneticGHQField, fields.jl:464-637) against the
oracle, through the C ABI: slice matrices, full sweeps under the shared counter RNG (both update kernels), the explicit
uniform table with its choice block, the 2-bit recorder wire format, global updates, the user-level API."""
import numpy as np
import pytest

from oracle import model as OM
from oracle.rng import uniforms_for_sweep

from test_gpu_parity import GTOL, check_sweeps, make_pair, relerr, rng

pytestmark = pytest.mark.gpu

KINDS = [(2, 4.0), (3, -4.0)]          # (field kind, U): real couplings only


@pytest.mark.parametrize("fk,U", KINDS)
def test_ghq_slice_matrices(b200, fk, U):
    """interaction_matrix_exp! of the GHQ fields inside the fused GEMMs (test/slice_matrices.jl:13-42)."""
    ctx, chains = make_pair(b200, "chain", (8,), U=U, beta=3.0, B=2, field_kind=fk)
    g = rng(3)
    X = np.asfortranarray(g.random((8, 8, ctx.nb, 2)))
    for which in ("left", "right", "inv_left", "inv_right", "daggered_left"):
        Y = ctx.multiply_slice_matrix(which, 17, X)
        for b, c in enumerate(chains):
            assert relerr(Y[:, :, :, b], c.multiply_slice_matrix(which, 17, X[:, :, :, b])) < 1e-13


@pytest.mark.parametrize("fk,U", KINDS)
@pytest.mark.parametrize("kind,Ls,beta", [("square", (4, 4), 2.0), ("square", (8, 8), 1.0), ("square", (10, 10), 0.5),
                                          ("honeycomb", (3, 3), 1.0)])
def test_ghq_sweep_parity(b200, fk, U, kind, Ls, beta):
    """Free-running sweeps: identical decisions (Metropolis AND choice draws), conf in 1..4, G <= 1e-10.
    4x4 / 8x8 / honeycomb run update.cu, 10x10 runs update3.cu."""
    ctx, chains = make_pair(b200, kind, Ls, U=U, beta=beta, B=2, field_kind=fk)
    check_sweeps(ctx, chains, 2)
    conf = ctx.get_conf()
    assert set(np.unique(conf)) <= {1, 2, 3, 4}
    assert (conf != make_pair(b200, kind, Ls, U=U, beta=beta, B=2, field_kind=fk)[0].get_conf()).any()


@pytest.mark.parametrize("version", [1, 3])
def test_ghq_uniform_table_equals_counter_rng(b200, version):
    """dqmc_sweep_traced with the explicit [B][2M][2][N] table == the counter RNG, on both update kernels."""
    out = []
    for use_table in (False, True):
        ctx, _ = make_pair(b200, "square", (6, 6), U=-4.0, beta=1.0, B=2, field_kind=3, update_variant=version)
        ctx.build_stack()
        u = np.stack([uniforms_for_sweep(11, b, 0, 2 * ctx.M, ctx.N, ghq=True) for b in range(2)]) if use_table else None
        acc, p, d = ctx.sweep_traced(uniforms=u)
        out.append((acc, d, ctx.get_conf(), ctx.greens()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2], out[1][2]) and np.array_equal(out[0][3], out[1][3])


def test_ghq_conf_packed_on_device(b200):
    """compress / decompress of the GHQ fields (fields.jl:476-489) on the device == the oracle's restatement."""
    ctx, chains = make_pair(b200, "square", (7, 7), U=-3.0, beta=1.3, B=3, field_kind=3)      # 2 x 637 bits
    conf = ctx.get_conf()
    packed = ctx.get_conf_packed()
    assert packed.shape == ((2 * 49 * 13 + 63) // 64, 3)
    for b in range(3):
        assert np.array_equal(packed[:, b], OM.ghq_compress(conf[:, :, b]))
    ctx.set_conf(np.asfortranarray(5 - conf))
    ctx.set_conf_packed(packed[:, 1:], chain0=1)
    now = ctx.get_conf()
    assert np.array_equal(now[:, :, 0], 5 - conf[:, :, 0]) and np.array_equal(now[:, :, 1:], conf[:, :, 1:])


@pytest.mark.parametrize("fk,U", [(3, -1.0), (2, 1.0)])
def test_ghq_global_shuffle(b200, fk, U):
    """global_update with a shuffled proposal (test/updates.jl:186-245): p and decisions == the oracle's; GlobalFlip is
    rejected for 4-state fields."""
    ctx, chains = make_pair(b200, "square", (2, 2), U=U, beta=2.0, B=3, mu=0.5, field_kind=fk, seed=5)
    ctx.build_stack()
    g = rng(8)
    old = ctx.get_conf()
    new = np.asfortranarray(np.stack([g.permutation(old[:, :, b].ravel()).reshape(4, 20) for b in range(3)], axis=2))
    u = np.array([0.0, 0.5, 0.999])
    acc, p = ctx.global_update(10, proposed=new, uniforms=u)
    G = ctx.greens()
    for b, c in enumerate(chains):
        c.init()
        a_ref, p_ref = c.global_update(new[:, :, b], uniform=u[b])
        assert acc[b] == a_ref and np.isclose(p[b], p_ref, rtol=1e-9)
        assert relerr(G[:, :, :, b], c.greens) < GTOL
    with pytest.raises(b200.DQMCError):
        ctx.global_update(10)
    with pytest.raises(b200.DQMCError):
        ctx.set_conf(np.zeros((4, 20, 3), dtype=np.int8))


def test_ghq_user_api(b200):
    """DQMC(model; field = MagneticGHQField) through run: U = 0 is not expressible (alpha = 0 gives the free G), so check
    against the analytic free Green's function at U -> 0 and that a finite-U run stays in 1..4."""
    model = b200.HubbardModel(b200.SquareLattice(4), U=0.0)
    mc = b200.DQMC(model, beta=1.0, safe_mult=5, thermalization=1, sweeps=2, measure_rate=1, seed=3, n_chains=2,
                   field="MagneticGHQField")
    mc["G"] = b200.greens_measurement(mc, model)
    assert b200.run(mc) == "SUCCESS"
    Gan = OM.analytic_greens(OM.hopping_matrix("square", (4, 4)), 1.0)
    assert np.allclose(mc["G"].mean(), np.stack([Gan, Gan], axis=2), atol=1e-12)
    model = b200.HubbardModel(b200.SquareLattice(4), U=4.0)
    mc = b200.DQMC(model, beta=1.0, thermalization=2, sweeps=2, seed=3, n_chains=2, field="DensityGHQField")
    assert b200.run(mc) == "SUCCESS"
    assert set(np.unique(mc.field.confs)) <= {1, 2, 3, 4} and mc.accepted.sum() > 0
    with pytest.raises(NotImplementedError):
        b200.DQMC(model, beta=1.0, field="MagneticGHQField")           # U > 0: complex coupling
