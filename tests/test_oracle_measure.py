"""Pins oracle/measure.py (Wick kernels, EachSitePairByDistance sums, TimeIntegral weights) against
definition-level arbiters.  No GPU.

1. Exact diagonalisation of free spinful fermions on small lattices, evaluated from the operator
   definitions <O_i(tau) O_j(0)> -- the reference pins the same kernels against its ED code
   (test/ED/ED_tests.jl:186-330); at U = 0 DQMC is exact, so the comparison is deterministic.
2. The reference's general `Matrix` (2N x 2N) kernels, restated per element, against the BlockDiagonal /
   DiagonallyRepeatingMatrix specialisations on random matrices.
3. Lattice index maps: Bravais srctrg2dir (lattice_cache.jl:224-240).
"""
import numpy as np
import pytest

from oracle import measure as OMS
from oracle import model as M
from oracle import ref as R


# ------------------------------------------------------------------------------------------------ ED
from oracle.ed import HubbardED as FreeED  # noqa: E402  (U = 0 here)


def free_chain(kind, Ls, beta, field_kind, safe_mult=5):
    T = M.hopping_matrix(kind, Ls)
    c = R.RefChain(T, U=0.0, beta=beta, safe_mult=safe_mult, field_kind=field_kind)
    g = np.random.default_rng(1)
    c.set_conf(np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(c.N, c.M))))
    c.init()
    return T, c


@pytest.mark.parametrize("kind,Ls,field_kind", [("square", (2, 2), 0), ("square", (2, 2), 1), ("chain", (3,), 0)])
def test_kernels_against_exact_diagonalisation(kind, Ls, field_kind):
    beta = 1.0
    T, c = free_chain(kind, Ls, beta, field_kind)
    ed = FreeED(T, beta)
    N = c.N
    G00 = c.measured_greens()
    triples = list(c.combined_greens_iterator(recalculate=c.safe_mult))
    assert len(triples) == c.M + 1
    nop = [ed.n(i, 0) + ed.n(i, 1) for i in range(N)]
    sop = [ed.spin_ops(i) for i in range(N)]
    for (l, G0l, Gl0, Gll) in triples[::3] + [triples[-1]]:
        tau = l * c.delta_tau
        Kc = OMS.full_cdc(G00, G0l, Gl0, Gll, l)
        Kx = OMS.full_sdc_x(G00, G0l, Gl0, Gll, l)
        Ky = OMS.full_sdc_y(G00, G0l, Gl0, Gll, l)
        Kz = OMS.full_sdc_z(G00, G0l, Gl0, Gll, l)
        for i in range(N):
            for j in range(N):
                assert abs(Kc[i, j] - ed.corr(nop[i], nop[j], tau)) < 1e-9, ("cdc", l, i, j)
                assert abs(Kx[i, j] - ed.corr(sop[i][0], sop[j][0], tau)) < 1e-9, ("sdc x", l, i, j)
                assert abs(Ky[i, j] - ed.corr(sop[i][1], sop[j][1], tau)) < 1e-9, ("sdc y", l, i, j)
                assert abs(Kz[i, j] - ed.corr(sop[i][2], sop[j][2], tau)) < 1e-9, ("sdc z", l, i, j)
    # equal-time scalars and vectors
    et = OMS.equal_time(G00, T, 0.0, OMS.bravais_srctrg2dir(Ls), 1)
    for i in range(N):
        for s in range(c.nb):
            assert abs(et["occ"][i + N * s] - np.trace(ed.rho @ ed.n(i, s))) < 1e-12
    Hkin = sum(T[i, j] * ed.c[i + N * s].T @ ed.c[j + N * s] for s in range(2) for i in range(N) for j in range(N))
    assert abs(et["K"] - np.trace(ed.rho @ Hkin)) < 1e-10


def test_time_integral_and_pair_sums_by_definition():
    """apply!(::TimeIntegral) + EachSitePairByDistance, literally looped, from ED correlators."""
    Ls, beta = (2, 2), 1.0
    T, c = free_chain("square", Ls, beta, 0)
    ed = FreeED(T, beta)
    N = c.N
    s2d = OMS.bravais_srctrg2dir(Ls)
    G00 = c.measured_greens()
    got = OMS.time_integral(G00, c.combined_greens_iterator(recalculate=c.safe_mult), c.delta_tau, c.M, s2d, 1)
    nop = [ed.n(i, 0) + ed.n(i, 1) for i in range(N)]
    mz = [ed.spin_ops(i)[2] for i in range(N)]
    want_c = np.zeros(N); want_z = np.zeros(N)
    for l in range(c.M + 1):
        w = (0.5 if l in (0, c.M) else 1.0) * c.delta_tau
        for trg in range(N):
            for src in range(N):
                d = s2d[src, trg]
                want_c[d] += w * ed.corr(nop[src], nop[trg], l * c.delta_tau)
                want_z[d] += w * ed.corr(mz[src], mz[trg], l * c.delta_tau)
    assert np.allclose(got["cds"][:, 0, 0], want_c / N, atol=1e-9, rtol=0)
    assert np.allclose(got["sdzs"][:, 0, 0], want_z / N, atol=1e-9, rtol=0)


# ------------------------------------------------------------------------- Matrix kernels, per element
def _matrix_kernels(G00, G0l, Gl0, Gll, N, i, j, l):
    """The `_GM4{<: Matrix}` methods (charge_density.jl:68-86 summed over the flavor iterator,
    spin_density.jl:70-91, 119-141, 172-191) on 2N x 2N matrices, 0-based i, j."""
    ident = 1.0 if (i == j and l == 0) else 0.0
    cdc = 0.0
    for f1 in range(2):
        for f2 in range(2):
            s1, s2 = N * f1, N * f2
            idf = ident if f1 == f2 else 0.0
            cdc += (1 - Gll[i + s1, i + s1]) * (1 - G00[j + s2, j + s2]) + (idf - G0l[j + s1, i + s2]) * Gl0[i + s1, j + s2]
    sx = (Gll[i + N, i] * G00[j + N, j] + Gll[i + N, i] * G00[j, j + N] + Gll[i, i + N] * G00[j + N, j]
          + Gll[i, i + N] * G00[j, j + N]
          + (0 - G0l[j, i + N]) * Gl0[i + N, j] + (ident - G0l[j, i]) * Gl0[i + N, j + N]
          + (ident - G0l[j + N, i + N]) * Gl0[i, j] + (0 - G0l[j + N, i]) * Gl0[i, j + N])
    sz = ((1 - Gll[i, i]) * (1 - G00[j, j]) - (1 - Gll[i, i]) * (1 - G00[j + N, j + N])
          - (1 - Gll[i + N, i + N]) * (1 - G00[j, j]) + (1 - Gll[i + N, i + N]) * (1 - G00[j + N, j + N])
          + (ident - G0l[j, i]) * Gl0[i, j] - (0 - G0l[j + N, i]) * Gl0[i, j + N]
          - (0 - G0l[j, i + N]) * Gl0[i + N, j] + (ident - G0l[j + N, i + N]) * Gl0[i + N, j + N])
    return cdc, sx, sz


@pytest.mark.parametrize("nb", [1, 2])
@pytest.mark.parametrize("l", [0, 3])
def test_specialised_kernels_equal_general_matrix_kernels(nb, l):
    g = np.random.default_rng(nb * 10 + l)
    N = 5
    blocks = [g.random((N, N, nb)) for _ in range(4)]

    def embed(B):
        out = np.zeros((2 * N, 2 * N))
        out[:N, :N] = B[:, :, 0]
        out[N:, N:] = B[:, :, nb - 1]
        return out

    full = [embed(B) for B in blocks]
    Kc = OMS.full_cdc(*blocks, l); Kx = OMS.full_sdc_x(*blocks, l); Kz = OMS.full_sdc_z(*blocks, l)
    for i in range(N):
        for j in range(N):
            c_, x_, z_ = _matrix_kernels(*full, N, i, j, l)
            assert abs(Kc[i, j] - c_) < 1e-13
            assert abs(Kx[i, j] - x_) < 1e-13
            assert abs(Kz[i, j] - z_) < 1e-13


# ------------------------------------------------------------------------- lattice maps
def test_bravais_srctrg2dir():
    """lattice_cache.jl:224-240: output[flat_shift][flat_src] = flat(mod1(src + shift, Ls))."""
    Ls = (3, 4)
    s2d = OMS.bravais_srctrg2dir(Ls)
    n = 12
    for src in range(n):
        assert sorted(s2d[src]) == list(range(n))          # every direction once per source
        sx, sy = src % 3, src // 3
        for shift in range(n):
            dx, dy = shift % 3, shift // 3
            trg = (sx + dx) % 3 + 3 * ((sy + dy) % 4)
            assert s2d[src, trg] == shift
    assert np.all(np.diag(s2d) == 0)                        # on-site is direction 1 (0-based 0)


def test_log_binner_levels_are_block_means():
    """LogBinner (BinningAnalysis.jl semantics): level l holds the statistics of the means over 2^l successive values."""
    from oracle.measure import LogBinner
    g = np.random.default_rng(0)
    x = g.normal(size=(37, 3))
    B = LogBinner(shape=(3,), levels=8)
    for v in x:
        B.push(v)
    for l in range(6):
        nb = 37 >> l
        blocks = x[:nb << l].reshape(nb, 1 << l, 3).mean(axis=1)
        assert B.count[l] == nb
        assert np.allclose(B.sum[l], blocks.sum(axis=0)) and np.allclose(B.sumsq[l], (blocks ** 2).sum(axis=0))
    assert B.count[6] == 0
