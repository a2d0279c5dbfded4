"""Pins oracle/model.py against the reference's golden vectors (no GPU).

test/lattices.jl:80-93 (SquareLattice(3) directed bonds), :169-182 (Honeycomb(2)),
test/flavortests_DQMC.jl:4-18 (slice arithmetic), :244-262 (chunk invariants).
"""
import numpy as np

from oracle import model as M


def test_square3_directed_bonds_golden():
    b = M.directed_bonds("square", (3, 3))
    assert [x[0] for x in b] == [1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6,
                                 7, 7, 7, 7, 8, 8, 8, 8, 9, 9, 9, 9]
    assert [x[1] for x in b] == [2, 4, 3, 7, 3, 5, 1, 8, 1, 6, 2, 9, 5, 7, 6, 1, 6, 8, 4, 2, 4, 9, 5, 3,
                                 8, 1, 9, 4, 9, 2, 7, 5, 7, 3, 8, 6]


def test_honeycomb2_directed_bonds_golden():
    b = M.directed_bonds("honeycomb", (2, 2))
    assert [x[0] for x in b] == [1, 1, 1, 5, 5, 5, 2, 2, 2, 6, 6, 6, 3, 3, 3, 7, 7, 7, 4, 4, 4, 8, 8, 8]
    assert [x[1] for x in b] == [5, 6, 7, 1, 2, 3, 6, 5, 8, 2, 1, 4, 7, 8, 5, 3, 4, 1, 8, 7, 6, 4, 3, 2]


def test_hopping_matrix_is_symmetric_and_counts_bonds():
    for kind, Ls, z in (("square", (4, 4), 4), ("honeycomb", (3, 3), 3), ("chain", (8,), 2)):
        T = M.hopping_matrix(kind, Ls, t=1.0, mu=0.3)
        assert np.array_equal(T, T.T)
        assert np.allclose(np.diag(T), -0.3)
        assert np.allclose((T - np.diag(np.diag(T))).sum(axis=0), -z)
    # L=2 double counts (+x and -x neighbour coincide): HubbardModel.jl:118-120
    T = M.hopping_matrix("square", (2, 2))
    assert T[0, 1] == -2.0


def test_slice_arithmetic():
    assert M.n_slices(5.0) == 50
    assert M.n_slices(5.0, 0.01) == 500
    assert M.n_slices(16.0, 0.1) == 160


def test_generate_chunks_invariants():
    rng = np.random.default_rng(0)
    for _ in range(300):
        slices = int(rng.integers(1, 101)); cs = int(rng.integers(1, 13))
        ch = M.generate_chunks(slices, cs)
        assert ch[0][0] == 1 and ch[-1][1] == slices
        for (a, b), nxt in zip(ch, ch[1:] + [None]):
            assert a > 0 and b <= slices and 0 < b - a + 1 <= cs
            if nxt is not None:
                assert nxt[0] == b + 1
    # round-half-to-even case called out in SURVEY section 7
    assert M.generate_chunks(5, 3) == [(1, 2), (3, 5)]
    assert M.generate_chunks(160, 10)[3] == (31, 40)


def test_hopping_exponentials():
    T = M.hopping_matrix("square", (4, 4))
    e2, e2i, eh, ehi = M.hopping_exponentials(T, 0.1)
    I = np.eye(16)
    assert np.abs(e2 @ e2i - I).max() < 1e-14
    assert np.abs(eh @ eh - e2).max() < 1e-14
    assert np.abs(ehi @ ehi - e2i).max() < 1e-14


def test_alpha():
    import math
    assert math.isclose(math.cosh(M.hirsch_alpha(4.0, 0.1, 0)), math.exp(0.2))
    assert math.isclose(math.cosh(M.hirsch_alpha(-4.0, 0.1, 1)), math.exp(0.2))
    assert M.choose_field(-1.0) == 1 and M.choose_field(1.0) == 0
