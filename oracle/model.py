"""oracle/model.py -- numpy restatement of the *inputs* of the DQMC sweep path.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Restates, independently of the
product package, how the reference builds what crosses the C-ABI boundary:

* lattice site ordering and directed bond lists
  (src/lattices/constructors.jl:12-26,46-58, src/lattices/lattice.jl:327-374),
* the Hubbard hopping matrix (src/models/HubbardModel.jl:112-124),
* the four hopping exponentials (src/flavors/DQMC/stack.jl:235-239),
* slice arithmetic (src/flavors/DQMC/parameters.jl:98-125),
* safe-multiplication ranges (src/flavors/DQMC/stack.jl:154-158),
* the Hirsch coupling alpha (src/flavors/DQMC/fields.jl:370-376, 419-425).

Pinned by tests/test_oracle_model.py against the reference's golden bond lists
(test/lattices.jl:80-93,169-182) and parameter tests (test/flavortests_DQMC.jl:4-18,244-262).
"""
from __future__ import annotations

import math

import numpy as np

# unit cells: (n_basis, [(from, to, shift), ...]) in the reference's bond order
UNIT_CELLS = {
    # constructors.jl:1-10
    "chain": (1, [(1, 1, (1,)), (1, 1, (-1,))]),
    # constructors.jl:12-26
    "square": (1, [(1, 1, (1, 0)), (1, 1, (0, 1)), (1, 1, (-1, 0)), (1, 1, (0, -1))]),
    # constructors.jl:46-58
    "honeycomb": (2, [(1, 2, (0, 0)), (1, 2, (-1, 0)), (1, 2, (0, -1)),
                      (2, 1, (0, 0)), (2, 1, (1, 0)), (2, 1, (0, 1))]),
    # constructors.jl:60-72
    "triangular": (1, [(1, 1, (1, 0)), (1, 1, (0, 1)), (1, 1, (-1, 1)),
                       (1, 1, (-1, 0)), (1, 1, (0, -1)), (1, 1, (1, -1))]),
}


def directed_bonds(kind: str, Ls: tuple[int, ...]):
    """All directed bonds (from, to), 1-based, in `bonds(l, Val(true))` order.

    lattice.jl:366-374 iterates Bravais cells (x fastest) and, per cell, the unit
    cell's bonds; `_shift_Bravais` (lattice.jl:327-357) maps (cell, bond) to flat
    indices with site = cell + (basis-1) * prod(Ls).
    """
    nbasis, ucbonds = UNIT_CELLS[kind]
    ncell = int(np.prod(Ls))
    out = []
    for idx in range(1, ncell + 1):
        for (bf, bt, shift) in ucbonds:
            flat_out, flat_fld, f = 1, idx, 1
            for d, L in enumerate(Ls):
                # fldmod1(flat_fld, L)
                t = (flat_fld - 1) % L + 1
                flat_fld = (flat_fld - 1) // L + 1
                t = (t + shift[d] - 1) % L + 1  # mod1
                flat_out += f * (t - 1)
                f *= L
            out.append((idx + (bf - 1) * f, flat_out + (bt - 1) * f))
    return out


def n_sites(kind: str, Ls: tuple[int, ...]) -> int:
    return UNIT_CELLS[kind][0] * int(np.prod(Ls))


def hopping_matrix(kind: str, Ls: tuple[int, ...], t: float = 1.0, mu: float = 0.0) -> np.ndarray:
    """HubbardModel.jl:112-124: T = diagm(-mu); T[to, from] += -t over directed bonds."""
    N = n_sites(kind, Ls)
    T = np.zeros((N, N))
    T[np.arange(N), np.arange(N)] = -mu
    for (frm, to) in directed_bonds(kind, Ls):
        T[to - 1, frm - 1] += -t
    return T


def sym_expm(A: np.ndarray) -> np.ndarray:
    """exp of a real symmetric matrix via eigh (Julia's exp(::Hermitian) path, real.jl:228-235)."""
    w, V = np.linalg.eigh(0.5 * (A + A.T))
    return (V * np.exp(w)) @ V.T


def hopping_exponentials(T: np.ndarray, delta_tau: float):
    """stack.jl:235-239 -> (eT2, eT2inv, eThalf, eThalfinv), Fortran-ordered."""
    f = np.asfortranarray
    return (f(sym_expm(-delta_tau * T)), f(sym_expm(+delta_tau * T)),
            f(sym_expm(-0.5 * delta_tau * T)), f(sym_expm(+0.5 * delta_tau * T)))


def julia_round(x: float) -> int:
    """Julia's round(Int, x): round-half-to-even (Python's round has the same rule)."""
    return int(round(x))


def n_slices(beta: float, delta_tau: float = 0.1) -> int:
    """parameters.jl:98,118: slices = round(beta / delta_tau)."""
    return julia_round(beta / delta_tau)


def generate_chunks(length: int, max_chunk_size: int):
    """stack.jl:154-158 -> list of (first, last), 1-based inclusive."""
    n_chunks = -(-length // max_chunk_size)
    step = length / n_chunks
    return [(julia_round((i - 1) * step) + 1, julia_round(i * step)) for i in range(1, n_chunks + 1)]


def hirsch_alpha(U: float, delta_tau: float, field_kind: int) -> float:
    """fields.jl:372 (density, kind 0): acosh(exp(+dt U / 2)); :421 (magnetic, kind 1): acosh(exp(-dt U / 2))."""
    s = 0.5 if field_kind == 0 else -0.5
    return math.acosh(math.exp(s * delta_tau * U))


# field kinds of the ABI (include/dqmc_b200.h): 0 DensityHirsch, 1 MagneticHirsch, 2 DensityGHQ, 3 MagneticGHQ
FIELD_KINDS = {"DensityHirschField": 0, "MagneticHirschField": 1, "DensityGHQField": 2, "MagneticGHQField": 3}


def ghq_tables():
    """fields.jl:517-519, 579-581 in the reference's own double arithmetic -> (eta[4], gamma[4], choices[4][3])."""
    s6 = math.sqrt(6.0)
    gam = np.array([1 - s6 / 3, 1 + s6 / 3, 1 + s6 / 3, 1 - s6 / 3])
    eta = np.array([-math.sqrt(6 + 2 * s6), -math.sqrt(6 - 2 * s6), math.sqrt(6 - 2 * s6), math.sqrt(6 + 2 * s6)])
    choices = np.array([[2, 3, 4], [1, 3, 4], [1, 2, 4], [1, 2, 3]], dtype=np.int8)
    return eta, gam, choices


def ghq_alpha(U: float, delta_tau: float, field_kind: int) -> float:
    """fields.jl:514 (magnetic GHQ, kind 3): sqrt(-dt U / 2); :576 (density GHQ, kind 2): sqrt(+dt U / 2).  Real only."""
    x = (0.5 if field_kind == 2 else -0.5) * delta_tau * U
    if x < 0:
        raise ValueError("complex GHQ coupling (DensityGHQ needs U > 0, MagneticGHQ U < 0): out of scope")
    return math.sqrt(x)


def field_alpha(U: float, delta_tau: float, field_kind: int) -> float:
    return hirsch_alpha(U, delta_tau, field_kind) if field_kind < 2 else ghq_alpha(U, delta_tau, field_kind)


def ghq_compress(conf) -> np.ndarray:
    """compress(::AbstractGHQField) (fields.jl:476-480): (1,2,3,4) -> bit pairs (00,01,10,11), high bit first; returns the
    UInt64 chunks of the BitArray."""
    v = np.asarray(conf, dtype=np.int64).ravel(order="F") - 1
    bits = np.empty(2 * v.size, dtype=np.uint8)
    bits[0::2] = v >> 1
    bits[1::2] = v & 1
    by = np.packbits(bits, bitorder="little")
    by = np.concatenate([by, np.zeros((-len(by)) % 8, dtype=np.uint8)])
    return by.view("<u8")


def ghq_decompress(chunks, shape) -> np.ndarray:
    """decompress(::AbstractGHQField, c) (fields.jl:481-489): 1 + 2 bit1 + bit2."""
    n = int(np.prod(shape))
    bits = np.unpackbits(np.asarray(chunks, dtype="<u8").view(np.uint8), bitorder="little")[:2 * n]
    return (1 + 2 * bits[0::2] + bits[1::2]).astype(np.int8).reshape(shape, order="F")


def choose_field(U: float) -> int:
    """HubbardModel.jl:83: U < 0 -> MagneticHirschField (1) else DensityHirschField (0)."""
    return 1 if U < 0.0 else 0


def analytic_greens(T: np.ndarray, beta: float) -> np.ndarray:
    """test/testfunctions.jl:120-165: G = V diag(1 / (1 + exp(-beta eps))) V^-1 for U = 0."""
    w, V = np.linalg.eigh(T)
    return (V * (1.0 / (1.0 + np.exp(-beta * w)))) @ V.T
