"""HubbardModel: the model-side inputs of the sweep path (hopping matrix, field choice).

Mirror of src/models/HubbardModel.jl:22-124.  Positive U is ATTRACTIVE, negative repulsive
(HubbardModel.jl:15-16); HubbardModelRepulsive(U=x) flips the sign (:49-53).
"""
from __future__ import annotations

import numpy as np

from .lattices import Chain, Lattice, SquareLattice


class HubbardModel:
    def __init__(self, l: Lattice | None = None, *, L=2, dims=2, U=1.0, mu=0.0, t=1.0):
        if l is None:
            if dims == 1:
                l = Chain(L)
            elif dims == 2:
                l = SquareLattice(L)
            else:
                raise NotImplementedError("CubicLattice is outside the sweep path's configs")
        self.l, self.U, self.mu, self.t = l, float(U), float(mu), float(t)

    def __repr__(self):
        kind = "repulsive" if self.U < 0 else "attractive"
        return f"{kind} Hubbard model (t={self.t}, mu={self.mu}, U={self.U}, {len(self.l)} sites)"


def HubbardModelAttractive(*a, **kw):
    return HubbardModel(*a, **kw)


def HubbardModelRepulsive(*a, **kw):
    kw["U"] = -kw.get("U", 1.0)
    return HubbardModel(*a, **kw)


def lattice(m: HubbardModel) -> Lattice:
    return m.l


def hopping_matrix(m: HubbardModel) -> np.ndarray:
    """T = diagm(-mu); T[to, from] += -t over directed bonds (HubbardModel.jl:112-124)."""
    N = len(m.l)
    T = np.zeros((N, N), order="F")
    T[np.arange(N), np.arange(N)] = -m.mu
    for b in m.l.bonds(directed=True):
        T[b.to - 1, b.frm - 1] += -m.t
    return T


def choose_field(m: HubbardModel) -> str:
    """HubbardModel.jl:83"""
    return "MagneticHirschField" if m.U < 0.0 else "DensityHirschField"
