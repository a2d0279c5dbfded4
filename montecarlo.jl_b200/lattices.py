"""Lattices: the input contract of `hopping_matrix` (site ordering + directed bond lists).

Host-side mirror of the reference's src/lattices/lattice.jl:38-167, 327-374 and
src/lattices/constructors.jl:1-72 -- only what the DQMC sweep path consumes.
Site index (1-based) = x + Lx*(y-1) + ... + prod(Ls)*(basis-1).
"""
from __future__ import annotations

from dataclasses import dataclass
from math import prod, sqrt


@dataclass(frozen=True)
class Bond:
    """Bond(from, to, uc_shift, label) like src/lattices/lattice.jl Bond{N}; from/to are 1-based."""
    frm: int
    to: int
    uc_shift: tuple = ()
    label: int = 1


@dataclass(frozen=True)
class UnitCell:
    name: str
    lattice_vectors: tuple
    sites: tuple
    bonds: tuple          # Bond(from_basis, to_basis, uc_shift)


class Lattice:
    def __init__(self, unitcell: UnitCell, Ls):
        self.unitcell = unitcell
        self.Ls = tuple(int(L) for L in Ls)

    def __len__(self):
        return len(self.unitcell.sites) * prod(self.Ls)

    @property
    def size(self):
        return self.Ls

    def _shift_bravais(self, flat: int, b: Bond) -> Bond:
        flat_out, flat_fld, f = 1, flat, 1
        for d, L in enumerate(self.Ls):
            t = (flat_fld - 1) % L + 1
            flat_fld = (flat_fld - 1) // L + 1
            t = (t + b.uc_shift[d] - 1) % L + 1
            flat_out += f * (t - 1)
            f *= L
        return Bond(flat + (b.frm - 1) * f, flat_out + (b.to - 1) * f, b.uc_shift, b.label)

    def bravais_srctrg2dir(self):
        """l[:Bravais_srctrg2dir] (lattices/lattice_cache.jl:69-78, 224-240), 0-based: the direction of the
        pair (src, trg) of Bravais cells is the flat index of mod(trg - src, Ls), x fastest."""
        n = prod(self.Ls)
        out = [[0] * n for _ in range(n)]
        for src in range(n):
            for trg in range(n):
                d, f, a, b = 0, 1, src, trg
                for L in self.Ls:
                    d += f * (((b % L) - (a % L)) % L)
                    a //= L; b //= L; f *= L
                out[src][trg] = d
        return out

    def bonds(self, directed: bool = False):
        """bonds(l, Val(directed)): Bravais cell major, unit-cell bond minor."""
        ucb = self.unitcell.bonds
        if not directed:
            # the reference keeps one representative per undirected pair (unitcell._directed_indices)
            keep, seen = [], set()
            for j, b in enumerate(ucb):
                rev = (b.to, b.frm, tuple(-s for s in b.uc_shift))
                if (b.frm, b.to, tuple(b.uc_shift)) in seen:
                    continue
                seen.add(rev)
                keep.append(j)
            ucb = [ucb[j] for j in keep]
        return [self._shift_bravais(idx, b) for idx in range(1, prod(self.Ls) + 1) for b in ucb]


def Chain(Lx):
    uc = UnitCell("Chain", ((1.0,),), ((0.0,),), (Bond(1, 1, (1,)), Bond(1, 1, (-1,))))
    return Lattice(uc, (Lx,))


def SquareLattice(Lx, Ly=None):
    Ly = Lx if Ly is None else Ly
    uc = UnitCell("Square", ((1.0, 0.0), (0.0, 1.0)), ((0.0, 0.0),),
                  (Bond(1, 1, (1, 0)), Bond(1, 1, (0, 1)), Bond(1, 1, (-1, 0)), Bond(1, 1, (0, -1))))
    return Lattice(uc, (Lx, Ly))


def Honeycomb(Lx, Ly=None):
    Ly = Lx if Ly is None else Ly
    uc = UnitCell("Honeycomb", ((sqrt(3.0) / 2, -0.5), (sqrt(3.0) / 2, 0.5)),
                  ((0.0, 0.0), (1 / sqrt(3.0), 0.0)),
                  (Bond(1, 2, (0, 0)), Bond(1, 2, (-1, 0)), Bond(1, 2, (0, -1)),
                   Bond(2, 1, (0, 0)), Bond(2, 1, (1, 0)), Bond(2, 1, (0, 1))))
    return Lattice(uc, (Lx, Ly))


def TriangularLattice(Lx, Ly=None):
    Ly = Lx if Ly is None else Ly
    uc = UnitCell("Triangular", ((sqrt(3.0) / 2, -0.5), (sqrt(3.0) / 2, 0.5)), ((0.0, 0.0),),
                  (Bond(1, 1, (1, 0)), Bond(1, 1, (0, 1)), Bond(1, 1, (-1, 1)),
                   Bond(1, 1, (-1, 0)), Bond(1, 1, (0, -1)), Bond(1, 1, (1, -1))))
    return Lattice(uc, (Lx, Ly))
