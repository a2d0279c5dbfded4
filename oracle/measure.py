"""oracle/measure.py -- numpy restatement of the reference's Wick kernels and lattice sums for the
equal-time and time-integrated observables the device evaluates (montecarlo.jl_b200/csrc/measure.cu).

TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Restates (paths relative to /root/reference/src):
  lattices/lattice_cache.jl:69-78, 224-240           Bravais srctrg2dir (dir = flat index of the shift)
  flavors/DQMC/measurements/generic.jl:337-372        apply!(::TimeIntegral): weights 0.5 dtau at l = 0, M
  flavors/DQMC/measurements/generic.jl:434-461        apply!(temp, ::EachSitePairByDistance, ...)
  flavors/DQMC/measurements/generic.jl:578-583        finalize_temp!: temp ./= length(lattice)
  flavors/DQMC/measurements/constructors/charge_density.jl:62-110   full_cdc_kernel
  flavors/DQMC/measurements/constructors/spin_density.jl:66-218     full_sdc_{x,y,z}_kernel
  flavors/DQMC/measurements/constructors/occupation.jl:44-70        occupation
  flavors/DQMC/measurements/constructors/energy.jl:119-165, models/HubbardModel.jl:160-183  energies
  flavors/DQMC/measurements/constructors/main.jl:5-36  FlavorIterator

Green's functions are (N, N, nb) arrays of *measured* G (after the exp(+-dtau T/2) transform); nb = 1 is
the reference's DiagonallyRepeatingMatrix (both spins equal), nb = 2 its BlockDiagonal.

Parity status: PINNED by tests/test_oracle_measure.py against exact diagonalisation of the 2x2 (and
3-site-chain) free-fermion problem evaluated from the operator definitions (the reference pins the same
kernels against its ED code, test/ED/ED_tests.jl:186-330).
"""
from __future__ import annotations

import numpy as np


def bravais_srctrg2dir(Ls):
    """construct_srctrg2dir(Bravais(l)): dir[src, trg] (0-based) = flat index of mod(trg - src, Ls), x fastest."""
    Ls = tuple(int(L) for L in Ls)
    n = int(np.prod(Ls))
    subs = np.array(np.unravel_index(np.arange(n), Ls, order="F")).T      # (n, ndim), x fastest
    out = np.zeros((n, n), dtype=np.int32)
    for s in range(n):
        shift = (subs - subs[s]) % np.array(Ls)
        out[s, :] = np.ravel_multi_index(tuple(shift.T), Ls, order="F")
    return out


# ----------------------------------------------------------------------------------------------- kernels
# Every function returns K[i, j] for all site pairs at once; `l` is G0l.l (id = delta_ij * delta_{0 l}).
def _id(N, l):
    return np.eye(N) if l == 0 else np.zeros((N, N))


def full_cdc(G00, G0l, Gl0, Gll, l):
    """charge_density.jl:62-110 summed over FlavorIterator(mc, 2)."""
    N, _, nb = G00.shape
    ident = _id(N, l)
    if nb == 1:       # DiagonallyRepeatingMatrix, flv = 2, one flavor iteration
        a = 1.0 - np.diag(Gll[:, :, 0]); b = 1.0 - np.diag(G00[:, :, 0])
        return 4.0 * np.outer(a, b) + 2.0 * (ident - G0l[:, :, 0].T) * Gl0[:, :, 0]
    out = np.zeros((N, N))
    for f1 in range(2):
        for f2 in range(2):
            out += np.outer(1.0 - np.diag(Gll[:, :, f1]), 1.0 - np.diag(G00[:, :, f2]))
            if f1 == f2:
                out += (ident - G0l[:, :, f1].T) * Gl0[:, :, f1]
    return out


def full_sdc_x(G00, G0l, Gl0, Gll, l):
    """spin_density.jl:66-113 (x); the y kernel :116-166 has the same value for these matrix types."""
    N, _, nb = G00.shape
    ident = _id(N, l)
    if nb == 1:
        return 2.0 * (ident - G0l[:, :, 0].T) * Gl0[:, :, 0]
    return (ident - G0l[:, :, 0].T) * Gl0[:, :, 1] + (ident - G0l[:, :, 1].T) * Gl0[:, :, 0]


full_sdc_y = full_sdc_x


def full_sdc_z(G00, G0l, Gl0, Gll, l):
    """spin_density.jl:169-218."""
    N, _, nb = G00.shape
    ident = _id(N, l)
    if nb == 1:
        return 2.0 * (ident - G0l[:, :, 0].T) * Gl0[:, :, 0]
    a = [1.0 - np.diag(Gll[:, :, f]) for f in range(2)]
    b = [1.0 - np.diag(G00[:, :, f]) for f in range(2)]
    out = np.outer(a[0], b[0]) - np.outer(a[0], b[1]) - np.outer(a[1], b[0]) + np.outer(a[1], b[1])
    return out + (ident - G0l[:, :, 0].T) * Gl0[:, :, 0] + (ident - G0l[:, :, 1].T) * Gl0[:, :, 1]


PAIR_KERNELS = {"cd": full_cdc, "sdx": full_sdc_x, "sdy": full_sdc_y, "sdz": full_sdc_z}


def each_site_pair_by_distance(K, s2d, nbasis):
    """generic.jl:434-461 + finalize_temp! (:578-583): temp[dir, b1, b2] = sum K[src + uc1, trg + uc2] / N."""
    nbr = s2d.shape[0]
    N = nbr * nbasis
    out = np.zeros((nbr, nbasis, nbasis))
    for b2 in range(nbasis):
        for b1 in range(nbasis):
            blk = K[b1 * nbr:(b1 + 1) * nbr, b2 * nbr:(b2 + 1) * nbr]
            np.add.at(out[:, b1, b2], s2d.ravel(), blk.ravel())
    return out / N


def occupation(G):
    """occupation.jl:44-70: 1 - G[i, i], flavor-major."""
    return np.concatenate([1.0 - np.diag(G[:, :, f]) for f in range(G.shape[2])])


def kinetic_energy(G, T):
    """energy.jl:119-152 (T is symmetric and real here)."""
    N, _, nb = G.shape
    e = sum(np.sum(T * (np.eye(N) - G[:, :, f])) for f in range(nb))
    return 2.0 * e if nb == 1 else e


def interaction_energy(G, U):
    """HubbardModel.jl:160-183."""
    if G.shape[2] == 1:
        return -U * np.sum((np.diag(G[:, :, 0]) - 0.5) ** 2)
    return -U * np.sum((np.diag(G[:, :, 0]) - 0.5) * (np.diag(G[:, :, 1]) - 0.5))


def equal_time(G, T, U, s2d, nbasis):
    """All equal-time observables of one measured G: kernels evaluated on (G, G, G, G) with k = l = 0."""
    out = {"occ": occupation(G), "K": kinetic_energy(G, T), "V": interaction_energy(G, U)}
    out["E"] = out["K"] + out["V"]
    for name, fn in PAIR_KERNELS.items():
        out[name + "c"] = each_site_pair_by_distance(fn(G, G, G, G, 0), s2d, nbasis)
    return out


def time_integral(G00, triples, delta_tau, M, s2d, nbasis):
    """apply!(::TimeIntegral) (generic.jl:337-372): triples = iterable of (l, G0l, Gl0, Gll), l = 0 .. M."""
    out = {name + "s": np.zeros((s2d.shape[0], nbasis, nbasis)) for name in PAIR_KERNELS}
    for (l, G0l, Gl0, Gll) in triples:
        w = (0.5 if l in (0, M) else 1.0) * delta_tau
        for name, fn in PAIR_KERNELS.items():
            out[name + "s"] += w * each_site_pair_by_distance(fn(G00, G0l, Gl0, Gll, l), s2d, nbasis)
    return out


class LogBinner:
    """numpy restatement of BinningAnalysis.jl 0.6's LogBinner (`_push!`: every value that reaches a level is added to that
    level's {sum, sum of squares, count}; a compressor per level averages two successive values and hands the mean to the
    next level) -- the accumulator behind every DQMCMeasurement (measurements/generic.jl:62-65, 586-587).  Third-party
    arithmetic: pinned here only against its definition (level l = statistics of means over 2^l successive values)."""

    def __init__(self, shape=(), levels=20):
        self.levels = levels
        self.sum = np.zeros((levels,) + tuple(shape))
        self.sumsq = np.zeros((levels,) + tuple(shape))
        self.count = np.zeros(levels, dtype=np.int64)
        self.pending = [None] * levels

    def push(self, value):
        v = np.array(value, dtype=np.float64, copy=True)
        for l in range(self.levels):
            self.sum[l] += v
            self.sumsq[l] += v * v
            self.count[l] += 1
            if self.pending[l] is None or l == self.levels - 1:
                self.pending[l] = v
                return
            v = 0.5 * (self.pending[l] + v)
            self.pending[l] = None

    def std_error(self, level):
        n = self.count[level]
        mean = self.sum[level] / n
        return np.sqrt(np.maximum(self.sumsq[level] / n - mean ** 2, 0.0) / (n - 1))
