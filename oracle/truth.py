"""oracle/truth.py -- ctypes front-end of the extended-precision arbiter (oracle/truth_ld.c).

TEST INFRASTRUCTURE ONLY.  `greens_truth` evaluates G(slice) = [1 + B_slice ... B_1 B_M ... B_{slice+1}]^-1 for one
configuration in x87 long double with an independent stabilisation; the double-precision oracle and the CUDA library
are both measured against it (tests/test_gpu_parity_configs.py, tests/test_oracle_truth.py, bench.py's cpu_baseline leg).
"""
from __future__ import annotations

import ctypes as C
import math
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so, src = _HERE / "libdqmc_truth.so", _HERE / "truth_ld.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-B", "libdqmc_truth.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        dp = C.POINTER(C.c_double)
        L.truth_greens_ld.restype = C.c_int
        L.truth_greens_ld.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp]
        _LIB = L
    return _LIB


def slice_diagonals(conf, alpha: float, field_kind: int, block: int) -> np.ndarray:
    """exp(V_l) diagonals [M, N] of flavor block `block` for a Hirsch configuration conf[site, slice]
    (fields.jl:380-386 density: exp(alpha x); :429-438 magnetic: block 2 uses -alpha)."""
    sign = -1.0 if ((field_kind & 1) and block == 1) else 1.0
    if field_kind >= 2:                                    # GHQ: exp(+-alpha eta(x)), fields.jl:533-546, 596-602
        from .model import ghq_tables
        eta = ghq_tables()[0]
        lut = np.array([math.exp(sign * alpha * e) for e in eta])
        return np.ascontiguousarray(lut[np.asarray(conf, dtype=np.int64).T - 1], dtype=np.float64)
    ep, em = math.exp(sign * alpha), math.exp(-sign * alpha)
    return np.ascontiguousarray(np.where(np.asarray(conf).T > 0, ep, em), dtype=np.float64)


def greens_truth(eT2, ev, slice0: int = 0, chunk: int = 5) -> np.ndarray:
    """G(slice0) from the slice diagonals ev[M, N] and the hopping exponential eT2 (both taken as exact)."""
    eT2 = np.asfortranarray(eT2, dtype=np.float64)
    ev = np.ascontiguousarray(ev, dtype=np.float64)
    M, n = ev.shape
    G = np.zeros((n, n), order="F")
    dp = C.POINTER(C.c_double)
    rc = lib().truth_greens_ld(n, M, int(chunk), int(slice0), eT2.ctypes.data_as(dp), ev.ctypes.data_as(dp),
                               G.ctypes.data_as(dp))
    if rc != 0:
        raise RuntimeError(f"truth_greens_ld failed ({rc})")
    return G


def greens_truth_chain(chain, conf=None, slice0: int = 0, chunk: int = 5) -> np.ndarray:
    """All flavor blocks (N, N, nb) for an oracle RefChain-like object (eT2, alpha, kind, nb) and conf[site, slice]."""
    conf = chain.get_conf() if conf is None else conf
    return np.stack([greens_truth(chain.eT2, slice_diagonals(conf, chain.alpha, chain.kind, b), slice0, chunk)
                     for b in range(chain.nb)], axis=2)
