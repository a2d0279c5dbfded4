"""oracle/rng.py -- numpy restatement of include/dqmc_rng.h (Philox4x32-10 -> 53-bit uniform).

TEST INFRASTRUCTURE ONLY.  Used to build explicit uniform tables that must make the
table-driven and counter-driven sweeps take identical decisions.
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox_uniform(seed, chain, sweep, step, site, second=False):
    """Vectorised dqmc_uniform(); step/site may be arrays (broadcast)."""
    step = np.asarray(step, dtype=np.uint64)
    site = np.asarray(site, dtype=np.uint64)
    shape = np.broadcast(step, site).shape
    c0 = np.full(shape, np.uint64(chain & 0xFFFFFFFF))
    c1 = np.full(shape, np.uint64(sweep & 0xFFFFFFFF))
    c2 = np.broadcast_to(step, shape).copy()
    c3 = np.broadcast_to(site, shape).copy()
    k0 = ((seed & 0xFFFFFFFF) ^ ((chain >> 32) & 0xFFFFFFFF)) & 0xFFFFFFFF
    k1 = (((seed >> 32) & 0xFFFFFFFF) ^ ((sweep >> 32) & 0xFFFFFFFF)) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & _MASK
        n1 = p1 & _MASK
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & _MASK
        n3 = p0 & _MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    if second:
        bits = ((c2 << np.uint64(32)) | c3) >> np.uint64(11)      # dqmc_uniform_choice: the other two output words
    else:
        bits = ((c0 << np.uint64(32)) | c1) >> np.uint64(11)
    return bits.astype(np.float64) * (1.0 / 9007199254740992.0)


def uniforms_for_sweep(seed, chain, sweep, nsteps, nsites, ghq=False):
    """Table of the uniforms the library uses for one sweep of one chain: [nsteps, nsites] (Hirsch fields) or
    [nsteps, 2, nsites] (GHQ fields: Metropolis uniforms, then choice uniforms)."""
    st = np.arange(nsteps, dtype=np.uint64)[:, None]
    si = np.arange(nsites, dtype=np.uint64)[None, :]
    u = np.ascontiguousarray(philox_uniform(seed, chain, sweep, st, si))
    if not ghq:
        return u
    return np.ascontiguousarray(np.stack([u, philox_uniform(seed, chain, sweep, st, si, second=True)], axis=1))


def ghq_choice(x_old, u):
    """dqmc_ghq_choice (include/dqmc_rng.h)."""
    r = min(2, max(0, int(3.0 * u)))
    c = r + 1
    return c + 1 if c >= x_old else c
