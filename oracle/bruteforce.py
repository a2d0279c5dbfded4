"""oracle/bruteforce.py -- definition-level (numpy / LAPACK) arbiters used to pin the oracle
and, for small systems, the GPU path.  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import scipy.linalg as sla


def field_values(c, conf):
    """The number the field couples with: x itself (Hirsch, +-1) or eta(x) (GHQ, x = 1..4; fields.jl:533-546, 596-602)."""
    if getattr(c, "kind", 0) >= 2:
        from .model import ghq_tables
        return ghq_tables()[0][np.asarray(conf, dtype=np.int64) - 1]
    return np.asarray(conf).astype(float)


def slice_B(c, conf, l, b=0):
    """B_l = eT2 * eV_l for flavor block b (stack.jl:319-327, fields.jl:380-386, 429-438)."""
    s = -1.0 if ((c.kind & 1) and b == 1) else 1.0
    return c.eT2 @ np.diag(np.exp(s * c.alpha * field_values(c, conf[:, l - 1])))


def decompose_udt(A):
    """test/linalg/old_linalg.jl:16-24 with LAPACK dgeqp3."""
    Q, Rm, p = sla.qr(A, pivoting=True)
    D = np.abs(np.diag(Rm))
    T = (Rm / D[:, None])[:, np.argsort(p)]
    return Q, D, T


def greens_lapack(c, conf, slice_, b=0, safe_mult=None):
    """test/testfunctions.jl:10-118 (calculate_greens_and_logdet) restated with scipy."""
    sm = safe_mult or c.safe_mult
    N = c.N

    def chain(ks, dagger):
        Uq, D, T = np.eye(N), np.ones(N), np.eye(N)
        for k in ks:
            B = slice_B(c, conf, k, b)
            Uq = (B.T if dagger else B) @ Uq
            if k % sm == 0:
                Uq, D, Tn = decompose_udt(Uq * D[None, :])
                T = Tn @ T
        Uq, D, Tn = decompose_udt(Uq * D[None, :])
        return Uq, D, Tn @ T

    if slice_ + 1 <= c.M:
        Ur, Dr, Tr = chain(range(c.M, slice_, -1), True)
    else:
        Ur, Dr, Tr = np.eye(N), np.ones(N), np.eye(N)
    if slice_ >= 1:
        Ul, Dl, Tl = chain(range(1, slice_ + 1), False)
    else:
        Ul, Dl, Tl = np.eye(N), np.ones(N), np.eye(N)
    Uq, D, T = decompose_udt(Dl[:, None] * (Tl @ Tr.T) * Dr[None, :])
    Uq = Ul @ Uq
    T = T @ Ur.T
    u, d, t = decompose_udt(Uq.T @ np.linalg.inv(T) + np.diag(D))
    T = np.linalg.inv(t @ T)
    Uq = (Uq @ u).T
    return T @ np.diag(1.0 / d) @ Uq


def greens_brute(c, conf, slice_, b=0):
    """G(slice) = [I + B_{slice-1}...B_1 B_M ... B_slice]^-1; only meaningful for small beta."""
    P = np.eye(c.N)
    for l in list(range(slice_, c.M + 1)) + list(range(1, slice_)):
        P = slice_B(c, conf, l, b) @ P
    return np.linalg.inv(np.eye(c.N) + P)


def log_weight(c, conf):
    """log W(conf) = log of [bosonic factor * prod_flavors det(I + B_M ... B_1)]."""
    lw = 0.0
    for b in range(c.nb):
        P = np.eye(c.N)
        for l in range(1, c.M + 1):
            P = slice_B(c, conf, l, b) @ P
        s, ld = np.linalg.slogdet(np.eye(c.N) + P)
        assert s > 0
        lw += ld * (2.0 if c.nb == 1 else 1.0)
    if (c.kind & 1) == 0:
        # exp(alpha x (n_up + n_dn - 1)): the "-1" is the bosonic weight exp(-alpha * sum(x)) (GHQ: x -> eta(x))
        lw += -c.alpha * float(field_values(c, conf).sum())
    if c.kind >= 2:
        # exp(dtau U A^2) ~ sum_x gamma(x) exp(alpha eta(x) A) (fields.jl:499-501): the weights gamma(x) of every site / slice
        from .model import ghq_tables
        lw += float(np.log(ghq_tables()[1][np.asarray(conf, dtype=np.int64) - 1]).sum())
    return lw
