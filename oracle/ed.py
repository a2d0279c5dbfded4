"""oracle/ed.py -- exact diagonalisation of small Hubbard clusters from the operator definitions.

TEST INFRASTRUCTURE ONLY.  Restates the Hamiltonian of the reference's ED arbiter (test/ED/ED.jl:68-113):
    H = sum_{ij sigma} T_ij c^dag_{i sigma} c_{j sigma}  -  U sum_i (n_i up - 1/2)(n_i dn - 1/2)
with T the model's hopping matrix (-t on bonds, -mu on the diagonal; positive U is attractive), as dense
matrices on the 4^N-dimensional Fock space (Jordan-Wigner).  Used to pin the Wick kernels (U = 0, deterministic)
and, statistically, the whole sweep + measurement path (test/ED/ED_tests.jl:402-560).
"""
import numpy as np


def fermion_ops(nmodes):
    """Jordan-Wigner annihilation operators c_0 .. c_{nmodes-1} as dense matrices."""
    a = np.array([[0.0, 1.0], [0.0, 0.0]])
    Z = np.diag([1.0, -1.0])
    I2 = np.eye(2)
    ops = []
    for m in range(nmodes):
        mats = [Z] * m + [a] + [I2] * (nmodes - m - 1)
        out = mats[0]
        for x in mats[1:]:
            out = np.kron(out, x)
        ops.append(out)
    return ops


class HubbardED:
    def __init__(self, T, beta, U=0.0):
        self.N = T.shape[0]
        N = self.N
        self.c = fermion_ops(2 * N)                       # mode = site + N * spin
        dim = self.c[0].shape[0]
        H = np.zeros((dim, dim))
        for s in range(2):
            for i in range(N):
                for j in range(N):
                    if T[i, j] != 0.0:
                        H += T[i, j] * self.c[i + N * s].T @ self.c[j + N * s]
        if U != 0.0:
            Id = np.eye(dim)
            for i in range(N):
                H -= U * (self.n(i, 0) - 0.5 * Id) @ (self.n(i, 1) - 0.5 * Id)
        self.H, self.T, self.U = H, T, U
        self.w, self.V = np.linalg.eigh(H)
        self.beta = beta
        self.rho = (self.V * np.exp(-beta * (self.w - self.w.min()))) @ self.V.T
        self.rho /= np.trace(self.rho)

    def n(self, i, s):
        return self.c[i + self.N * s].T @ self.c[i + self.N * s]

    def expect(self, O):
        return np.trace(self.rho @ O)

    def evolve(self, O, tau):
        ep = (self.V * np.exp(tau * self.w)) @ self.V.T
        em = (self.V * np.exp(-tau * self.w)) @ self.V.T
        return ep @ O @ em

    def corr(self, A, B, tau=0.0):
        """<A(tau) B(0)>"""
        return np.trace(self.rho @ self.evolve(A, tau) @ B) if tau != 0.0 else np.trace(self.rho @ A @ B)

    def spin_ops(self, i):
        N = self.N
        up, dn = self.c[i], self.c[i + N]
        mx = up.T @ dn + dn.T @ up
        my = -1j * (up.T @ dn - dn.T @ up)
        mz = up.T @ up - dn.T @ dn
        return mx, my, mz

    def density(self, i):
        return self.n(i, 0) + self.n(i, 1)

    def kinetic(self):
        N = self.N
        return sum(self.T[i, j] * self.c[i + N * s].T @ self.c[j + N * s]
                   for s in range(2) for i in range(N) for j in range(N) if self.T[i, j] != 0.0)

    def interaction(self):
        Id = np.eye(self.c[0].shape[0])
        return -self.U * sum((self.n(i, 0) - 0.5 * Id) @ (self.n(i, 1) - 0.5 * Id) for i in range(self.N))

    def pair_by_distance(self, ops, s2d, tau=0.0):
        """sum over (src, trg) with direction d of <O_src(tau) O_trg(0)> / N  (EachSitePairByDistance + finalize)."""
        out = np.zeros(s2d.shape[0], dtype=complex)
        for src in range(self.N):
            for trg in range(self.N):
                out[s2d[src, trg]] += self.corr(ops[src], ops[trg], tau)
        return (out / self.N).real
