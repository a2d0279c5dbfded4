"""oracle/ref.py -- ctypes front-end of the C oracle (oracle/dqmc_ref.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from . import model as M

_HERE = Path(__file__).resolve().parent
_LIB = None

c_double_p = C.POINTER(C.c_double)
c_i8_p = C.POINTER(C.c_int8)
c_u8_p = C.POINTER(C.c_uint8)
c_i64_p = C.POINTER(C.c_int64)
c_int_p = C.POINTER(C.c_int)


class RefStats(C.Structure):
    _fields_ = [("neg_count", C.c_int64), ("neg_sumlog", C.c_double), ("neg_min", C.c_double),
                ("neg_max", C.c_double), ("prop_count", C.c_int64), ("prop_sumlog", C.c_double),
                ("prop_min", C.c_double), ("prop_max", C.c_double)]


def build(force: bool = False) -> Path:
    so = _HERE / "libdqmc_ref.so"
    src = _HERE / "dqmc_ref.c"
    hdr = _HERE.parent / "include" / "dqmc_rng.h"
    deps = [src, _HERE / "dqmc_ref_ut.inc.c", _HERE / "dqmc_ref_global.inc.c", hdr]
    stale = (not so.exists()) or so.stat().st_mtime < max(d.stat().st_mtime for d in deps)
    # the library is compiled -march=native and travels with the tree: rebuild when the host CPU is not the one it
    # was built on (an illegal-instruction fault is the alternative)
    tag, cpu = _HERE / "libdqmc_ref.so.cpu", _cpu_signature()
    if not stale and (not tag.exists() or tag.read_text() != cpu):
        stale = True
    if force or stale:
        subprocess.check_call(["make", "-C", str(_HERE), "-B", "libdqmc_ref.so"],
                              stdout=subprocess.DEVNULL)
        tag.write_text(cpu)
    return so


def _cpu_signature() -> str:
    import hashlib
    try:
        lines = [l for l in open("/proc/cpuinfo") if l.startswith(("model name", "flags"))][:2]
    except OSError:
        lines = []
    return hashlib.sha1("".join(lines).encode()).hexdigest()


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        L.ref_chain_create.restype = C.c_void_p
        L.ref_chain_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p,
                                       C.c_double, c_double_p, c_double_p, c_double_p, c_double_p,
                                       C.c_int, C.c_int, C.c_uint64, C.c_int64]
        L.ref_chain_destroy.argtypes = [C.c_void_p]
        for name in ("ref_build_stack", "ref_reverse_build_stack", "ref_propagate"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.ref_calculate_greens_at.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p]
        L.ref_sweep_spatial.restype = C.c_int
        L.ref_sweep_spatial.argtypes = [C.c_void_p, c_u8_p, c_double_p, c_u8_p]
        L.ref_local_sweep.restype = C.c_int64
        L.ref_local_sweep.argtypes = [C.c_void_p, c_double_p, c_u8_p, c_double_p, c_u8_p]
        L.ref_measured_greens.argtypes = [C.c_void_p, c_double_p]
        L.ref_set_conf.argtypes = [C.c_void_p, c_i8_p]
        L.ref_get_conf.argtypes = [C.c_void_p, c_i8_p]
        L.ref_get_greens.argtypes = [C.c_void_p, c_double_p]
        L.ref_set_greens.argtypes = [C.c_void_p, c_double_p]
        L.ref_get_state.argtypes = [C.c_void_p, c_int_p]
        L.ref_set_state.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_get_stats.argtypes = [C.c_void_p, C.POINTER(RefStats)]
        L.ref_set_sweep_index.argtypes = [C.c_void_p, C.c_int64]
        L.ref_get_array.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p]
        L.ref_vmul.argtypes = [C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]
        L.ref_udt_pivot.argtypes = [C.c_int, c_double_p, c_double_p, c_double_p, c_i64_p,
                                    c_double_p, C.c_int]
        L.ref_rdivp.argtypes = [C.c_int, c_double_p, c_double_p, c_double_p, c_i64_p]
        L.ref_calculate_greens_block.argtypes = [C.c_int] + [c_double_p] * 7 + [c_i64_p, c_double_p]
        L.ref_multiply_slice_matrix.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p]
        L.ref_wrap_greens.argtypes = [C.c_void_p, c_double_p, C.c_int, C.c_int]
        L.ref_propose_local.restype = C.c_double
        L.ref_propose_local.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_propose_local_choice.restype = C.c_double
        L.ref_propose_local_choice.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.ref_max_threads.restype = C.c_int
        L.ref_run_chains.restype = C.c_double
        L.ref_run_chains.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, c_i64_p]
        # unequal-time path (dqmc_ref_ut.inc.c)
        L.ref_ut_build_stack.argtypes = [C.c_void_p]
        L.ref_ut_lazy_build.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_ut_get_array.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p]
        L.ref_ut_calculate_greens.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p]
        L.ref_ut_greens.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p]
        L.ref_find_range_with_value.restype = C.c_int
        L.ref_find_range_with_value.argtypes = [C.c_void_p, C.c_int]
        L.ref_cgi_first.restype = C.c_int
        L.ref_cgi_first.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]
        L.ref_cgi_next.restype = C.c_int
        L.ref_cgi_next.argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p]
        L.ref_global_update.restype = C.c_int
        L.ref_global_update.argtypes = [C.c_void_p, c_i8_p, C.c_double, C.c_int, c_double_p]
        _LIB = L
    return _LIB


def _dp(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)


def _f(a, shape=None) -> np.ndarray:
    a = np.array(a, dtype=np.float64, order="F", copy=True)
    return a if shape is None else a.reshape(shape, order="F")


# ---------------------------------------------------------------- operator level
def vmul(A, B, ta=False, tb=False):
    A, B = _f(A), _f(B)
    n = A.shape[0]
    Cm = np.zeros((n, n), order="F")
    lib().ref_vmul(n, int(ta), int(tb), _dp(Cm), _dp(A), _dp(B))
    return Cm


def udt_pivot(X, apply_pivot=True):
    """-> U, D, T, pivot (0-based).  UDT.jl:216-334."""
    T = _f(X)
    n = T.shape[0]
    U = np.zeros((n, n), order="F")
    D = np.zeros(n)
    piv = np.zeros(n, dtype=np.int64)
    tmp = np.zeros(n)
    lib().ref_udt_pivot(n, _dp(U), _dp(D), _dp(T), piv.ctypes.data_as(c_i64_p), _dp(tmp), int(apply_pivot))
    return U, D, T, piv


def rdivp(A, T, pivot):
    A, T = _f(A), _f(T)
    n = A.shape[0]
    O = np.zeros((n, n), order="F")
    piv = np.ascontiguousarray(pivot, dtype=np.int64)
    lib().ref_rdivp(n, _dp(A), _dp(T), _dp(O), piv.ctypes.data_as(c_i64_p))
    return A


def calculate_greens_udt(Ul, Dl, Tl, Ur, Dr, Tr):
    """stack.jl:442-496 on copies of the inputs -> G."""
    Ul, Tl, Ur, Tr = _f(Ul), _f(Tl), _f(Ur), _f(Tr)
    Dl, Dr = np.array(Dl, dtype=np.float64), np.array(Dr, dtype=np.float64)
    n = Ul.shape[0]
    G = np.zeros((n, n), order="F")
    piv = np.zeros(n, dtype=np.int64)
    tmp = np.zeros(n)
    lib().ref_calculate_greens_block(n, _dp(Ul), _dp(Dl), _dp(Tl), _dp(Ur), _dp(Dr), _dp(Tr), _dp(G),
                                     piv.ctypes.data_as(c_i64_p), _dp(tmp))
    return G


# ---------------------------------------------------------------- chain level
class RefChain:
    """One reference simulation: DQMCStack + Hirsch field (stack.jl, fields.jl)."""

    def __init__(self, T, *, U, beta=None, delta_tau=0.1, slices=None, safe_mult=10,
                 field_kind=None, check_sign_problem=True, check_propagation_error=True,
                 seed=1234, chain_id=0, conf=None):
        self.T = np.asfortranarray(T, dtype=np.float64)
        self.N = self.T.shape[0]
        self.delta_tau = float(delta_tau)
        self.M = int(slices) if slices is not None else M.n_slices(beta, delta_tau)
        self.beta = self.M * self.delta_tau
        self.safe_mult = int(safe_mult)
        self.kind = M.choose_field(U) if field_kind is None else int(field_kind)
        self.nb = 2 if (self.kind & 1) else 1
        self.ghq = self.kind >= 2
        self.U = float(U)
        self.alpha = M.field_alpha(U, delta_tau, self.kind)
        self.ranges = M.generate_chunks(self.M, self.safe_mult)
        self.C = len(self.ranges)
        self.eT2, self.eT2inv, self.eThalf, self.eThalfinv = M.hopping_exponentials(self.T, delta_tau)
        rf = np.array([r[0] for r in self.ranges], dtype=np.int32)
        rl = np.array([r[1] for r in self.ranges], dtype=np.int32)
        self._h = lib().ref_chain_create(self.N, self.M, self.nb, self.kind, self.C,
                                         rf.ctypes.data_as(c_int_p), rl.ctypes.data_as(c_int_p),
                                         self.alpha, _dp(self.eT2), _dp(self.eT2inv), _dp(self.eThalf),
                                         _dp(self.eThalfinv), int(check_sign_problem),
                                         int(check_propagation_error), int(seed), int(chain_id))
        if conf is not None:
            self.set_conf(conf)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ref_chain_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # conf is (N, M) int8, Fortran order == conf[site, slice] like fields.jl:363-368
    def set_conf(self, conf):
        c = np.asfortranarray(conf, dtype=np.int8)
        assert c.shape == (self.N, self.M)
        lib().ref_set_conf(self._h, c.ctypes.data_as(c_i8_p))

    def get_conf(self):
        c = np.zeros((self.N, self.M), dtype=np.int8, order="F")
        lib().ref_get_conf(self._h, c.ctypes.data_as(c_i8_p))
        return c

    def _mat(self):
        return np.zeros((self.N, self.N, self.nb), order="F")

    @property
    def greens(self):
        """G_eff blocks, shape (N, N, nb)."""
        g = self._mat()
        lib().ref_get_greens(self._h, _dp(g))
        return g

    def set_greens(self, g):
        g = _f(g, (self.N, self.N, self.nb))
        lib().ref_set_greens(self._h, _dp(g))

    def measured_greens(self):
        g = self._mat()
        lib().ref_measured_greens(self._h, _dp(g))
        return g

    @property
    def state(self):
        s = (C.c_int * 3)()
        lib().ref_get_state(self._h, s)
        return tuple(s)  # (current_slice, current_range, direction)

    def set_state(self, slice_, range_, direction):
        lib().ref_set_state(self._h, slice_, range_, direction)

    def build_stack(self):
        lib().ref_build_stack(self._h)

    def reverse_build_stack(self):
        lib().ref_reverse_build_stack(self._h)

    def propagate(self):
        lib().ref_propagate(self._h)

    def init(self):
        """initialize_run's stack part: reverse_build_stack + propagate (DQMC.jl:178-179)."""
        self.reverse_build_stack()
        self.propagate()

    def calculate_greens_at(self, slice_, safe_mult=None):
        g = self._mat()
        lib().ref_calculate_greens_at(self._h, int(slice_), int(safe_mult or self.safe_mult), _dp(g))
        return g

    def array(self, which, slot=0):
        names = {"u_stack": 0, "d_stack": 1, "t_stack": 2, "Ul": 3, "Dl": 4, "Tl": 5, "Ur": 6,
                 "Dr": 7, "Tr": 8, "greens_temp": 9}
        w = names[which]
        out = np.zeros((self.N, self.nb), order="F") if w in (1, 4, 7) else self._mat()
        lib().ref_get_array(self._h, w, int(slot), _dp(out))
        return out

    def multiply_slice_matrix(self, which, slice_, Mx):
        names = {"left": 0, "right": 1, "inv_right": 2, "inv_left": 3, "daggered_left": 4}
        Mx = _f(Mx, (self.N, self.N, self.nb))
        lib().ref_multiply_slice_matrix(self._h, names[which], int(slice_), _dp(Mx))
        return Mx

    def wrap_greens(self, gf, curr_slice, direction):
        gf = _f(gf, (self.N, self.N, self.nb))
        lib().ref_wrap_greens(self._h, _dp(gf), int(curr_slice), int(direction))
        return gf

    def propose_local(self, site, accept=False, u_choice=0.0):
        """site is 0-based; returns p = exp(-dE_boson) * detratio (local_updates.jl:31); u_choice: the uniform behind the
        GHQ fields' rand(1:3) (x_new = dqmc_ghq_choice(x_old, u_choice))."""
        return lib().ref_propose_local_choice(self._h, int(site), int(accept), float(u_choice))

    def sweep_spatial(self, forced=None):
        probs = np.zeros(self.N)
        dec = np.zeros(self.N, dtype=np.uint8)
        f = None if forced is None else np.ascontiguousarray(forced, dtype=np.uint8).ctypes.data_as(c_u8_p)
        acc = lib().ref_sweep_spatial(self._h, f, _dp(probs), dec.ctypes.data_as(c_u8_p))
        return acc, probs, dec

    def local_sweep(self, uniforms=None, forced=None, trace=False):
        """One full sweep (local_updates.jl:7-14).  uniforms/forced: arrays [2M, N] (C order)."""
        u = None if uniforms is None else _dp(np.ascontiguousarray(uniforms, dtype=np.float64))
        self._keep = (uniforms, forced)
        f = None if forced is None else np.ascontiguousarray(forced, dtype=np.uint8).ctypes.data_as(c_u8_p)
        if trace:
            probs = np.zeros((2 * self.M, self.N))
            dec = np.zeros((2 * self.M, self.N), dtype=np.uint8)
            acc = lib().ref_local_sweep(self._h, u, f, _dp(probs), dec.ctypes.data_as(c_u8_p))
            return acc, probs, dec
        return lib().ref_local_sweep(self._h, u, f, None, None)

    # ---- unequal-time stack (unequal_time_stack.jl) and CombinedGreensIterator (greens_iterators.jl)
    def ut_build_stack(self):
        lib().ref_ut_build_stack(self._h)

    def ut_lazy_build(self, forward_upto=0, backward_downto=0):
        lib().ref_ut_lazy_build(self._h, int(forward_upto), int(backward_downto))

    def ut_array(self, which, slot):
        """which: forward_u/d/t, backward_u/d/t, inv_u/d/t; slot 0-based."""
        names = ["forward_u", "forward_d", "forward_t", "backward_u", "backward_d", "backward_t",
                 "inv_u", "inv_d", "inv_t"]
        w = names.index(which)
        out = np.zeros((self.N, self.nb), order="F") if w % 3 == 1 else self._mat()
        lib().ref_ut_get_array(self._h, w, int(slot), _dp(out))
        return out

    def find_range_with_value(self, val):
        return lib().ref_find_range_with_value(self._h, int(val))

    def ut_calculate_greens(self, k, l):
        """calculate_greens(mc, k, l): effective G(k <- l), shape (N, N, nb)."""
        g = self._mat()
        lib().ref_ut_calculate_greens(self._h, int(k), int(l), _dp(g))
        return g

    def ut_greens(self, k, l):
        """greens(mc, k, l): measured G(k <- l)."""
        g = self._mat()
        lib().ref_ut_greens(self._h, int(k), int(l), _dp(g))
        return g

    def combined_greens_iterator(self, recalculate=None, start=0, stop=None):
        """Yields (l, G0l, Gl0, Gll) like CombinedGreensIterator (greens_iterators.jl:154-435)."""
        recalculate = 2 * self.safe_mult if recalculate is None else int(recalculate)
        stop = self.M if stop is None else int(stop)
        bufs = [self._mat() for _ in range(3)]
        nxt = lib().ref_cgi_first(self._h, recalculate, int(start), stop, self.safe_mult, *[_dp(b) for b in bufs])
        while nxt >= 0:
            yield (nxt - 1, *[b.copy() for b in bufs])
            nxt = lib().ref_cgi_next(self._h, nxt, *[_dp(b) for b in bufs])

    def global_update(self, new_conf, uniform=0.5, safe_mult=None):
        """global_update (global_updates.jl:203-219) with the proposed configuration -> (accepted, p)."""
        nc = np.asfortranarray(new_conf, dtype=np.int8)
        assert nc.shape == (self.N, self.M)
        p = C.c_double(0.0)
        acc = lib().ref_global_update(self._h, nc.ctypes.data_as(c_i8_p), float(uniform),
                                      int(safe_mult or self.safe_mult), C.byref(p))
        return acc, p.value

    def set_sweep_index(self, s):
        lib().ref_set_sweep_index(self._h, int(s))

    @property
    def stats(self):
        s = RefStats()
        lib().ref_get_stats(self._h, C.byref(s))
        return {k: getattr(s, k) for k, _ in RefStats._fields_}


def run_chains(chains, nthreads=None, warm=0, nsweeps=1):
    """CPU baseline: every chain runs init + warm + nsweeps sweeps, one chain per thread.
    Returns (elapsed seconds of the nsweeps part, accepted per chain)."""
    L = lib()
    n = len(chains)
    arr = (C.c_void_p * n)(*[c._h for c in chains])
    acc = np.zeros(n, dtype=np.int64)
    nthreads = nthreads or min(n, os.cpu_count() or 1)
    dt = L.ref_run_chains(arr, n, int(nthreads), int(warm), int(nsweeps), acc.ctypes.data_as(c_i64_p))
    return dt, acc
