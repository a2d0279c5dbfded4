"""Thin numpy-facing wrapper of the C ABI (include/dqmc_b200.h).  One `Context` == one
`dqmc_ctx`: a batch of independent Markov chains on one B200.  All array arguments are
host numpy arrays in the reference's (Julia, column-major) layouts:

    conf     int8   (N, M, B)        conf[site, slice, chain]
    greens   float64 (N, N, nb, B)   G[:, :, flavor block, chain]

Arrays are Fortran-ordered so that the memory is exactly what Julia would pass to `ccall`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

FIELD_DENSITY_HIRSCH = 0
FIELD_MAGNETIC_HIRSCH = 1
FIELD_DENSITY_GHQ = 2        # 4-state Gauss-Hermite fields (fields.jl:464-637): conf in 1..4
FIELD_MAGNETIC_GHQ = 3


class DQMCError(RuntimeError):
    pass


def _dp(a):
    return a.ctypes.data_as(_lib.dp)


def _f64(a, shape=None):
    a = np.array(a, dtype=np.float64, order="F", copy=True)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


class Context:
    def __init__(self, *, n_sites, n_slices, field_kind, n_chains, ranges, alpha,
                 hopping_exp_squared, hopping_exp_inv_squared, hopping_exp, hopping_exp_inv,
                 check_sign_problem=True, check_propagation_error=True, seed=1234, chain_offset=0,
                 device=0, delay_block=0, update_variant=0):
        self._L = _lib.load()
        self.N, self.M, self.kind, self.B = int(n_sites), int(n_slices), int(field_kind), int(n_chains)
        self.nb = 2 if (self.kind & 1) else 1
        self.ghq = self.kind >= FIELD_DENSITY_GHQ
        self._uf = 2 if self.ghq else 1                  # uniforms per proposal in explicit tables (GHQ: + choice)
        self.ranges = [(int(a), int(b)) for a, b in ranges]
        self.C = len(self.ranges)
        rf = np.array([r[0] for r in self.ranges], dtype=np.int32)
        rl = np.array([r[1] for r in self.ranges], dtype=np.int32)
        mats = [_f64(m, (self.N, self.N)) for m in
                (hopping_exp_squared, hopping_exp_inv_squared, hopping_exp, hopping_exp_inv)]
        d = _lib.Desc()
        d.n_sites, d.n_slices, d.field_kind, d.n_chains, d.n_ranges = self.N, self.M, self.kind, self.B, self.C
        d.range_first = rf.ctypes.data_as(_lib.i32p)
        d.range_last = rl.ctypes.data_as(_lib.i32p)
        d.alpha = float(alpha)
        d.hopping_exp_squared, d.hopping_exp_inv_squared = _dp(mats[0]), _dp(mats[1])
        d.hopping_exp, d.hopping_exp_inv = _dp(mats[2]), _dp(mats[3])
        d.check_sign_problem, d.check_propagation_error = int(check_sign_problem), int(check_propagation_error)
        d.seed, d.chain_offset, d.device, d.delay_block = int(seed), int(chain_offset), int(device), int(delay_block)
        d.update_variant = int(update_variant)
        h = C.c_void_p()
        rc = self._L.dqmc_create(C.byref(d), C.byref(h))
        if rc != 0:
            raise DQMCError(f"dqmc_create failed ({rc}): {self._L.dqmc_last_error(None).decode()}")
        self._h = h

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != 0:
            raise DQMCError(f"dqmc_b200 error {rc}: {self._L.dqmc_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self._L.dqmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ field
    def set_conf(self, conf, chain0=0):
        conf = np.asfortranarray(conf, dtype=np.int8)
        if conf.ndim == 2:
            conf = conf.reshape(self.N, self.M, 1, order="F")
        assert conf.shape[:2] == (self.N, self.M)
        self._ck(self._L.dqmc_set_conf(self._h, chain0, conf.shape[2], conf.ctypes.data_as(_lib.i8p)))

    def get_conf(self, chain0=0, nchains=None, out=None):
        """out: optional preallocated (e.g. pinned) int8 buffer of N*M*nchains bytes."""
        nchains = self.B - chain0 if nchains is None else nchains
        if out is None:
            out = np.zeros((self.N, self.M, nchains), dtype=np.int8, order="F")
        assert out.dtype == np.int8 and out.size == self.N * self.M * nchains
        self._ck(self._L.dqmc_get_conf(self._h, chain0, nchains, out.ctypes.data_as(_lib.i8p)))
        return out

    def get_conf_packed(self, chain0=0, nchains=None):
        """BitArray(conf .== 1).chunks of every chain (fields.jl:331): uint64 (words, nchains)."""
        nchains = self.B - chain0 if nchains is None else nchains
        words = (self.N * self.M * self._uf + 63) // 64          # GHQ: two bits per value
        out = np.zeros((words, nchains), dtype=np.uint64, order="F")
        self._ck(self._L.dqmc_get_conf_packed(self._h, chain0, nchains, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def set_conf_packed(self, chunks, chain0=0):
        """decompress!(field, bits) (fields.jl:334) on the device."""
        chunks = np.asfortranarray(chunks, dtype=np.uint64)
        if chunks.ndim == 1:
            chunks = chunks.reshape(-1, 1, order="F")
        assert chunks.shape[0] == (self.N * self.M * self._uf + 63) // 64
        self._ck(self._L.dqmc_set_conf_packed(self._h, chain0, chunks.shape[1], chunks.ctypes.data_as(C.POINTER(C.c_uint64))))

    # ------------------------------------------------------------------ stack
    def build_stack(self):
        self._ck(self._L.dqmc_build_stack(self._h))

    def forward_build_stack(self):
        self._ck(self._L.dqmc_forward_build_stack(self._h))

    def propagate(self, n=1):
        self._ck(self._L.dqmc_propagate(self._h, int(n)))

    @property
    def state(self):
        s = (C.c_int32 * 3)()
        self._ck(self._L.dqmc_get_state(self._h, s))
        return tuple(s)

    # ------------------------------------------------------------------ sweeps
    def sweep(self, nsweeps=1, uniforms=None):
        """-> accepted flips per chain.  uniforms: (nsweeps, B, 2M, N) C-ordered table (GHQ fields: (nsweeps, B, 2M, 2, N),
        Metropolis then choice uniforms) or None."""
        acc = np.zeros(self.B, dtype=np.int64)
        u = None
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
            assert uniforms.size == nsweeps * self.B * 2 * self.M * self.N * self._uf
            u = _dp(uniforms)
        self._ck(self._L.dqmc_sweep(self._h, int(nsweeps), u, acc.ctypes.data_as(_lib.i64p)))
        return acc

    def sweep_traced(self, uniforms=None, forced=None):
        """One sweep -> (accepted[B], probs[B, 2M, N], decisions[B, 2M, N])."""
        shp = (self.B, 2 * self.M, self.N)
        probs = np.zeros(shp)
        dec = np.zeros(shp, dtype=np.uint8)
        acc = np.zeros(self.B, dtype=np.int64)
        u = f = None
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64); assert uniforms.size == self._uf * probs.size
            u = _dp(uniforms)
        if forced is not None:
            forced = np.ascontiguousarray(forced, dtype=np.uint8); assert forced.shape == shp
            f = forced.ctypes.data_as(_lib.u8p)
        self._ck(self._L.dqmc_sweep_traced(self._h, u, f, _dp(probs), dec.ctypes.data_as(_lib.u8p),
                                           acc.ctypes.data_as(_lib.i64p)))
        return acc, probs, dec

    def sweep_spatial(self, uniforms=None, forced=None):
        shp = (self.B, self.N)
        probs = np.zeros(shp)
        dec = np.zeros(shp, dtype=np.uint8)
        acc = np.zeros(self.B, dtype=np.int64)
        u = f = None
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64); assert uniforms.size == self._uf * probs.size
            u = _dp(uniforms)
        if forced is not None:
            forced = np.ascontiguousarray(forced, dtype=np.uint8); assert forced.shape == shp
            f = forced.ctypes.data_as(_lib.u8p)
        self._ck(self._L.dqmc_sweep_spatial(self._h, u, f, _dp(probs), dec.ctypes.data_as(_lib.u8p),
                                            acc.ctypes.data_as(_lib.i64p)))
        return acc, probs, dec

    def global_update(self, safe_mult, proposed=None, uniforms=None):
        """global_update (global_updates.jl:203-219) for every chain.  proposed: (N, M, B) int8 configurations or
        None for GlobalFlip.  -> (accepted[B] in {0, 1}, p[B])."""
        acc = np.zeros(self.B, dtype=np.int64)
        probs = np.zeros(self.B)
        pc = uu = None
        if proposed is not None:
            proposed = np.asfortranarray(proposed, dtype=np.int8)
            assert proposed.shape == (self.N, self.M, self.B)
            pc = proposed.ctypes.data_as(_lib.i8p)
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
            assert uniforms.shape == (self.B,)
            uu = _dp(uniforms)
        self._ck(self._L.dqmc_global_update(self._h, pc, uu, int(safe_mult), acc.ctypes.data_as(_lib.i64p), _dp(probs)))
        return acc, probs

    def set_sweep_index(self, s):
        self._ck(self._L.dqmc_set_sweep_index(self._h, int(s)))

    def set_global_update_index(self, i):
        self._ck(self._L.dqmc_set_global_update_index(self._h, int(i)))

    # ------------------------------------------------------------------ NCCL communicator owned by the context
    @staticmethod
    def comm_unique_id():
        """ncclGetUniqueId -> 128 bytes (rank 0; ship them to every rank)."""
        buf = (C.c_uint8 * 128)()
        rc = _lib.load().dqmc_comm_unique_id(buf)
        if rc != 0:
            raise DQMCError(f"dqmc_comm_unique_id failed ({rc}): {_lib.load().dqmc_last_error(None).decode()}")
        return bytes(buf)

    def comm_init(self, n_ranks, rank, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self._L.dqmc_comm_init(self._h, int(n_ranks), int(rank), buf))

    def reduce_observables(self, comm=None):
        """ncclAllReduce(sum) of the accumulator blocks on the context's stream (dqmc_reduce_observables)."""
        self._ck(self._L.dqmc_reduce_observables(self._h, C.c_void_p(comm) if comm else None))

    # ------------------------------------------------------------------ results
    def _gshape(self, nchains):
        return (self.N, self.N, self.nb, nchains)

    def greens(self, chain0=0, nchains=None, out=None):
        """out: optional preallocated (e.g. pinned) float64 buffer of N*N*nb*nchains elements."""
        nchains = self.B - chain0 if nchains is None else nchains
        if out is None:
            out = np.zeros(self._gshape(nchains), order="F")
        assert out.dtype == np.float64 and out.size == self.N * self.N * self.nb * nchains
        self._ck(self._L.dqmc_get_greens(self._h, chain0, nchains, _dp(out)))
        return out

    def set_greens(self, G, chain0=0):
        G = _f64(G)
        if G.ndim == 3:
            G = G.reshape(self.N, self.N, self.nb, 1, order="F")
        self._ck(self._L.dqmc_set_greens(self._h, chain0, G.shape[3], _dp(G)))

    def measured_greens(self, chain0=0, nchains=None):
        nchains = self.B - chain0 if nchains is None else nchains
        out = np.zeros(self._gshape(nchains), order="F")
        self._ck(self._L.dqmc_get_measured_greens(self._h, chain0, nchains, _dp(out)))
        return out

    def calculate_greens_at(self, slice_, safe_mult):
        out = np.zeros(self._gshape(self.B), order="F")
        self._ck(self._L.dqmc_calculate_greens_at(self._h, int(slice_), int(safe_mult), _dp(out)))
        return out

    def stats(self):
        arr = (_lib.Stats * self.B)()
        self._ck(self._L.dqmc_get_stats(self._h, 0, self.B, arr))
        return [{k: getattr(s, k) for k, _ in _lib.Stats._fields_} for s in arr]

    _WHICH = {"u_stack": 0, "d_stack": 1, "t_stack": 2, "Ul": 3, "Dl": 4, "Tl": 5, "Ur": 6, "Dr": 7, "Tr": 8}

    def stack_array(self, which, slot=1, chain=0):
        w = self._WHICH[which]
        out = np.zeros((self.N, self.nb), order="F") if w in (1, 4, 7) else np.zeros((self.N, self.N, self.nb), order="F")
        self._ck(self._L.dqmc_get_stack_array(self._h, int(chain), w, int(slot), _dp(out)))
        return out

    # ------------------------------------------------------------------ unequal-time Green's functions
    _UT_WHICH = {"forward_u": 0, "forward_d": 1, "forward_t": 2, "backward_u": 3, "backward_d": 4,
                 "backward_t": 5, "inv_u": 6, "inv_d": 7, "inv_t": 8}

    def ut_build_stack(self):
        """build_stack(mc, mc.ut_stack) (unequal_time_stack.jl:128-185)."""
        self._ck(self._L.dqmc_ut_build_stack(self._h))

    def ut_lazy_build(self, forward_upto=0, backward_downto=0):
        self._ck(self._L.dqmc_ut_lazy_build(self._h, int(forward_upto), int(backward_downto)))

    def ut_stack_array(self, which, slot=1, chain=0):
        w = self._UT_WHICH[which]
        out = np.zeros((self.N, self.nb), order="F") if w % 3 == 1 else np.zeros((self.N, self.N, self.nb), order="F")
        self._ck(self._L.dqmc_ut_get_stack_array(self._h, int(chain), w, int(slot), _dp(out)))
        return out

    def ut_greens(self, k, l, measured=True):
        """greens(mc, k, l) (measured=True) / calculate_greens(mc, k, l) for all chains -> (N, N, nb, B)."""
        out = np.zeros(self._gshape(self.B), order="F")
        self._ck(self._L.dqmc_ut_greens(self._h, int(k), int(l), int(bool(measured)), _dp(out)))
        return out

    def combined_greens_iterator(self, safe_mult, recalculate=None, start=0, stop=None, fetch=True):
        """CombinedGreensIterator (greens_iterators.jl:154-435): yields (l, G0l, Gl0, Gll), each
        (N, N, nb, B); with fetch=False the matrices stay on the device and None is yielded instead."""
        recalculate = 2 * int(safe_mult) if recalculate is None else int(recalculate)
        stop = self.M if stop is None else int(stop)
        self._ck(self._L.dqmc_cgi_begin(self._h, recalculate, int(start), stop, int(safe_mult)))
        l = C.c_int32(0)
        while True:
            bufs = [np.zeros(self._gshape(self.B), order="F") for _ in range(3)] if fetch else [None] * 3
            ptrs = [(_dp(b) if b is not None else None) for b in bufs]
            self._ck(self._L.dqmc_cgi_next(self._h, C.byref(l), *ptrs))
            if l.value < 0:
                return
            yield (l.value, *bufs)

    # ------------------------------------------------------------------ device-side Wick kernels
    OBS_NAMES = ("occ", "K", "V", "E", "cdc", "sdxc", "sdyc", "sdzc", "cds", "sdxs", "sdys", "sdzs")

    def set_lattice(self, srctrg2dir, n_basis, hopping_matrix, U):
        """srctrg2dir: (n_bravais, n_bravais) 0-based Bravais direction of every (src, trg) pair."""
        s2d = np.asfortranarray(np.array(srctrg2dir, dtype=np.int32) + 1)
        nbr = s2d.shape[0]
        T = _f64(hopping_matrix, (self.N, self.N))
        self._ck(self._L.dqmc_set_lattice(self._h, nbr, int(n_basis), s2d.ctypes.data_as(_lib.i32p), _dp(T), float(U)))
        self._nbr, self._nbasis = nbr, int(n_basis)
        off = (C.c_int32 * (len(self.OBS_NAMES) + 1))()
        self._ck(self._L.dqmc_measurement_layout(self._h, off))
        self._obs_off = list(off)

    def measure_equal_time(self):
        self._ck(self._L.dqmc_measure_equal_time(self._h))

    def measure_time_integral(self, safe_mult, delta_tau, recalculate=None):
        recalculate = 2 * int(safe_mult) if recalculate is None else int(recalculate)
        self._ck(self._L.dqmc_measure_time_integral(self._h, recalculate, int(safe_mult), float(delta_tau)))

    def _split_obs(self, vec):
        out = {}
        for k, name in enumerate(self.OBS_NAMES):
            v = vec[..., self._obs_off[k]:self._obs_off[k + 1]]
            if name == "occ":
                pass
            elif name in ("K", "V", "E"):
                v = v[..., 0]
            else:
                v = v.reshape(v.shape[:-1] + (self._nbasis, self._nbasis, self._nbr)).swapaxes(-1, -3)
            out[name] = v
        return out

    def measurements(self, chain0=0, nchains=None):
        """-> dict name -> per-chain values of the last measurement; pair observables are (B, n_bravais, basis, basis)."""
        nchains = self.B - chain0 if nchains is None else nchains
        out = np.zeros((nchains, self._obs_off[-1]))
        self._ck(self._L.dqmc_get_measurements(self._h, chain0, nchains, _dp(out)))
        return self._split_obs(out)

    def measurement_stats(self):
        """-> (count_equal_time, count_time_integral, dict of sums, dict of sums of squares)."""
        n = self._obs_off[-1]
        cnt = np.zeros(2); s = np.zeros(n); s2 = np.zeros(n)
        self._ck(self._L.dqmc_get_measurement_stats(self._h, _dp(cnt), _dp(s), _dp(s2)))
        return cnt[0], cnt[1], self._split_obs(s), self._split_obs(s2)

    def measurement_binning(self):
        """Log-binning levels of every observable element (LogBinner semantics): -> (counts[2, L], sums, sumsqs) with
        sums / sumsqs dicts of arrays whose leading axis is the level."""
        L, n = int(self._L.dqmc_binning_levels()), self._obs_off[-1]
        cnt = np.zeros((2, L)); s = np.zeros((L, n)); s2 = np.zeros((L, n))
        self._ck(self._L.dqmc_get_measurement_binning(self._h, _dp(cnt), _dp(s), _dp(s2)))
        return cnt, self._split_obs(s), self._split_obs(s2)

    def binning_std_errors(self, key, time_integral=False):
        """std_error of observable `key` at every populated binning level (per element) -> (levels, n_l, err[levels, ...])."""
        cnt, s, s2 = self.measurement_binning()
        c = cnt[1 if time_integral else 0]
        lv = np.nonzero(c > 1)[0]
        shape = (-1,) + (1,) * (s[key].ndim - 1)
        mean = s[key][lv] / c[lv].reshape(shape)
        var = np.maximum(s2[key][lv] / c[lv].reshape(shape) - mean ** 2, 0.0)
        return lv, c[lv], np.sqrt(var / (c[lv].reshape(shape) - 1))

    # ------------------------------------------------------------------ observables
    def accumulate_greens(self):
        self._ck(self._L.dqmc_accumulate_greens(self._h))

    def observable_buffer(self):
        p = C.c_void_p(); n = C.c_int64()
        self._ck(self._L.dqmc_observable_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def observables(self):
        cnt = np.zeros(1)
        s = np.zeros((self.N, self.N, self.nb), order="F")
        s2 = np.zeros((self.N, self.N, self.nb), order="F")
        self._ck(self._L.dqmc_get_observables(self._h, _dp(cnt), _dp(s), _dp(s2)))
        return float(cnt[0]), s, s2

    # ------------------------------------------------------------------ operator level on the context
    _SLICE_OPS = {"left": 0, "right": 1, "inv_right": 2, "inv_left": 3, "daggered_left": 4}

    def multiply_slice_matrix(self, which, slice_, X):
        X = _f64(X, self._gshape(self.B))
        self._ck(self._L.dqmc_op_multiply_slice_matrix(self._h, self._SLICE_OPS[which], int(slice_), _dp(X)))
        return X

    def wrap_greens(self, X, curr_slice, direction):
        X = _f64(X, self._gshape(self.B))
        self._ck(self._L.dqmc_op_wrap_greens(self._h, int(curr_slice), int(direction), _dp(X)))
        return X

    def kernel_launches(self):
        return int(self._L.dqmc_kernel_launches(self._h))

    def stream_handle(self):
        """cudaStream_t of the context as an integer (torch.cuda.ExternalStream(handle))."""
        p = C.c_void_p()
        self._ck(self._L.dqmc_get_stream(self._h, C.byref(p)))
        return p.value or 0

    PROF_NAMES = ("gemm", "udt", "rdivp", "update", "other")

    def profile(self, enable=True):
        self._ck(self._L.dqmc_profile(self._h, int(enable)))

    def profile_report(self):
        ms = np.zeros(len(self.PROF_NAMES)); cnt = np.zeros(len(self.PROF_NAMES), dtype=np.int64)
        self._ck(self._L.dqmc_profile_report(self._h, _dp(ms), cnt.ctypes.data_as(_lib.i64p)))
        return {n: {"ms": float(ms[i]), "count": int(cnt[i])} for i, n in enumerate(self.PROF_NAMES)}

    def observable_tensor(self, torch):
        """The device accumulator block [count | sum | sumsq] as a torch tensor (no copy), for
        torch.distributed.all_reduce over NCCL."""
        ptr, n = self.observable_buffer()

        class _Buf:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Buf(), device="cuda")


# ---------------------------------------------------------------------- operator level, stateless
def _op_err(rc):
    if rc != 0:
        raise DQMCError(f"dqmc_b200 op error {rc}: {_lib.load().dqmc_last_error(None).decode()}")


def _batched(a):
    a = np.array(a, dtype=np.float64, order="F", copy=True)
    return a.reshape(a.shape[0], a.shape[1], 1, order="F") if a.ndim == 2 else a


def vmul(A, B, transA=False, transB=False, device=0):
    """vmul!(C, op(A), op(B)); A, B: (n, n) or (n, n, batch)."""
    A3, B3 = _batched(A), _batched(B)
    out = np.zeros_like(A3, order="F")
    _op_err(_lib.load().dqmc_op_vmul(device, A3.shape[0], A3.shape[2], int(transA), int(transB), _dp(A3), _dp(B3), _dp(out)))
    return out if np.ndim(A) == 3 else out[:, :, 0]


def udt_AVX_pivot(X, apply_pivot=True, device=0):
    """-> U, D, T, pivot (1-based int64) like udt_AVX_pivot!(U, D, T, pivot, temp, Val(apply_pivot))."""
    X3 = _batched(X)
    n, _, b = X3.shape
    U = np.zeros_like(X3, order="F"); T = np.zeros_like(X3, order="F")
    D = np.zeros((n, b), order="F"); piv = np.zeros((n, b), dtype=np.int64, order="F")
    _op_err(_lib.load().dqmc_op_udt(device, n, b, int(apply_pivot), _dp(X3), _dp(U), _dp(D), _dp(T),
                                     piv.ctypes.data_as(_lib.i64p)))
    if np.ndim(X) == 2:
        return U[:, :, 0], D[:, 0], T[:, :, 0], piv[:, 0]
    return U, D, T, piv


def rdivp(A, T, pivot, device=0):
    A3, T3 = _batched(A), _batched(T)
    n, _, b = A3.shape
    piv = np.asfortranarray(np.array(pivot, dtype=np.int64).reshape(n, b, order="F"))
    _op_err(_lib.load().dqmc_op_rdivp(device, n, b, _dp(A3), _dp(T3), piv.ctypes.data_as(_lib.i64p)))
    return A3 if np.ndim(A) == 3 else A3[:, :, 0]


def calculate_greens_AVX(Ul, Dl, Tl, Ur, Dr, Tr, device=0):
    m = [_batched(x) for x in (Ul, Tl, Ur, Tr)]
    n, _, b = m[0].shape
    dl = np.asfortranarray(np.array(Dl, dtype=np.float64).reshape(n, b, order="F"))
    dr = np.asfortranarray(np.array(Dr, dtype=np.float64).reshape(n, b, order="F"))
    G = np.zeros_like(m[0], order="F")
    _op_err(_lib.load().dqmc_op_calculate_greens(device, n, b, _dp(m[0]), _dp(dl), _dp(m[1]), _dp(m[2]), _dp(dr),
                                                  _dp(m[3]), _dp(G)))
    return G if np.ndim(Ul) == 3 else G[:, :, 0]
