"""DQMC(model; beta, delta_tau, safe_mult, ...), run!, greens, measurements -- the user-facing
API of the reference for this path, served by the CUDA library.

Mirrors src/flavors/DQMC/DQMC.jl:32-64 (constructor), :144-191 (init!), :200-225 (sweep_once!),
:252-394 (run!), parameters.jl:84-139, fields.jl:363-451, greens.jl:94-125 and the
`mc[:G] = greens_measurement(mc, model)` sugar (Measurements.jl:318-323).  Julia's `run!` is
`run_b`/`run` here (no `!` in Python identifiers).  One DQMC object drives `n_chains`
independent Markov chains of the same model and parameters (the reference would use a
Vector{DQMC}); chain 0 is what `mc.field.conf` / `mc.stack.greens` show by default.
"""
from __future__ import annotations

import math
import time

import numpy as np

from .context import (FIELD_DENSITY_GHQ, FIELD_DENSITY_HIRSCH, FIELD_MAGNETIC_GHQ, FIELD_MAGNETIC_HIRSCH,
                      Context)
from .models import HubbardModel, choose_field, hopping_matrix


def _julia_round(x: float) -> int:
    return int(round(x))          # round-half-even, like Julia's round(Int, x)


class DQMCParameters:
    """parameters.jl:33-139 (same keyword semantics, same slice arithmetic)."""

    def __init__(self, *, thermalization=100, sweeps=100, silent=False, check_sign_problem=True,
                 check_propagation_error=True, safe_mult=10, measure_rate=10, print_rate=None,
                 checkerboard=False, beta=None, delta_tau=None, slices=None):
        given = {k for k, v in (("beta", beta), ("delta_tau", delta_tau), ("slices", slices)) if v is not None}
        if given == {"beta"}:
            delta_tau = 0.1
            given = {"beta", "delta_tau"}
        if len(given) < 2:
            raise ValueError(f"Invalid keyword arguments to DQMCParameters: {sorted(given)}")
        if given == {"beta", "delta_tau", "slices"}:
            s = _julia_round(beta / delta_tau)
            if s != slices:
                raise ValueError(f"Given slices ({slices}) does not match calculated slices beta/delta_tau ~ {s}")
        elif given == {"beta", "slices"}:
            delta_tau = beta / slices
        elif given == {"delta_tau", "slices"}:
            beta = delta_tau * slices
        else:
            slices = _julia_round(beta / delta_tau)
        if checkerboard:
            raise NotImplementedError("checkerboard decomposition is outside the B200 sweep path (SURVEY 2, row 19)")
        self.thermalization, self.sweeps, self.silent = int(thermalization), int(sweeps), bool(silent)
        self.check_sign_problem, self.check_propagation_error = bool(check_sign_problem), bool(check_propagation_error)
        self.safe_mult, self.measure_rate = int(safe_mult), int(measure_rate)
        self.print_rate = max(10, (self.thermalization + self.sweeps) // 100) if print_rate is None else int(print_rate)
        self.beta, self.delta_tau, self.slices = float(beta), float(delta_tau), int(slices)
        self.checkerboard = False


def generate_chunks(length: int, max_chunk_size: int):
    """stack.jl:154-158 -> [(first, last)] 1-based inclusive."""
    n_chunks = -(-length // max_chunk_size)
    step = length / n_chunks
    return [(_julia_round((i - 1) * step) + 1, _julia_round(i * step)) for i in range(1, n_chunks + 1)]


def sym_exp(A: np.ndarray) -> np.ndarray:
    """exp of a Hermitian matrix through its eigen-decomposition (stack.jl:235-239 -> fallback_exp)."""
    w, V = np.linalg.eigh(0.5 * (A + A.T))
    return np.asfortranarray((V * np.exp(w)) @ V.T)


class Field:
    """The Hubbard-Stratonovich fields of the path: DensityHirschField / MagneticHirschField (fields.jl:363-451, conf
    = +-1) and the 4-node Gauss-Hermite DensityGHQField / MagneticGHQField (fields.jl:464-637, conf in 1..4, real
    coupling only).  Holds alpha and the Int8 configurations of every chain."""

    def __init__(self, name: str, param: DQMCParameters, model: HubbardModel, n_chains: int):
        self.name = name
        dtU = param.delta_tau * model.U
        if name == "DensityHirschField":
            self.kind = FIELD_DENSITY_HIRSCH
            self.alpha = math.acosh(math.exp(0.5 * dtU))                     # fields.jl:372
        elif name == "MagneticHirschField":
            self.kind = FIELD_MAGNETIC_HIRSCH
            self.alpha = math.acosh(math.exp(-0.5 * dtU))                    # fields.jl:421
        elif name in ("DensityGHQField", "MagneticGHQField"):
            self.kind = FIELD_DENSITY_GHQ if name == "DensityGHQField" else FIELD_MAGNETIC_GHQ
            x = (0.5 if self.kind == FIELD_DENSITY_GHQ else -0.5) * dtU      # fields.jl:576 / :514
            if x < 0:
                raise NotImplementedError(f"{name} with U = {model.U}: complex coupling sqrt({x}) -- complex matrix types "
                                          "are outside the B200 sweep path (SURVEY 2)")
            self.alpha = math.sqrt(x)
        else:
            raise NotImplementedError(f"{name}: not a field of the B200 path")
        self.ghq = self.kind >= FIELD_DENSITY_GHQ
        self.confs = np.ones((len(model.l), param.slices, n_chains), dtype=np.int8, order="F")

    @property
    def conf(self):
        return self.confs[:, :, 0]

    def rand(self, rng: np.random.Generator):
        """rand!(field) (fields.jl:330 iid +-1; :472-473 iid 1..4 for the GHQ fields)."""
        vals = np.array([1, 2, 3, 4] if self.ghq else [-1, 1], dtype=np.int8)
        self.confs[...] = rng.choice(vals, size=self.confs.shape)

    def compress(self, chain=0):
        """compress(field) -> the chunks of the BitArray: uint64 words, bit i at chunks[i >> 6], position i & 63 (Julia's
        BitArray layout).  Hirsch: BitArray(conf .== 1) (fields.jl:331); GHQ: two bits per value, (v - 1) >> 1 then
        (v - 1) & 1 (fields.jl:476-480)."""
        v = self.confs[:, :, chain].ravel(order="F")
        if self.ghq:
            bits = np.empty(2 * v.size, dtype=np.uint8)
            bits[0::2] = (v - 1) >> 1
            bits[1::2] = (v - 1) & 1
        else:
            bits = (v == 1)
        by = np.packbits(bits, bitorder="little")
        by = np.concatenate([by, np.zeros((-len(by)) % 8, dtype=np.uint8)])
        return by.view("<u8").copy()

    def decompress(self, chunks, chain=0):
        """decompress!(field, bits) (fields.jl:334: conf = 2 bit - 1; GHQ :481-489: 1 + 2 bit1 + bit2)."""
        n = self.confs.shape[0] * self.confs.shape[1]
        raw = np.unpackbits(np.asarray(chunks, dtype="<u8").view(np.uint8), bitorder="little")
        if self.ghq:
            vals = (1 + 2 * raw[0:2 * n:2] + raw[1:2 * n:2]).astype(np.int8)
        else:
            vals = raw[:n].astype(np.int8) * 2 - 1
        self.confs[:, :, chain] = vals.reshape(self.confs.shape[:2], order="F")


HirschField = Field        # the name round 1 used


class _StackView:
    """Read-only window on the device-resident DQMCStack of chain 0 (stack.jl:1-74)."""

    def __init__(self, mc):
        self._mc = mc

    @property
    def greens(self):
        g = self._mc.ctx.greens(0, 1)[:, :, :, 0]
        return g[:, :, 0] if self._mc.ctx.nb == 1 else g

    @property
    def current_slice(self):
        return self._mc.ctx.state[0]

    @property
    def direction(self):
        return self._mc.ctx.state[2]

    @property
    def ranges(self):
        return [range(a, b + 1) for a, b in self._mc.ctx.ranges]


class GreensMeasurement:
    """greens_measurement(mc, model) (measurements/constructors/greens.jl:17-40) with a plain
    {count, sum, sum of squares} accumulator per chain standing in for BinningAnalysis' LogBinner
    (third-party, arithmetic unpinned -- SURVEY 8c)."""

    def __init__(self, mc):
        nb = mc.ctx.nb
        self.count = 0
        self.sum = np.zeros((mc.ctx.N, mc.ctx.N, nb, mc.ctx.B), order="F")
        self.sumsq = np.zeros_like(self.sum)

    def push(self, G):
        self.count += 1
        self.sum += G
        self.sumsq += G * G

    def mean(self, chain=None):
        m = self.sum / max(self.count, 1)
        m = m.mean(axis=3) if chain is None else m[:, :, :, chain]
        return m[:, :, 0] if m.shape[2] == 1 else m

    def std_error(self):
        n = self.count * self.sum.shape[3]
        mean = self.sum.sum(axis=3) / n
        var = np.maximum(self.sumsq.sum(axis=3) / n - mean ** 2, 0.0)
        return np.sqrt(var / max(n - 1, 1))


def greens_measurement(mc, model=None):
    return GreensMeasurement(mc)


class DeviceMeasurement:
    """A DQMCMeasurement (measurements/generic.jl:1-120) whose Wick kernel runs on the device
    (csrc/measure.cu).  `key` names the observable of Context.OBS_NAMES; `time_integral` marks the
    TimeIntegral greens iterator (susceptibilities), otherwise Greens() (equal time).  Every measurement
    pushes one value per chain into a {count, sum, sum of squares} accumulator (LogBinner stand-in)."""

    def __init__(self, key: str, time_integral: bool):
        self.key, self.time_integral = key, time_integral
        self.count, self.sum, self.sumsq = 0, None, None

    def push(self, values):
        values = np.asarray(values, dtype=np.float64)
        if self.sum is None:
            self.sum, self.sumsq = np.zeros_like(values), np.zeros_like(values)
        self.count += 1
        self.sum += values
        self.sumsq += values * values

    def mean(self, chain=None):
        m = self.sum / max(self.count, 1)
        return m.mean(axis=0) if chain is None else m[chain]

    def std_error(self):
        n = self.count * self.sum.shape[0]
        mean = self.sum.sum(axis=0) / n
        var = np.maximum(self.sumsq.sum(axis=0) / n - mean ** 2, 0.0)
        return np.sqrt(var / max(n - 1, 1))


# constructors with the reference's names (measurements/constructors/*.jl)
def occupation(mc, model=None):
    return DeviceMeasurement("occ", False)


def kinetic_energy(mc, model=None):
    return DeviceMeasurement("K", False)


def interaction_energy(mc, model=None):
    return DeviceMeasurement("V", False)


def total_energy(mc, model=None):
    return DeviceMeasurement("E", False)


def charge_density_correlation(mc, model=None):
    return DeviceMeasurement("cdc", False)


def charge_density_susceptibility(mc, model=None):
    return DeviceMeasurement("cds", True)


def spin_density_correlation(mc, model=None, dir="z"):
    return DeviceMeasurement({"x": "sdxc", "y": "sdyc", "z": "sdzc"}[str(dir).lstrip(":")], False)


def spin_density_susceptibility(mc, model=None, dir="z"):
    return DeviceMeasurement({"x": "sdxs", "y": "sdys", "z": "sdzs"}[str(dir).lstrip(":")], True)


# ---- configuration recorder (src/configurations.jl:12-60) and replay! (DQMC.jl:418-505)
class ConfigRecorder:
    """ConfigRecorder(rate): every `rate` sweeps after thermalization the configuration of every chain is stored in the
    reference's compressed form -- the chunks of compress(field) (BitArray(conf .== 1), or two bits per value for the GHQ
    fields), packed on the device (dqmc_get_conf_packed).  configs[i] is a uint64 array (words, n_chains)."""

    def __init__(self, rate=10):
        self.rate, self.configs, self.sweeps = int(rate), [], []

    def push(self, mc, sweep):
        if sweep % self.rate == 0:                      # configurations.jl:30-33
            self.configs.append(mc.ctx.get_conf_packed())
            self.sweeps.append(int(sweep))

    def __len__(self):
        return len(self.configs)

    def __getitem__(self, i):
        return self.configs[i]

    def __iter__(self):
        return iter(self.configs)


class Discarder:
    """Discarder (configurations.jl): records nothing."""
    rate = 0

    def push(self, mc, sweep):
        pass

    def __len__(self):
        return 0


# ---- updates (updates/local_updates.jl:70-84, updates/global_updates.jl:229-270, updates/scheduler.jl:236-289)
class LocalSweep:
    is_full_sweep = True

    def __init__(self, N=1):
        self.N = int(N)


class GlobalFlip:
    is_full_sweep = True         # scheduler.jl:67 (default; global_updates.jl does not override it)


class GlobalShuffle:
    is_full_sweep = True


class SimpleScheduler:
    """SimpleScheduler(updates...): cycles through the given updates, one per sweep unless an update declares
    is_full_sweep = False (scheduler.jl:62-67, 281-289).  LocalSweep(N) expands to N local sweeps and at least one
    local sweep is required (scheduler.jl:267-278)."""

    def __init__(self, *updates):
        seq = []
        for u in updates:
            seq.extend([LocalSweep()] * u.N if isinstance(u, LocalSweep) else [u])
        if not any(isinstance(u, LocalSweep) for u in seq):
            raise ValueError("The scheduler requires local updates, but none were passed (scheduler.jl:269)")
        self.sequence, self.idx = seq, 0

    def next(self):
        u = self.sequence[self.idx]
        self.idx = (self.idx + 1) % len(self.sequence)
        return u


class DQMC:
    """DQMC(model; beta, delta_tau, safe_mult, thermalization, sweeps, measure_rate, seed, field, ...)."""

    def __init__(self, model: HubbardModel, *, seed=-1, field=None, n_chains=1, device=0, chain_offset=0,
                 delay_block=0, scheduler=None, recalculate=None, recorder=None, recording_rate=None, **kwargs):
        self.model = model
        self.parameters = DQMCParameters(**kwargs)
        self._rng = np.random.default_rng(None if seed == -1 else seed)
        self.field = Field(field or choose_field(model), self.parameters, model, n_chains)
        self.field.rand(self._rng)                                  # DQMC.jl:50
        self.measurements = {}
        self.thermalization_measurements = {}
        self.last_sweep = 0
        self.accepted = np.zeros(n_chains, dtype=np.int64)
        self.total = 0
        # init_hopping_matrices (stack.jl:209-249)
        T = hopping_matrix(model)
        if not np.allclose(T, T.T):
            raise ValueError("The hopping matrix is not approximately Hermitian (stack.jl:213-224)")
        dt = self.parameters.delta_tau
        self.hopping_matrix = T
        self.ctx = Context(
            n_sites=len(model.l), n_slices=self.parameters.slices, field_kind=self.field.kind, n_chains=n_chains,
            ranges=generate_chunks(self.parameters.slices, self.parameters.safe_mult), alpha=self.field.alpha,
            hopping_exp_squared=sym_exp(-dt * T), hopping_exp_inv_squared=sym_exp(dt * T),
            hopping_exp=sym_exp(-0.5 * dt * T), hopping_exp_inv=sym_exp(0.5 * dt * T),
            check_sign_problem=self.parameters.check_sign_problem,
            check_propagation_error=self.parameters.check_propagation_error,
            seed=(1234 if seed == -1 else seed), chain_offset=chain_offset, device=device, delay_block=delay_block)
        self.stack = _StackView(self)
        # DQMC.jl:36-37, 57: recorder = ConfigRecorder, recording_rate = measure_rate
        self.recorder = recorder if recorder is not None else ConfigRecorder(
            self.parameters.measure_rate if recording_rate is None else recording_rate)
        self.scheduler = scheduler or SimpleScheduler(LocalSweep())
        self.recalculate = recalculate          # CombinedGreensIterator's recalculate (default 2 safe_mult)
        self.global_accepted = np.zeros(n_chains, dtype=np.int64)
        self.global_total = 0
        self._lattice_set = False
        self._initialized = False

    # mc[:G] = greens_measurement(mc, model)   (Measurements.jl:318-323)
    def __setitem__(self, key, m):
        self.measurements[key] = m

    def __getitem__(self, key):
        return self.measurements[key]

    def init(self):
        """init!(mc) + the stack part of initialize_run (DQMC.jl:144-191)."""
        self.ctx.set_conf(self.field.confs)
        self.ctx.build_stack()
        self._initialized = True

    def _apply_update(self, u, uniforms=None):
        """update(u, mc, model, field) -> acceptance per chain (scheduler.jl:176-181)."""
        if isinstance(u, LocalSweep):
            acc = self.ctx.sweep(1, uniforms)
            self.accepted += acc
            self.total += 2 * self.ctx.N * self.ctx.M
            return acc / (2 * self.ctx.N * self.ctx.M)
        proposed = None
        if isinstance(u, GlobalShuffle):                       # global_updates.jl:262-270: shuffle!(conf)
            conf = self.ctx.get_conf()
            proposed = np.asfortranarray(np.stack(
                [self._rng.permutation(conf[:, :, b].ravel()).reshape(conf.shape[:2]) for b in range(self.ctx.B)], axis=2))
        acc, _ = self.ctx.global_update(self.parameters.safe_mult, proposed=proposed)
        self.global_accepted += acc
        self.global_total += 1
        return acc.astype(np.float64)

    def sweep_once(self, uniforms=None):
        """update(::SimpleScheduler, mc, model) (scheduler.jl:281-289): scheduled updates are executed until one with
        is_full_sweep(update) == true has run, then last_sweep advances.  is_full_sweep defaults to true for every
        AbstractUpdate (scheduler.jl:62-67) and neither LocalSweep nor the global updates override it
        (global_updates.jl:217-270) -- only NoUpdate and ChemicalPotentialTuning return false -- so in the reference a
        GlobalFlip / GlobalShuffle DOES count as a sweep: SimpleScheduler(LocalSweep(), GlobalFlip()) alternates them,
        one per sweep.  Returns the acceptance of the last executed update per chain."""
        while True:
            u = self.scheduler.next()
            rate = self._apply_update(u, uniforms)
            if getattr(u, "is_full_sweep", True):
                break
        self.last_sweep += 1
        return rate

    def max_acceptance(self):
        """max_acceptance(scheduler) (scheduler.jl:291-299): the largest accepted / total over the update kinds."""
        rates = [self.accepted.sum() / max(1, self.total * self.ctx.B)]
        if self.global_total:
            rates.append(self.global_accepted.sum() / max(1, self.global_total * self.ctx.B))
        return float(max(rates))

    # ---- unequal-time Green's functions (unequal_time_stack.jl, measurements/greens_iterators.jl)
    def greens_kl(self, k, l, chain=None):
        """greens(mc, k, l): G(k <- l), 0 <= k, l <= slices."""
        g = self.ctx.ut_greens(k, l)
        g = g[:, :, :, 0] if chain is None else g[:, :, :, chain]
        return g[:, :, 0] if self.ctx.nb == 1 else g

    def combined_greens_iterator(self, recalculate=None, start=0, stop=None):
        """CombinedGreensIterator(mc; recalculate, start, stop): yields (l, G0l, Gl0, Gll) for all chains."""
        return self.ctx.combined_greens_iterator(self.parameters.safe_mult, recalculate or self.recalculate, start, stop)

    def _set_lattice(self):
        if not self._lattice_set:
            l = self.model.l
            self.ctx.set_lattice(np.array(l.bravais_srctrg2dir(), dtype=np.int32), len(l.unitcell.sites),
                                 self.hopping_matrix, self.model.U)
            self._lattice_set = True

    def measure(self):
        """The measurement hand-off of sweep_once! (DQMC.jl:217-221): one equal-time and/or one TimeIntegral pass
        on the device, then every registered measurement is pushed its per-chain values."""
        dev = [m for m in self.measurements.values() if isinstance(m, DeviceMeasurement)]
        if dev:
            self._set_lattice()
            if any(not m.time_integral for m in dev):
                self.ctx.measure_equal_time()
            if any(m.time_integral for m in dev):
                self.ctx.measure_time_integral(self.parameters.safe_mult, self.parameters.delta_tau, self.recalculate)
            vals = self.ctx.measurements()
            for m in dev:
                m.push(vals[m.key])
        others = [m for m in self.measurements.values() if not isinstance(m, DeviceMeasurement)]
        if others:
            G = self.ctx.measured_greens()
            for m in others:
                m.push(G)

    def greens(self, chain=None):
        """greens(mc) (greens.jl:94-125): exp(+dtau T / 2) G_eff exp(-dtau T / 2)."""
        g = self.ctx.measured_greens()
        g = g[:, :, :, 0] if chain is None else g[:, :, :, chain]
        return g[:, :, 0] if self.ctx.nb == 1 else g

    def sync_field(self):
        self.field.confs[...] = self.ctx.get_conf()

    def analysis(self):
        """DQMCAnalysis (statistics.jl:40-50) per chain."""
        return self.ctx.stats()


def run(mc: DQMC, *, verbose=False, min_update_rate=0.001):
    """run!(mc) (DQMC.jl:252-394): thermalization sweeps, then measured sweeps."""
    p = mc.parameters
    if not mc._initialized:
        mc.init()
    t0 = time.time()
    total = p.thermalization + p.sweeps
    min_sweeps = round(1.0 / min_update_rate)                   # DQMC.jl:283
    while mc.last_sweep < total:
        mc.sweep_once()
        i = mc.last_sweep
        # sweep_once! (DQMC.jl:200-225): after thermalization the configuration goes to the recorder and the measurements
        # fire on last_sweep % measure_rate == 0
        if i > p.thermalization:
            mc.recorder.push(mc, i)
            if i % p.measure_rate == 0 and mc.measurements:
                mc.measure()
        if i > min_sweeps and mc.max_acceptance() < min_update_rate:          # DQMC.jl:294-307, every sweep
            mc.sync_field()
            return "CANCELLED_LOW_ACCEPTANCE"                                   # helpers.jl:17-22
        if verbose and i % p.print_rate == 0:
            print(f"\t{i}\n\t\tsweep dur: {(time.time() - t0) / i:.3f}s\n\t\tacc rate (local): {mc.max_acceptance():.3f}")
    mc.sync_field()
    return "SUCCESS"


def replay(mc: DQMC, configurations=None, *, measure_rate=1):
    """replay!(mc, configurations = mc.recorder; measure_rate = 1) (DQMC.jl:418-505): every measure_rate-th recorded
    configuration is decompressed on the device (dqmc_set_conf_packed), the stack and G(0) are rebuilt from it
    (calculate_greens(mc, 0) -- here reverse_build_stack + propagate, which also leaves a valid stack for the
    unequal-time measurements) and every measurement is applied."""
    configurations = mc.recorder if configurations is None else configurations
    if not mc._initialized:
        mc.init()
    for i in range(0, len(configurations), int(measure_rate)):
        mc.ctx.set_conf_packed(configurations[i])
        mc.ctx.build_stack()
        if mc.measurements:
            mc.measure()
        mc.last_sweep = i + 1
    mc.sync_field()
    return "SUCCESS"


run_b = run
