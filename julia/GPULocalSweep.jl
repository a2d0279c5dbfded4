# GPULocalSweep.jl -- the reference-side binding of libdqmc_b200.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image nor on the GPU box.
# This file is what a MonteCarlo.jl maintainer would add (e.g. as ext/MonteCarloB200Ext.jl); it uses only
# documented extension points of the package:
#   * `AbstractLocalUpdate`             src/flavors/DQMC/updates/scheduler.jl:44-45
#   * `init!(mc, update)`               scheduler.jl:211-216 (called from init!(mc), DQMC.jl:147)
#   * `update(u, mc, model, field)`     scheduler.jl:176-181, 281-289 -> accepted fraction ::Float64
# User code stays
#   mc = DQMC(model; beta=16.0, delta_tau=0.1, safe_mult=10, scheduler=SimpleScheduler(GPULocalSweep()))
#   mc[:G] = greens_measurement(mc, model);  run!(mc)
module MonteCarloB200

using MonteCarlo
import MonteCarlo: AbstractLocalUpdate, DQMC, init!, update, name, field, conf, nslices, lattice
import MonteCarlo: DensityHirschField, MagneticHirschField, DensityGHQField, MagneticGHQField

const LIB = get(ENV, "DQMC_B200_LIB", "libdqmc_b200.so")

# mirrors `dqmc_desc` of include/dqmc_b200.h field by field
struct DQMCDesc
    n_sites::Int32; n_slices::Int32; field_kind::Int32; n_chains::Int32; n_ranges::Int32
    range_first::Ptr{Int32}; range_last::Ptr{Int32}
    alpha::Float64
    hopping_exp_squared::Ptr{Float64}; hopping_exp_inv_squared::Ptr{Float64}
    hopping_exp::Ptr{Float64}; hopping_exp_inv::Ptr{Float64}
    check_sign_problem::Int32; check_propagation_error::Int32
    seed::UInt64; chain_offset::Int64; device::Int32; delay_block::Int32; update_variant::Int32
end

# One library context = one batch of independent Markov chains on one GPU.  The reference's DQMC object owns exactly
# one configuration (flavors/DQMC/main.jl:26-62), so a batch is a Vector{DQMC} of the same model / parameters whose
# GPULocalSweep updates share ONE GPUChains object: the first `update` of a sweep round launches the sweep of all
# chains, the other chains' `update` calls of that round only copy their own conf / G / accepted count back.
mutable struct GPUChains
    ctx::Ptr{Cvoid}
    n_chains::Int
    device::Int
    seed::UInt64
    chain_offset::Int64              # global index of chain 1 (multi-GPU sharding)
    rng::Symbol                      # :counter -- the library's Philox stream (include/dqmc_rng.h), nothing crosses
                                     #             the bus per sweep;  :julia -- uniforms drawn with Julia's rand
    round::Int                       # sweeps launched so far
    harvested::Vector{Int}           # per chain: the round its host state reflects
    accepted::Vector{Int64}
    confs::Vector{Matrix{Int8}}      # the chains' conf matrices, registered by init!
end
GPUChains(n_chains::Int; device = 0, seed = 0x1234, chain_offset = 0, rng = :counter) =
    GPUChains(C_NULL, n_chains, device, UInt64(seed), Int64(chain_offset), rng, 0, zeros(Int, n_chains),
              zeros(Int64, n_chains), Vector{Matrix{Int8}}(undef, n_chains))

struct GPULocalSweep <: AbstractLocalUpdate
    chains::GPUChains
    chain::Int                       # 1-based index of this DQMC object inside the batch
end
# single-chain form: `scheduler = SimpleScheduler(GPULocalSweep())`
GPULocalSweep(; device = 0, seed = 0x1234, rng = :julia) = GPULocalSweep(GPUChains(1; device, seed, rng), 1)
name(::GPULocalSweep) = "GPULocalSweep"

"""
    gpu_batch(make_mc, n_chains; device, seed, chain_offset, rng) -> Vector{DQMC}

`make_mc(scheduler)` builds one DQMC (same model and parameters for every call); the returned simulations share one
library context.  Drive them in lockstep: `for sweep in ...; for mc in mcs; MonteCarlo.sweep_once!(mc, ...) end end`
(or `run_batch!`).  Every `mc` keeps its own measurements and LogBinners.
"""
function gpu_batch(make_mc, n_chains::Int; kwargs...)
    chains = GPUChains(n_chains; kwargs...)
    return [make_mc(SimpleScheduler(GPULocalSweep(chains, b))) for b in 1:n_chains]
end

function check(c::GPUChains, rc::Int32)
    rc == 0 && return
    msg = unsafe_string(ccall((:dqmc_last_error, LIB), Cstring, (Ptr{Cvoid},), c.ctx))
    error("dqmc_b200 ($rc): $msg")     # maps to error()/ExitCode, src/helpers.jl:17-22
end
check(u::GPULocalSweep, rc::Int32) = check(u.chains, rc)
Base.getproperty(u::GPULocalSweep, s::Symbol) = s === :ctx ? getfield(u, :chains).ctx : getfield(u, s)

field_kind(::DensityHirschField) = Int32(0)
field_kind(::MagneticHirschField) = Int32(1)
field_kind(f::DensityGHQField) = f.α isa Real ? Int32(2) : error("DensityGHQField with complex α is outside the B200 path")
field_kind(f::MagneticGHQField) = f.α isa Real ? Int32(3) : error("MagneticGHQField with complex α is outside the B200 path")
is_ghq(f) = f isa MonteCarlo.AbstractGHQField

# init!(mc) has already run init_hopping_matrices + initialize_stack (DQMC.jl:144-148), so the
# exponentials and ranges exist.  Both flavor blocks of a BlockDiagonal hold the same N x N matrix
# (DQMC_interface.jl:288), so block 1 is passed.
dense(H::Hermitian) = Matrix{Float64}(H)
dense(H::Hermitian{<:Any, <:MonteCarlo.BlockDiagonal}) = Matrix{Float64}(parent(H).blocks[1])

# The context is created by the first chain that is initialised (the model data are the same for all chains); every
# chain uploads its own configuration; the stack is built once the last chain has registered.
function init!(mc::DQMC, u::GPULocalSweep)
    c, f = u.chains, field(mc)
    if c.ctx == C_NULL
        s, p = mc.stack, mc.parameters
        rf = Int32[first(r) for r in s.ranges]; rl = Int32[last(r) for r in s.ranges]
        e2, e2i = dense(s.hopping_matrix_exp_squared), dense(s.hopping_matrix_exp_inv_squared)
        eh, ehi = dense(s.hopping_matrix_exp), dense(s.hopping_matrix_exp_inv)
        ctx = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve rf rl e2 e2i eh ehi begin
            d = DQMCDesc(length(lattice(mc)), p.slices, field_kind(f), c.n_chains, length(rf), pointer(rf), pointer(rl),
                         Float64(f.α), pointer(e2), pointer(e2i), pointer(eh), pointer(ehi),
                         p.check_sign_problem, p.check_propagation_error, c.seed, c.chain_offset, c.device, 0, 0)
            rc = ccall((:dqmc_create, LIB), Int32, (Ref{DQMCDesc}, Ref{Ptr{Cvoid}}), d, ctx)
            rc == 0 || error("dqmc_create: " * unsafe_string(ccall((:dqmc_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        end
        c.ctx = ctx[]
        finalizer(x -> x.ctx != C_NULL && ccall((:dqmc_destroy, LIB), Int32, (Ptr{Cvoid},), x.ctx), c)
    end
    cf = conf(f)                                  # Matrix{Int8}(N, M), column-major == ABI layout
    c.confs[u.chain] = cf
    check(c, ccall((:dqmc_set_conf, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int8}), c.ctx, u.chain - 1, 1, cf))
    if all(b -> isassigned(c.confs, b), 1:c.n_chains)
        check(c, ccall((:dqmc_build_stack, LIB), Int32, (Ptr{Cvoid},), c.ctx))   # reverse_build_stack + propagate
        c.round = 0; fill!(c.harvested, 0)
    end
    nothing
end

# One sweep of ALL chains.  rng = :julia keeps Julia's RNG in charge: the whole table of Metropolis uniforms (and, for
# the GHQ fields, the choice uniforms (rand(1:3) - 0.5) / 3) is drawn up front -- the reference draws rand() only when
# p <= 1 (local_updates.jl:53), so the stream position differs from a CPU run with the same seed (documented
# divergence; the Markov chain is equally valid).
function launch_round!(c::GPUChains, N::Int, M::Int, ghq::Bool)
    if c.rng === :julia
        u = ghq ? Array{Float64}(undef, N, 2, 2M, c.n_chains) : rand(Float64, N, 2M, c.n_chains)   # C layout [B][2M]([2])[N]
        if ghq
            u[:, 1, :, :] .= rand.()
            u[:, 2, :, :] .= (rand.(Ref(1:3)) .- 0.5) ./ 3
        end
        check(c, ccall((:dqmc_sweep, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Int64}), c.ctx, 1, u, c.accepted))
    else
        check(c, ccall((:dqmc_sweep, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Int64}), c.ctx, 1, C_NULL, c.accepted))
    end
    c.round += 1
end

function update(u::GPULocalSweep, mc::DQMC, model, f)
    c, b = u.chains, u.chain
    N, M = size(conf(f))
    if c.harvested[b] == c.round
        launch_round!(c, N, M, is_ghq(f))          # first chain of a new round: sweep everybody
    elseif c.harvested[b] != c.round - 1
        error("GPULocalSweep: the chains of a batch must be swept in lockstep (chain $b is $(c.round - c.harvested[b]) rounds behind)")
    end
    # hand this chain's state back so that every existing measurement keeps working (generic.jl:287-288)
    check(c, ccall((:dqmc_get_conf, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int8}), c.ctx, b - 1, 1, conf(f)))
    G = mc.stack.greens
    if G isa MonteCarlo.BlockDiagonal
        buf = Array{Float64}(undef, N, N, length(G.blocks))
        check(c, ccall((:dqmc_get_greens, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}), c.ctx, b - 1, 1, buf))
        for k in eachindex(G.blocks); copyto!(G.blocks[k], view(buf, :, :, k)); end
    else
        check(c, ccall((:dqmc_get_greens, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}), c.ctx, b - 1, 1, G))
    end
    # the invariant after a sweep on both sides (stack.jl:50-52); the host copies of u/d/t_stack, Ul..Tr are NOT refreshed
    # (fetch them with dqmc_get_stack_array if a host-side global update or ut-stack needs them)
    mc.stack.current_slice = 1; mc.stack.direction = 1
    c.harvested[b] = c.round
    return c.accepted[b] / (2 * N * M)            # local_updates.jl:82
end

"""
    run_batch!(mcs; thermalization, sweeps, measure_rate)

run! (DQMC.jl:252-394) for a batch created by `gpu_batch`, without the file I/O: sweeps all chains in lockstep and lets
every chain's own measurement groups fire at its own `last_sweep`.
"""
function run_batch!(mcs::Vector{<:DQMC}; thermalization = mcs[1].parameters.thermalization, sweeps = mcs[1].parameters.sweeps)
    groups = map(mcs) do mc
        MonteCarlo.init!(mc)
        MonteCarlo.generate_groups(mc, mc.model, mc.thermalization_measurements), MonteCarlo.generate_groups(mc, mc.model, mc.measurements)
    end
    for _ in 1:(thermalization + sweeps), (mc, (thg, g)) in zip(mcs, groups)
        MonteCarlo.sweep_once!(mc, thg, g, thermalization)
    end
    return mcs
end

# ---------------------------------------------------------------------------------------------------
# Global updates on the GPU: one more AbstractGlobalUpdate (updates/global_updates.jl:229-236).  The
# scheduler calls update(u, mc, model, field) -> 0 / 1 like for the built-in GlobalFlip.
# ---------------------------------------------------------------------------------------------------
struct GPUGlobalFlip <: MonteCarlo.AbstractGlobalUpdate
    sweep::GPULocalSweep                      # shares the context of the local sweep
end
name(::GPUGlobalFlip) = "GPUGlobalFlip"
MonteCarlo.requires_temp_conf(::GPUGlobalFlip) = false     # temp_conf lives on the device

function update(u::GPUGlobalFlip, mc::DQMC, model, f)
    s = u.sweep
    accepted = Ref{Int64}(0)
    uniform = Ref(rand())                      # Julia keeps owning the RNG stream
    s.chains.n_chains == 1 || error("GPUGlobalFlip: batched chains decide per chain inside the library; call dqmc_global_update once per round")
    check(s, ccall((:dqmc_global_update, LIB), Int32,
                   (Ptr{Cvoid}, Ptr{Int8}, Ref{Float64}, Int32, Ref{Int64}, Ptr{Float64}),
                   s.ctx, C_NULL, uniform, mc.parameters.safe_mult, accepted, C_NULL))
    check(s, ccall((:dqmc_get_conf, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int8}), s.ctx, 0, 1, conf(f)))
    return Int(accepted[])
end

# ---------------------------------------------------------------------------------------------------
# Unequal-time Green's functions: a greens iterator type next to TimeIntegral / CombinedGreensIterator
# (measurements/greens_iterators.jl:79-104, 154-196).  init(mc, it) returns an iterable whose elements are
# the same (G0l, Gl0, Gll) GreensMatrix triples, so apply!(::TimeIntegral, ...) (generic.jl:337-372) and
# every existing susceptibility measurement work unchanged.
# ---------------------------------------------------------------------------------------------------
struct GPUCombinedGreensIterator <: MonteCarlo.AbstractUnequalTimeGreensIterator
    sweep::GPULocalSweep
    recalculate::Int
    start::Int
    stop::Int
end
struct _GPUCGI{T <: DQMC}
    mc::T
    spec::GPUCombinedGreensIterator
    bufs::NTuple{3, Array{Float64, 3}}
end
MonteCarlo.init(mc::DQMC, it::GPUCombinedGreensIterator) = begin
    N = length(lattice(mc)); nb = mc.stack.greens isa MonteCarlo.BlockDiagonal ? 2 : 1
    _GPUCGI(mc, it, ntuple(_ -> Array{Float64}(undef, N, N, nb), 3))
end
Base.length(it::_GPUCGI) = it.spec.stop - it.spec.start + 1

function Base.iterate(it::_GPUCGI, started::Bool = false)
    s = it.spec.sweep
    if !started
        check(s, ccall((:dqmc_cgi_begin, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Int32),
                       s.ctx, it.spec.recalculate, it.spec.start, it.spec.stop, it.mc.parameters.safe_mult))
    end
    l = Ref{Int32}(-1)
    check(s, ccall((:dqmc_cgi_next, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   s.ctx, l, it.bufs[1], it.bufs[2], it.bufs[3]))
    l[] < 0 && return nothing
    wrap(A) = size(A, 3) == 1 ? A[:, :, 1] : MonteCarlo.BlockDiagonal(A[:, :, 1], A[:, :, 2])
    return ((MonteCarlo.GreensMatrix(0, Int(l[]), wrap(it.bufs[1])),
             MonteCarlo.GreensMatrix(Int(l[]), 0, wrap(it.bufs[2])),
             MonteCarlo.GreensMatrix(Int(l[]), Int(l[]), wrap(it.bufs[3]))), true)
end

# greens(mc, k, l) served by the device-side UnequalTimeStack (unequal_time_stack.jl:302-335)
function gpu_greens(u::GPULocalSweep, mc::DQMC, k::Int, l::Int)
    N = length(lattice(mc)); nb = mc.stack.greens isa MonteCarlo.BlockDiagonal ? 2 : 1
    G = Array{Float64}(undef, N, N, nb)
    check(u, ccall((:dqmc_ut_greens, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}), u.ctx, k, l, 1, G))
    return MonteCarlo.GreensMatrix(k, l, nb == 1 ? G[:, :, 1] : MonteCarlo.BlockDiagonal(G[:, :, 1], G[:, :, 2]))
end

end # module
