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

const LIB = get(ENV, "DQMC_B200_LIB", "libdqmc_b200.so")

# mirrors `dqmc_desc` of include/dqmc_b200.h field by field
struct DQMCDesc
    n_sites::Int32; n_slices::Int32; field_kind::Int32; n_chains::Int32; n_ranges::Int32
    range_first::Ptr{Int32}; range_last::Ptr{Int32}
    alpha::Float64
    hopping_exp_squared::Ptr{Float64}; hopping_exp_inv_squared::Ptr{Float64}
    hopping_exp::Ptr{Float64}; hopping_exp_inv::Ptr{Float64}
    check_sign_problem::Int32; check_propagation_error::Int32
    seed::UInt64; chain_offset::Int64; device::Int32; delay_block::Int32
end

mutable struct GPULocalSweep <: AbstractLocalUpdate
    ctx::Ptr{Cvoid}
    device::Int
    seed::UInt64
    GPULocalSweep(; device = 0, seed = 0x1234) = new(C_NULL, device, seed)
end
name(::GPULocalSweep) = "GPULocalSweep"

function check(u::GPULocalSweep, rc::Int32)
    rc == 0 && return
    msg = unsafe_string(ccall((:dqmc_last_error, LIB), Cstring, (Ptr{Cvoid},), u.ctx))
    error("dqmc_b200 ($rc): $msg")     # maps to error()/ExitCode, src/helpers.jl:17-22
end

field_kind(::DensityHirschField) = Int32(0)
field_kind(::MagneticHirschField) = Int32(1)

# init!(mc) has already run init_hopping_matrices + initialize_stack (DQMC.jl:144-148), so the
# exponentials and ranges exist.  Both flavor blocks of a BlockDiagonal hold the same N x N matrix
# (DQMC_interface.jl:288), so block 1 is passed.
dense(H::Hermitian) = Matrix{Float64}(H)
dense(H::Hermitian{<:Any, <:MonteCarlo.BlockDiagonal}) = Matrix{Float64}(parent(H).blocks[1])

function init!(mc::DQMC, u::GPULocalSweep)
    s, p, f = mc.stack, mc.parameters, field(mc)
    rf = Int32[first(r) for r in s.ranges]; rl = Int32[last(r) for r in s.ranges]
    e2, e2i = dense(s.hopping_matrix_exp_squared), dense(s.hopping_matrix_exp_inv_squared)
    eh, ehi = dense(s.hopping_matrix_exp), dense(s.hopping_matrix_exp_inv)
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve rf rl e2 e2i eh ehi begin
        d = DQMCDesc(length(lattice(mc)), p.slices, field_kind(f), 1, length(rf), pointer(rf), pointer(rl),
                     Float64(f.α), pointer(e2), pointer(e2i), pointer(eh), pointer(ehi),
                     p.check_sign_problem, p.check_propagation_error, u.seed, 0, u.device, 0)
        rc = ccall((:dqmc_create, LIB), Int32, (Ref{DQMCDesc}, Ref{Ptr{Cvoid}}), d, ctx)
        rc == 0 || error("dqmc_create: " * unsafe_string(ccall((:dqmc_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    end
    u.ctx = ctx[]
    finalizer(x -> x.ctx != C_NULL && ccall((:dqmc_destroy, LIB), Int32, (Ptr{Cvoid},), x.ctx), u)
    c = conf(f)                                   # Matrix{Int8}(N, M), column-major == ABI layout
    check(u, ccall((:dqmc_set_conf, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int8}), u.ctx, 0, 1, c))
    check(u, ccall((:dqmc_build_stack, LIB), Int32, (Ptr{Cvoid},), u.ctx))   # reverse_build_stack + propagate
    nothing
end

# One full local sweep on the GPU; Julia keeps owning the RNG stream by passing the uniforms.
function update(u::GPULocalSweep, mc::DQMC, model, f)
    N, M = size(conf(f))
    uniforms = rand(Float64, N, 2M)               # [site, step] == C layout [2M][N]
    accepted = Ref{Int64}(0)
    check(u, ccall((:dqmc_sweep, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ref{Int64}),
                   u.ctx, 1, uniforms, accepted))
    # hand the state back so that every existing measurement keeps working (generic.jl:287-288)
    check(u, ccall((:dqmc_get_conf, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int8}), u.ctx, 0, 1, conf(f)))
    G = mc.stack.greens
    if G isa MonteCarlo.BlockDiagonal
        buf = Array{Float64}(undef, N, N, length(G.blocks))
        check(u, ccall((:dqmc_get_greens, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}), u.ctx, 0, 1, buf))
        for b in eachindex(G.blocks); copyto!(G.blocks[b], view(buf, :, :, b)); end
    else
        check(u, ccall((:dqmc_get_greens, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}), u.ctx, 0, 1, G))
    end
    # current_slice = 1, direction = +1 is the invariant after a sweep on both sides
    return accepted[] / (2 * N * M)               # local_updates.jl:82
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
