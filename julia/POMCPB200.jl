# POMCPB200.jl — ccall binding of libpomcp_b200.so behind POMCP.jl's own API surface.
#
# UNVERIFIED IN THIS ENVIRONMENT: no Julia toolchain exists in the build container or on the GPU
# box.  This file shows the reference-side binding a maintainer would add; the tested drop-in
# boundary is the C ABI itself (include/pomcp_b200.h), exercised through the Python mirror
# pomcp.jl_b200/api.py, whose structure this file follows one to one.
#
# It keeps the reference names: POMCPSolver / solve / action / updater / initialize_belief / update
# (src/POMCP.jl:20-67), the exceptions of src/exceptions.jl and the depletion ErrorException of
# src/particle_filter.jl:94-101.  Written for Julia >= 1.6 + POMDPs.jl >= 0.9.
module POMCPB200

import POMDPs
import POMDPs: solve, action, updater, update, initialize_belief

const LIB = get(ENV, "POMCP_B200_LIB", "libpomcp_b200")

# ---- POD blocks (must match include/pomcp_b200.h byte for byte)
struct CProblem
    problem_id::Int32; _pad0::Int32
    discount::Float64
    tiger_r_listen::Float64; tiger_r_findtiger::Float64; tiger_r_escapetiger::Float64; tiger_p_listen_correctly::Float64
    baby_r_feed::Float64; baby_r_hungry::Float64; baby_p_become_hungry::Float64
    baby_p_cry_when_hungry::Float64; baby_p_cry_when_not_hungry::Float64
    rs_n::Int32; rs_k::Int32
    rs_rock_x::NTuple{16,Int32}; rs_rock_y::NTuple{16,Int32}
    rs_start_x::Int32; rs_start_y::Int32
    rs_half_eff_dist::Float64; rs_r_good::Float64; rs_r_bad::Float64; rs_r_exit::Float64
    ld_correct_r::Float64; ld_incorrect_r::Float64; ld_init_mean::Float64; ld_init_std::Float64
end

struct CSolverParams
    eps::Float64; max_depth::Int64; c::Float64; tree_queries::Int64; seed::UInt64
    init_V::Float64; init_N::Int64
    num_sparse_actions::Int32; estimator::Int32; estimator_value::Float64
    belief_mode::Int32; search_mode::Int32
    max_nodes::Int64; max_log::Int64
    inflight_min::Int32; inflight_max::Int32; inflight_ramp_div::Int32; rollouts_per_leaf::Int32
    rank::Int32
    dpw_observation::Int32; k_observation::Float64; alpha_observation::Float64     # src/constructor.jl:40-73
    dpw_action::Int32; dpw_solver::Int32; k_action::Float64; alpha_action::Float64
    n_trees::Int32; _pad3::Int32                                                    # ABI version 2
end

struct CEpisodeParams                                                               # pomcp_run_episodes
    max_steps::Int32; initial_state::Int32; sim_eps::Float64; env_seed::UInt64
    sir_particles::Int64; initial_p1::Float64
end

# ---- exceptions (src/exceptions.jl:1-26)
abstract type NoDecision <: Exception end
struct AllSamplesTerminal <: NoDecision
    belief
end
struct ExceptionRethrow end
default_action(::ExceptionRethrow, belief, ex) = rethrow(ex)
default_action(f::Function, belief, ex) = f(belief, ex)
default_action(p::POMDPs.Policy, belief, ex) = action(p, belief)
default_action(a, belief, ex) = a

particle_depletion_error() = error("""
    POMCP.jl: Particle Depletion! To fix this, you have three options:
          1) use more tree_queries (will only work for very small problems)
          2) implement a ParticleReinvigorator with reinvigorate!() and handle_unseen_observation()
          3) implement a more advanced updater for the agent (POMCP can use any
             belief/state distribution that supports rand())
    """)

function check(rc::Int32, h::Ptr{Cvoid}=C_NULL, belief=nothing)
    rc == 0 && return
    rc == 1 && throw(AllSamplesTerminal(belief))
    rc == 2 && particle_depletion_error()
    msg = unsafe_string(ccall((:pomcp_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    error("pomcp_b200 status $rc: $msg")
end

# ---- solver / planner (src/constructor.jl:8-31, src/POMCP.jl:137-152,260-268)
Base.@kwdef mutable struct POMCPSolver <: POMDPs.Solver
    eps::Float64 = 0.01
    max_depth::Int = typemax(Int)
    c::Float64 = 1.0
    tree_queries::Int = 100
    rng::UInt64 = 0                 # Philox key instead of an AbstractRNG
    init_V::Float64 = 0.0
    init_N::Int = 0
    num_sparse_actions::Int = 0
    default_action = ExceptionRethrow()
    search_mode::Symbol = :batched  # :exact for bit-parity
    device::Int = 0
end

mutable struct POMCPPlanner{P} <: POMDPs.Policy
    problem::P
    solver::POMCPSolver
    handle::Ptr{Cvoid}
    generation::Int
end

struct BeliefNode            # belief == search root (src/updater.jl:20-26)
    planner::POMCPPlanner
    generation::Int
end

"Problem types opt in by defining `c_problem(p)::CProblem` (Tiger, Baby, RockSample, LightDark1D)."
function c_problem end

function solve(solver::POMCPSolver, pomdp::POMDPs.POMDP)   # src/solver.jl:30-32
    sp = CSolverParams(solver.eps, solver.max_depth, solver.c, solver.tree_queries, solver.rng,
                       solver.init_V, solver.init_N, solver.num_sparse_actions, 0, 0.0,
                       0, solver.search_mode == :exact ? 0 : 1, 0, 0, 0, 0, 0, 0, 0,
                       0, 10.0, 0.5,            # dpw_observation off; k/alpha defaults of src/constructor.jl:60-63
                       0, 0, 10.0, 0.5,         # dpw_action off, dpw_solver off
                       1, 0)                    # n_trees (ABI version 2)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:pomcp_create, LIB), Int32, (Ref{CProblem}, Ref{CSolverParams}, Int32, Ref{Ptr{Cvoid}}),
               c_problem(pomdp), sp, solver.device, h)
    check(rc)
    pl = POMCPPlanner(pomdp, solver, h[], 0)
    finalizer(p -> ccall((:pomcp_destroy, LIB), Int32, (Ptr{Cvoid},), p.handle), pl)
    return pl
end

function make_node(pl::POMCPPlanner, b)                     # src/solver.jl:276-281
    if b isa BeliefNode
        b.generation == pl.generation || error("stale BeliefNode")
        return b
    elseif b isa AbstractVector                              # particle collection
        GC.@preserve b check(ccall((:pomcp_set_belief_particles, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                                   pl.handle, pointer(b), length(b)), pl.handle)
    else                                                     # initial_state_distribution(pomdp)
        check(ccall((:pomcp_set_belief_initial, LIB), Int32, (Ptr{Cvoid},), pl.handle), pl.handle)
    end
    pl.generation += 1
    return BeliefNode(pl, pl.generation)
end

function action(pl::POMCPPlanner, belief)                   # src/solver.jl:16-23
    try
        make_node(pl, belief)
        a = Ref{Int32}(-1)
        nA = length(POMDPs.actions(pl.problem))
        N = zeros(Int64, nA); V = zeros(Float64, nA)
        rc = ccall((:pomcp_search, LIB), Int32, (Ptr{Cvoid}, Int64, Ref{Int32}, Ptr{Int64}, Ptr{Float64}, Int32),
                   pl.handle, pl.solver.tree_queries, a, N, V, nA)
        check(rc, pl.handle, belief)
        return POMDPs.actions(pl.problem)[a[] + 1]
    catch ex
        return default_action(pl.solver.default_action, belief, ex)
    end
end

struct RootUpdater <: POMDPs.Updater                        # src/updater.jl:8-10
    planner::POMCPPlanner
end
updater(pl::POMCPPlanner) = RootUpdater(pl)                 # src/updater.jl:41
initialize_belief(up::RootUpdater, b) = make_node(up.planner, b)   # src/updater.jl:45-47

obs_bits(o::Float64) = reinterpret(UInt64, o)
obs_bits(o::Bool) = UInt64(o)
obs_bits(o::Integer) = UInt64(o)

function update(up::RootUpdater, b_old::BeliefNode, a, o)  # src/updater.jl:13-27
    pl = up.planner
    ai = Int32(findfirst(==(a), POMDPs.actions(pl.problem)) - 1)
    check(ccall((:pomcp_update_unweighted, LIB), Int32, (Ptr{Cvoid}, Int32, UInt64), pl.handle, ai, obs_bits(o)),
          pl.handle)
    pl.generation += 1
    return BeliefNode(pl, pl.generation)
end

"Device-side reinvigorator: pass as `node_belief_updater`; both hooks of src/particle_filter.jl:37-73 then run
inside pomcp_update_reinvigorated (SIR posterior of the old root belief, at least `n` particles)."
mutable struct SIRReinvigorator <: ParticleReinvigorator
    n::Int
    seed::UInt64
    step::Int
    last_outcome::Int32                                     # 0 plain re-root, 1 topped up, 2 rebuilt
end
SIRReinvigorator(n; seed=0) = SIRReinvigorator(n, UInt64(seed), 0, Int32(0))

function update(up::RootUpdater, b_old::BeliefNode, a, o, r::SIRReinvigorator)
    pl = up.planner
    ai = Int32(findfirst(==(a), POMDPs.actions(pl.problem)) - 1)
    r.step += 1
    out = Ref{Int32}(0)
    check(ccall((:pomcp_update_reinvigorated, LIB), Int32, (Ptr{Cvoid}, Int32, UInt64, UInt64, Int64, Ref{Int32}),
                pl.handle, ai, obs_bits(o), r.seed * 0x9E3779B97F4A7C15 + UInt64(r.step), r.n, out), pl.handle)
    r.last_outcome = out[]
    pl.generation += 1
    return BeliefNode(pl, pl.generation)
end

"SIR particle filter on the device (ParticleFilters.SIRParticleFilter, test/runtests.jl:57)."
struct SIRParticleFilter <: POMDPs.Updater
    planner::POMCPPlanner
    n::Int
    seed::UInt64
end
function update(up::SIRParticleFilter, particles::AbstractVector, a, o)
    pl = up.planner
    ai = Int32(findfirst(==(a), POMDPs.actions(pl.problem)) - 1)
    GC.@preserve particles check(ccall((:pomcp_set_belief_particles, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                                       pl.handle, pointer(particles), length(particles)), pl.handle)
    check(ccall((:pomcp_pf_update_sir, LIB), Int32, (Ptr{Cvoid}, Int32, UInt64, UInt64, Int64),
                pl.handle, ai, obs_bits(o), up.seed, up.n), pl.handle)
    out = similar(particles, up.n)
    n = Ref{Int64}(0)
    GC.@preserve out check(ccall((:pomcp_get_belief_particles, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Int64}),
                                 pl.handle, pointer(out), up.n, n), pl.handle)
    pl.generation += 1
    return out
end

export POMCPSolver, POMCPPlanner, RootUpdater, SIRParticleFilter, BeliefNode, NoDecision, AllSamplesTerminal,
       ExceptionRethrow, solve, action, updater, update, initialize_belief

end # module
