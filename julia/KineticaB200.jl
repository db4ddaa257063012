# KineticaB200.jl — the Julia-side binding a Kinetica.jl maintainer would add to use
# libkinetica_b200.so as a new solver method.
#
# UNVERIFIED: the build image has no `julia` binary, so this file has never been executed.  It
# mirrors, call for call, what the Python mirror (kinetica.jl_b200/solve.py + _lib.py) does through
# ctypes, which IS exercised by the tests; treat it as the worked-out sketch of the binding, not as
# tested code.  File:line references are to Kinetica.jl v0.7.2.
#
# Hooking in.  The reference front doors `solve_network(method::StaticODESolve / VariableODESolve,
# sd, rd; copy_network, return_integrator)` (src/solving/methods.jl:105-130, 330-360) dispatch on
# `Val(split_method)` only, and `pars.solver` is an untyped field, so a new solver cannot be selected
# by dispatch from outside without overwriting those methods.  The intended patch to the reference
# is two lines at the top of each front door (INTEGRATION.md shows it):
#
#     method.pars.solver isa KineticaB200.B200Rodas4 &&
#         return KineticaB200.b200_solve_network(method, sd, rd; copy_network=copy_network)
#
# `B200EnsembleODESolve` is a new method type and gets its own `solve_network` method here.
#
# Usage (inside a session that has `using Kinetica`):
#   include("KineticaB200.jl"); using .KineticaB200
#   pars = ODESimulationParams(tspan=(0.0, tf), u0=Dict("C"=>1.0), solver=B200Rodas4())
#   res  = b200_solve_network(VariableODESolve(pars, conditions, calc), sd, rd)
#   ress = solve_network(B200EnsembleODESolve(pars, [cs1, cs2, ...], calc), sd, rd)
module KineticaB200

using Kinetica
using RecursiveArrayTools: DiffEqArray
import Kinetica: solve_network, AbstractODESolveMethod, ODESimulationParams, ConditionSet,
                 AbstractKineticCalculator, RxFilter, SpeciesData, RxData, ODESolveOutput

export B200Rodas4, B200EnsembleODESolve, b200_solve_network

const LIB = get(ENV, "KINETICA_B200_LIB", "libkinetica_b200.so")
const STOP_RATE, STOP_SAVE, STOP_CHUNK = Int32(1), Int32(2), Int32(4)

"Marker usable as `pars.solver`: the batched Rodas4 integrator of libkinetica_b200."
struct B200Rodas4 end

"One `ConditionSet` per ensemble member; network, calculator and parameters are shared."
struct B200EnsembleODESolve <: AbstractODESolveMethod
    pars::ODESimulationParams
    conditions::Vector{<:ConditionSet}
    calculator::AbstractKineticCalculator
    filter::RxFilter
    function B200EnsembleODESolve(pars, conditions, calculator, filter=RxFilter())
        isempty(conditions) && throw(ArgumentError("An ensemble needs at least one ConditionSet."))
        for cs in conditions
            Kinetica.has_conditions(calculator, cs.symbols) ||
                throw(ArgumentError("Calculator does not support all of the provided conditions."))
            (Kinetica.isstatic(cs) || cs.discrete_updates) ||
                throw(ArgumentError("The B200 ensemble path needs discrete rate updates (ts_update) for variable conditions."))
        end
        return new(pars, conditions, calculator, filter)
    end
end

mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(device::Integer=0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:kb2_create, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, out)
        rc == 0 || error("kb2_create failed ($rc): no usable CUDA device; there is no CPU fallback")
        h = new(out[])
        finalizer(x -> ccall((:kb2_destroy, LIB), Int32, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end

function check(h::Handle, rc::Int32)
    rc == 0 && return
    msg = unsafe_string(ccall((:kb2_last_error, LIB), Cstring, (Ptr{Cvoid},), h.ptr))
    throw(ErrorException("libkinetica_b200: $msg"))
end

"Flatten RxData (ragged, 1-based) into 0-based CSR (src/exploration/network.jl:193-203)."
function flatten(ids::Vector{Vector{Int}}, nus::Vector{Vector{Int}})
    ptr = Int64[0]; idx = Int64[]; nu = Int64[]
    for (r, s) in zip(ids, nus)
        append!(idx, r .- 1); append!(nu, s); push!(ptr, length(idx))
    end
    isempty(idx) && (push!(idx, 0); push!(nu, 0))
    return ptr, idx, nu
end

# profile -> (kind, params[16])   (include/kinetica_b200.h KB2_PROFILE_*)
desc(p::Kinetica.StaticConditionProfile) = (Int32(0), vcat(Float64(p.value), zeros(15)))
desc(p::Union{Kinetica.NullDirectProfile, Kinetica.NullGradientProfile}) = (Int32(1), vcat(p.X_start, zeros(15)))
desc(p::Kinetica.LinearDirectProfile) = (Int32(2), vcat(p.rate, p.X_start, p.X_end, p.t_end, zeros(12)))
desc(p::Kinetica.LinearGradientProfile) = (Int32(3), vcat(p.rate, p.X_start, p.X_end, p.t_end, zeros(12)))
desc(p::Kinetica.DoubleRampGradientProfile) =
    (Int32(4), vcat(p.X_start, p.rate1, p.rate2, p.t_startr1, p.t_endr1, p.t_startr2, p.t_endr2, p.t_blend, zeros(8)))

"""
Chunk boundaries and save times of a chunkwise solve (methods.jl:214-222, 757-765): local save points
`0:save_interval:chunkstep` (save_interval defaults to the chunk step), global time = local +
nc*chunkstep, `(length(saveat_local)-1)*n_chunks + 1` points.
"""
function chunk_grid(pars)
    step = pars.solve_chunkstep
    n_chunks = Int(pars.tspan[2] / step)
    si = isnothing(pars.save_interval) ? step : pars.save_interval
    loc = collect(0.0:si:step)
    save = Float64[loc[i] + nc * step for nc in 0:n_chunks-1 for i in 1:length(loc)-1]
    push!(save, loc[end] + (n_chunks - 1) * step)
    return Float64[nc * step for nc in 1:n_chunks-1], save
end

"Merged stop list of one member; times that differ only in the last bits are one stop (the tstop's value is kept)."
function merge_stops(tstops, saveat, chunks, t0, tf)
    tol = 1e-12 * max(1.0, abs(tf))
    ev = vcat([(t, STOP_RATE, 0) for t in tstops if t0 <= t <= tf],
              [(t, STOP_SAVE, 1) for t in saveat if t0 <= t <= tf],
              [(t, STOP_CHUNK, 1) for t in chunks if t0 < t < tf], [(tf, Int32(0), 1)])
    sort!(ev; by = e -> (e[1], e[3]))
    ts = Float64[]; fl = Int32[]
    for (t, f, _) in ev
        if !isempty(ts) && t - ts[end] <= tol
            fl[end] |= f
            f == STOP_RATE && (ts[end] = t)
        else
            push!(ts, t); push!(fl, f)
        end
    end
    abs(ts[end] - tf) <= tol && (ts[end] = tf)
    return ts, fl
end

"""
The device solve of `conds` (one member each) on the network `sd`/`rd`: returns per-member save
arrays, maxima, status words and the save times.  `pars` carries tolerances etc.
"""
function solve_members(h::Handle, pars, conds::Vector, calc, sd::SpeciesData, rd::RxData)
    B = length(conds)
    t0, tf = pars.tspan
    if pars.solve_chunks
        chunks, saveat = chunk_grid(pars)
    else
        chunks = Float64[]
        si = isnothing(pars.save_interval) ? tf / 1000 : pars.save_interval   # the reference saves every step for `nothing`
        saveat = Kinetica.create_savepoints(t0, tf, si)
    end
    lists = [merge_stops(Kinetica.isstatic(cs) ? Float64[] : Kinetica.get_tstops(cs), saveat, chunks, t0, tf) for cs in conds]
    nmax = maximum(length(l[1]) for l in lists)
    stop_t = zeros(nmax, B); flags = zeros(Int32, nmax, B); counts = Int32[length(l[1]) for l in lists]
    for (b, (t, f)) in enumerate(lists)
        stop_t[1:length(t), b] .= t; flags[1:length(f), b] .= f       # column b = member b: C layout [b*nmax + s]
    end
    check(h, ccall((:kb2_set_member_stops, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}),
                   h.ptr, B, nmax, counts, stop_t, flags))
    if calc isa Kinetica.PrecalculatedArrheniusCalculator
        kmax = isnothing(calc.k_max) ? NaN : Float64(calc.k_max)
        check(h, ccall((:kb2_set_arrhenius, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64),
            h.ptr, calc.A, calc.Ea, C_NULL, kmax, calc.t_mult))
        kinds = Int32[]; params = Float64[]
        for cs in conds
            k, p = desc(Kinetica.get_profile(cs, :T)); push!(kinds, k); append!(params, p)
        end
        check(h, ccall((:kb2_set_profiles, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Float64}),
                       h.ptr, B, kinds, params))
        # the reference evaluates the INTERPOLATED profile solution at every tstop (solve_utils.jl:101-104):
        # ship those values (NaN = not a rate stop)
        Ttab = fill(NaN, nmax, B)
        for (b, cs) in enumerate(conds)
            Kinetica.isstatic(cs) && continue
            vc = Kinetica.get_variable_conditions(cs)
            haskey(vc, :T) || continue
            for s in 1:counts[b]
                (flags[s, b] & STOP_RATE) != 0 && (Ttab[s, b] = vc[:T](stop_t[s, b])[1])
            end
        end
        check(h, ccall((:kb2_set_T_table, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}), h.ptr, B, nmax, Ttab))
    else
        # any other calculator (ASE, KPM, ...): host table, exactly the reference's discrete design
        B == 1 || error("host-tabulated calculators: single-member solves only")
        k_init = calc(; Kinetica.get_initial_conditions(conds[1])...)
        rate_t = [stop_t[s, 1] for s in 1:counts[1] if (flags[s, 1] & STOP_RATE) != 0]
        k_pre = isempty(rate_t) ? zeros(0, rd.nr) :
                reduce(hcat, Kinetica.calculate_discrete_rates(conds[1], calc, rd.nr).u)'
        kt = collect(vec(k_pre'))                       # row-major k[s*R + r]
        check(h, ccall((:kb2_set_rate_table, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}),
                       h.ptr, size(k_pre, 1), kt, k_init))
        check(h, ccall((:kb2_set_T_table, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}), h.ptr, 0, 0, C_NULL))
    end
    # chunkwise: adaptive_solve! per chunk runs on the device
    check(h, ccall((:kb2_set_chunking, LIB), Int32, (Ptr{Cvoid}, Int32, Int32), h.ptr,
                   Int32(pars.solve_chunks && pars.adaptive_tols), Int32(pars.update_tols)))
    u0 = Kinetica.make_u0(sd, pars)
    Ns = count(f -> (f & STOP_SAVE) != 0, @view flags[1:counts[1], 1])
    out_u = Vector{Float64}(undef, Ns * sd.n * B); out_umax = Vector{Float64}(undef, sd.n * B)
    status = Vector{Int32}(undef, B); stats = Vector{Int64}(undef, 8B)
    dtmin = eps(pars.solve_chunks ? pars.solve_chunkstep : tf)          # methods.jl:164 / :231
    check(h, ccall((:kb2_solve, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Float64, Float64, Float64, Float64, Int64, Int32, Int64,
         Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}),
        h.ptr, B, u0, 0, t0, pars.abstol, pars.reltol, dtmin, pars.maxiters, pars.ban_negatives, Ns,
        out_u, out_umax, status, stats))
    U = reshape(out_u, B, sd.n, Ns)                                     # C layout [(s*S + i)*B + b]
    save_t = [stop_t[s, 1] for s in 1:counts[1] if (flags[s, 1] & STOP_SAVE) != 0]
    return U, reshape(out_umax, B, sd.n), status, save_t
end

"adaptive_solve! (solve_utils.jl:376-424) per member: only the members that failed are solved again."
function solve_with_retry(h, pars, conds, calc, sd, rd)
    p = deepcopy(pars)
    B = length(conds)
    todo = collect(1:B)
    U = nothing; umax = nothing; save_t = nothing
    mintol = eps(Float64)
    iters = 0
    while true
        iters += 1
        Us, ms, st, save_t = solve_members(h, p, conds[todo], calc, sd, rd)
        if isnothing(U)
            U, umax = Us, ms
        else
            U[todo, :, :] .= Us; umax[todo, :] .= ms
        end
        bad = findall(!=(0), st)
        if isempty(bad)
            if pars.update_tols && p.abstol != pars.abstol
                pars.abstol = p.abstol; pars.reltol = p.reltol
            end
            return U, umax, save_t
        end
        # chunkwise solves have already repeated the failed chunk on the device
        if pars.solve_chunks || !pars.adaptive_tols || iters >= 5 || p.abstol / 10 <= mintol || p.reltol / 10 <= mintol
            throw(ErrorException("ODE solution failed."))               # solve_utils.jl:405-411
        end
        todo = todo[bad]
        p.abstol /= 10; p.reltol /= 10
    end
end

"apply_low_k_cutoff! (solve_utils.jl:213-245) with the ENSEMBLE-wide maximum rates."
function apply_low_k_cutoff_ensemble!(rd, calc, pars, conds)
    pars.low_k_cutoff == :none && return 0
    k_cutoff = pars.low_k_cutoff == :auto ? pars.reltol / pars.tspan[end] : Float64(pars.low_k_cutoff)
    max_rates = reduce((a, b) -> max.(a, b), [Kinetica.get_max_rates(cs, calc) for cs in conds]) .* pars.low_k_maxconc^2
    low = findall(<(k_cutoff), max_rates)
    splice!(rd, calc, low)
    return length(low)
end

function host_prepare(method, sd, rd; copy_network=true)
    sd_a, rd_a = copy_network ? (deepcopy(sd), deepcopy(rd)) : (sd, rd)
    conds = method.conditions isa ConditionSet ? [method.conditions] : method.conditions
    foreach(cs -> Kinetica.solve_variable_conditions!(cs, method.pars), conds)
    mask = Kinetica.get_filter_mask(method.filter, sd_a, rd_a)
    splice!(rd_a, findall(mask))
    Kinetica.setup_network!(sd_a, rd_a, method.calculator)
    apply_low_k_cutoff_ensemble!(rd_a, method.calculator, method.pars, conds)
    return sd_a, rd_a, conds
end

function device_network(sd, rd; device=0)
    h = Handle(device)
    rp, ri, rn = flatten(rd.id_reacs, rd.stoic_reacs)
    pp, pi_, pn = flatten(rd.id_prods, rd.stoic_prods)
    GC.@preserve rp ri rn pp pi_ pn begin
        check(h, ccall((:kb2_set_network, LIB), Int32,
            (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
            h.ptr, sd.n, rd.nr, rp, ri, rn, pp, pi_, pn))
    end
    nnz = Ref{Int64}(0); nlu = Ref{Int64}(0); nf = Ref{Int64}(0)
    check(h, ccall((:kb2_symbolic, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Int64}, Ref{Int64}, Ref{Int64}),
                   h.ptr, 4, nnz, nlu, nf))
    return h
end

"`res.sol_k` of a discrete solve: DiffEqArray(k_precalc, tstops) (solve_utils.jl:91-109, analysis/io.jl:36-38)."
rate_solution(cs, calc, rd) =
    (Kinetica.isstatic(cs) || !cs.discrete_updates) ? nothing : Kinetica.calculate_discrete_rates(cs, calc, rd.nr)

"Seven-field constructor of analysis/io.jl:3-11; `sol` is the DiffEqArray `load_output` also rebuilds (io.jl:229)."
make_output(sd, rd, U, b, save_t, cs, calc, pars) =
    ODESolveOutput(sd, rd, DiffEqArray([U[b, :, s] for s in 1:size(U, 3)], save_t), rate_solution(cs, calc, rd),
                   nothing, pars, cs)

"Front door for `StaticODESolve` / `VariableODESolve` with `pars.solver isa B200Rodas4` (see the header)."
function b200_solve_network(method::Union{Kinetica.StaticODESolve, Kinetica.VariableODESolve}, sd::SpeciesData,
                            rd::RxData; copy_network=true, device=0)
    (method isa Kinetica.VariableODESolve && !method.conditions.discrete_updates) &&
        throw(ArgumentError("continuous rate updates go through kb2_set_continuous; pass ts_update for discrete updates"))
    sd_a, rd_a, conds = host_prepare(method, sd, rd; copy_network)
    h = device_network(sd_a, rd_a; device)
    U, _, save_t = solve_with_retry(h, method.pars, conds, method.calculator, sd_a, rd_a)
    return make_output(sd_a, rd_a, U, 1, save_t, conds[1], method.calculator, method.pars)
end

function solve_network(method::B200EnsembleODESolve, sd::SpeciesData, rd::RxData; copy_network=true, device=0)
    sd_a, rd_a, conds = host_prepare(method, sd, rd; copy_network)
    h = device_network(sd_a, rd_a; device)
    U, _, save_t = solve_with_retry(h, method.pars, conds, method.calculator, sd_a, rd_a)
    return [make_output(sd_a, rd_a, U, b, save_t, cs, method.calculator, method.pars) for (b, cs) in enumerate(conds)]
end

end # module
