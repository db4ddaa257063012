# KineticaB200.jl — the Julia-side binding a Kinetica.jl maintainer would add to use
# libkinetica_b200.so as a new solver method.  NOT executed in this repository's CI: the build
# image has no `julia` binary.  It mirrors, call for call, what the Python mirror
# (kinetica.jl_b200/solve.py + _lib.py) does through ctypes, which IS exercised by the tests.
#
# Usage (inside a session that has `using Kinetica`):
#   include("KineticaB200.jl"); using .KineticaB200
#   pars = ODESimulationParams(tspan=(0.0, tf), u0=Dict("C"=>1.0), solver=B200Rodas4(), save_interval=0.1)
#   res  = solve_network(VariableODESolve(pars, conditions, calc), sd, rd)        # drop-in
#   ress = solve_network(B200EnsembleODESolve(pars, [cs1, cs2, ...], calc), sd, rd)
module KineticaB200

using Kinetica
using RecursiveArrayTools: DiffEqArray
import Kinetica: solve_network, AbstractODESolveMethod, ODESimulationParams, ConditionSet,
                 AbstractKineticCalculator, RxFilter, SpeciesData, RxData, ODESolveOutput

export B200Rodas4, B200EnsembleODESolve

const LIB = get(ENV, "KINETICA_B200_LIB", "libkinetica_b200.so")

"Marker usable as `pars.solver`: the batched Rodas4 integrator of libkinetica_b200."
struct B200Rodas4 end

struct B200EnsembleODESolve <: AbstractODESolveMethod
    pars::ODESimulationParams
    conditions::Vector{<:ConditionSet}
    calculator::AbstractKineticCalculator
    filter::RxFilter
end
B200EnsembleODESolve(pars, conditions, calculator) = B200EnsembleODESolve(pars, conditions, calculator, RxFilter())

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

"Flatten RxData (ragged, 1-based) into 0-based CSR (reference src/exploration/network.jl:193-203)."
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

function merge_stops(tstops, saveat, t0, tf)
    ts = filter(t -> t0 <= t <= tf, tstops); sv = filter(t -> t0 <= t <= tf, saveat)
    allt = sort(unique(vcat(ts, sv, tf)))
    flags = Int32[(t in ts ? 1 : 0) | (t in sv ? 2 : 0) for t in allt]
    return allt, flags
end

function solve_ensemble(pars, conds::Vector, calc, sd::SpeciesData, rd::RxData; device=0)
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
    B = length(conds)
    t0, tf = pars.tspan
    si = isnothing(pars.save_interval) ? tf / 1000 : pars.save_interval
    saveat = Kinetica.create_savepoints(t0, tf, si)
    tstops = Kinetica.isstatic(conds[1]) ? Float64[] : Kinetica.get_tstops(conds[1])
    stop_t, flags = merge_stops(tstops, saveat, t0, tf)
    check(h, ccall((:kb2_set_stops, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Int32}),
                   h.ptr, length(stop_t), stop_t, flags))
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
    else
        # any other calculator (ASE, KPM, ...): host table, exactly the reference's discrete design
        B == 1 || error("host-tabulated calculators: single-member solves only")
        k_init = calc(; Kinetica.get_initial_conditions(conds[1])...)
        k_pre = isempty(tstops) ? zeros(0, rd.nr) :
                reduce(hcat, Kinetica.calculate_discrete_rates(conds[1], calc, rd.nr).u)'
        kt = collect(vec(k_pre'))                       # row-major k[s*R + r]
        check(h, ccall((:kb2_set_rate_table, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}),
                       h.ptr, size(k_pre, 1), kt, k_init))
    end
    u0 = Kinetica.make_u0(sd, pars)
    Ns = count(f -> (f & 2) != 0, flags)
    out_u = Vector{Float64}(undef, Ns * sd.n * B); out_umax = Vector{Float64}(undef, sd.n * B)
    status = Vector{Int32}(undef, B); stats = Vector{Int64}(undef, 8B)
    check(h, ccall((:kb2_solve, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Float64, Float64, Float64, Float64, Int64, Int32, Int64,
         Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}),
        h.ptr, B, u0, 0, t0, pars.abstol, pars.reltol, eps(tf), pars.maxiters, pars.ban_negatives, Ns,
        out_u, out_umax, status, stats))
    all(==(0), status) || throw(ErrorException("ODE solution failed."))     # solve_utils.jl:405-411
    U = reshape(out_u, B, sd.n, Ns)                                         # [(s*S + i)*B + b]
    save_t = stop_t[(flags .& 2) .!= 0]
    return [DiffEqArray([U[b, :, s] for s in 1:Ns], save_t) for b in 1:B]
end

function host_prepare(method, sd, rd; copy_network=true)
    sd_a, rd_a = copy_network ? (deepcopy(sd), deepcopy(rd)) : (sd, rd)
    conds = method.conditions isa ConditionSet ? [method.conditions] : method.conditions
    foreach(cs -> Kinetica.solve_variable_conditions!(cs, method.pars), conds)
    mask = Kinetica.get_filter_mask(method.filter, sd_a, rd_a)
    splice!(rd_a, findall(mask))
    Kinetica.setup_network!(sd_a, rd_a, method.calculator)
    Kinetica.apply_low_k_cutoff!(rd_a, method.calculator, method.pars, conds[1])
    return sd_a, rd_a, conds
end

# drop-in: dispatch the reference's own method structs onto the B200 path when pars.solver isa B200Rodas4
function solve_network(method::Union{Kinetica.StaticODESolve, Kinetica.VariableODESolve}, sd::SpeciesData,
                       rd::RxData, ::B200Rodas4; copy_network=true)
    sd_a, rd_a, conds = host_prepare(method, sd, rd; copy_network)
    sol = solve_ensemble(method.pars, conds, method.calculator, sd_a, rd_a)[1]
    return ODESolveOutput(method, sol, sd_a, rd_a)
end

function solve_network(method::B200EnsembleODESolve, sd::SpeciesData, rd::RxData; copy_network=true)
    sd_a, rd_a, conds = host_prepare(method, sd, rd; copy_network)
    sols = solve_ensemble(method.pars, conds, method.calculator, sd_a, rd_a)
    return [ODESolveOutput(Kinetica.VariableODESolve(method.pars, cs, method.calculator, method.filter), s, sd_a, rd_a)
            for (cs, s) in zip(conds, sols)]
end

end # module
