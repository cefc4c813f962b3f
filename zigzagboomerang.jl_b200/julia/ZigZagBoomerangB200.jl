# ZigZagBoomerangB200.jl -- thin Julia wrapper over libzzb200.so (include/zzb200.h).
#
# Adds GPU methods to the reference's own `spdmp` / `pdmp` / `sspdmp` generic functions (ZigZag, LocalBound, sticky ZigZag,
# FactBoomerang): they are selected by dispatch when
# the "gradient" argument is a `GaussianPotential` / `LogisticSubsampled` descriptor instead of a closure, and return exactly what the
# reference returns: `Ξ::FactTrace, (t, x, θ), (acc, num), c` (src/sfact.jl:211).
#
# NOTE: Julia is not installed in the build image, so this file is UNTESTED there; it is kept in lock-step with the
# ctypes binding zigzagboomerang.jl_b200/_capi.py, which exercises the same entry points in the test-suite.
module ZigZagBoomerangB200

using ZigZagBoomerang
using ZigZagBoomerang: ZigZag, FactBoomerang, LocalBound, FactTrace, Trace, Seed
using SparseArrays
import ZigZagBoomerang: spdmp, pdmp, sspdmp, sspdmp2, sspdmp3, sspdmp4

const libzzb200 = get(ENV, "ZZB200_LIB", joinpath(@__DIR__, "..", "libzzb200.so"))
const cubin = get(ENV, "ZZB200_CUBIN", joinpath(@__DIR__, "..", "zzb200_kernels.cubin"))

const ZZB_E_BOUND = Int32(3)
const ZZB_FLAG_NO_TRACE = UInt32(1)
const ZZB_FLAG_LOCAL_BOUND = UInt32(2)
const ZZB_FLAG_STICKY_REVERSIBLE = UInt32(16)
const ZZB_FLAG_STICKY_STRONG_UB = UInt32(32)
const ZZB_FLAG_STICKY_ZZ = UInt32(128)

"""
    GaussianPotential(Γ, h = nothing)

Target descriptor standing in for the closure `∇ϕ(x, i) = idot(Γ, i, x) - h[i]`; also callable like it, so the same
object works with the reference's CPU `spdmp`.
"""
struct GaussianPotential{T<:SparseMatrixCSC{Float64,Int}, H}
    Γ::T
    h::H
end
GaussianPotential(Γ) = GaussianPotential(Γ, nothing)
(g::GaussianPotential)(x, i, args...) = ZigZagBoomerang.idot(g.Γ, i, x) - (g.h === nothing ? 0.0 : g.h[i])

function lasterror()
    buf = Vector{UInt8}(undef, 1024)
    ccall((:zzb_last_error, libzzb200), Int32, (Ptr{UInt8}, Int64), buf, 1024)
    unsafe_string(pointer(buf))
end

function check(st::Int32)
    st == 0 && return
    st == ZZB_E_BOUND && error("Tuning parameter `c` too small.")   # same message as src/sfact.jl:124
    error("zzb200 error $st: " * lasterror())
end

const initialised = Ref(false)
function init(device::Integer = 0)
    initialised[] && return
    ids = Int32[device]
    check(ccall((:zzb_init, libzzb200), Int32, (Int32, Ptr{Int32}, Cstring), 1, ids, cubin))
    initialised[] = true
end
"""
    LogisticSubsampled(A, At, y, ny, μ, γ0 = 0.01, k = 10)

Target descriptor standing in for the closure `∇ϕmoving(t, x, θ, i, t′, F, A, At, μ, y, ny, k)` of `scripts/logistic.jl:78-107`
together with the trailing arguments `SelfMoving(), A, At, μ, y, ny, k` the script forwards to it (`:167`): the subsampled
partial derivative of the logistic-regression potential with a control variate at `μ`.  Calling it evaluates the full-data
partial derivative (`∇ϕ(x, i, A, At, y, ny)`, `:104`), so the same object also works with the reference's CPU `spdmp`.
"""
struct LogisticSubsampled{T<:SparseMatrixCSC{Float64,Int}}
    A::T
    At::T
    y::Vector{Float64}
    ny::Vector{Float64}
    μ::Vector{Float64}
    γ0::Float64
    k::Int
end
LogisticSubsampled(A, At, y, ny, μ, γ0 = 0.01, k = 10) =
    LogisticSubsampled(A, At, Vector{Float64}(y), Vector{Float64}(ny), Vector{Float64}(μ), Float64(γ0), Int(k))
function (g::LogisticSubsampled)(x, i, args...)
    sig(u) = inv(one(u) + exp(-u))
    rows, vals = rowvals(g.A), nonzeros(g.A)
    s = 0.0
    for p in nzrange(g.A, i)
        u = ZigZagBoomerang.idot(g.At, rows[p], x)
        s += vals[p]*g.y[rows[p]]*sig(-u) - vals[p]*g.ny[rows[p]]*sig(u)
    end
    g.γ0*x[i] - s
end

# zzb_problem_create_logistic: Julia's CSC arrays of A and At as they are
function create_problem(prob, ∇ϕ::LogisticSubsampled, F, μ, d)
    Γb = F.Γ
    check(ccall((:zzb_problem_create_logistic, libzzb200), Int32,
                (Ref{Ptr{Cvoid}}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
                prob, d, size(∇ϕ.A, 1), ∇ϕ.A.colptr, ∇ϕ.A.rowval, ∇ϕ.A.nzval, ∇ϕ.At.colptr, ∇ϕ.At.rowval, ∇ϕ.At.nzval,
                ∇ϕ.y, ∇ϕ.ny, ∇ϕ.μ, ∇ϕ.γ0, ∇ϕ.k, Γb.colptr, Γb.rowval, Γb.nzval, μ))
end

# spdmp(∇ϕmoving, t0, x0, θ0, T, c, Zdrop, SelfMoving(), A, At, μ, y, ny, k; adapt, factor) (scripts/logistic.jl:167) becomes
# spdmp(LogisticSubsampled(A, At, y, ny, μ, γ0, k), t0, x0, θ0, T, c, Zdrop; adapt, factor); trailing arguments are ignored
function spdmp(∇ϕ::LogisticSubsampled, t0, x0, θ0, T, c, G::Union{ZigZagBoomerang.All,ZigZagBoomerang.Matched}, F::ZigZag, args...;
               factor = 1.8, adapt = false, adaptscale = false, progress = false, progress_stops = 20, seed = Seed())
    F.λref == 0 || error("ZigZag refreshments (λref > 0) are not implemented on the device path")
    Ξ, u, an, cv = device_run(:zigzag, ∇ϕ, t0, x0, θ0, T, c, F; factor = factor, adapt = adapt, seed = seed)
    c .= cv
    Ξ, u, an, c
end
spdmp(∇ϕ::LogisticSubsampled, t0, x0, θ0, T, c, F::ZigZag, args...; kargs...) =
    spdmp(∇ϕ, t0, x0, θ0, T, c, ZigZagBoomerang.Matched(), F, args...; kargs...)

# One device run: problem from (∇ϕ, F), then the one-call entry point selected by `kind`, then the results in the
# reference's return shape.  `kind`: :zigzag (zzb_spdmp_run, flags = 0 or ZZB_FLAG_LOCAL_BOUND), :sticky (zzb_sspdmp_run),
# :boomerang (zzb_spdmp_boomerang_run).
function device_run(kind::Symbol, ∇ϕ::Union{GaussianPotential,LogisticSubsampled}, t0, x0, θ0, T, c, F; κ = nothing, flags = UInt32(0),
                    factor = 1.8, adapt = false, seed = Seed())
    init()
    d = length(x0)
    logistic = ∇ϕ isa LogisticSubsampled
    Γt, Γb = logistic ? F.Γ : ∇ϕ.Γ, F.Γ
    μ = Vector{Float64}(F.μ)
    x0v, θ0v, cv = Vector{Float64}(x0), Vector{Float64}(θ0), Vector{Float64}(c)
    prob = Ref{Ptr{Cvoid}}(C_NULL)
    run = Ref{Ptr{Cvoid}}(C_NULL)
    h = (logistic || ∇ϕ.h === nothing) ? Ptr{Float64}(C_NULL) : pointer(∇ϕ.h)
    sd = UInt64[seed[1], seed[2]]
    κv = κ === nothing ? Float64[] : Vector{Float64}(κ isa Number ? fill(κ, d) : κ)
    σv = (kind === :boomerang || kind === :refresh) ? Vector{Float64}(F.σ) : Float64[]
    GC.@preserve Γt Γb μ x0v θ0v cv sd κv σv ∇ϕ begin
        if logistic
            create_problem(prob, ∇ϕ, F, μ, d)
        elseif flags & ZZB_FLAG_LOCAL_BOUND != 0     # the bound comes from the target itself (src/local.jl:2-6): bnd_* = NULL
            check(ccall((:zzb_problem_create_gaussian, libzzb200), Int32,
                        (Ref{Ptr{Cvoid}}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64},
                         Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
                        prob, d, Γt.colptr, Γt.rowval, Γt.nzval, h, C_NULL, C_NULL, C_NULL, C_NULL))
        else
            check(ccall((:zzb_problem_create_gaussian, libzzb200), Int32,
                        (Ref{Ptr{Cvoid}}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64},
                         Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
                        prob, d, Γt.colptr, Γt.rowval, Γt.nzval, h, Γb.colptr, Γb.rowval, Γb.nzval, μ))
        end
        st = if kind === :sticky          # sspdmp, src/ss_fact.jl:159-217 (with adapt: :132-136)
            ccall((:zzb_sspdmp_adapt_run, libzzb200), Int32,
                  (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{UInt64}, Int32, Float64, UInt32,
                   Ref{Ptr{Cvoid}}), prob[], t0, x0v, θ0v, T, cv, κv, sd, adapt, factor, flags, run)
        elseif kind === :refresh          # spdmp with Z.λref > 0 (hasrefresh, src/fact_samplers.jl:19): refresh branch src/sfact.jl:78-114
            ccall((:zzb_spdmp_refresh_run, libzzb200), Int32,
                  (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}, Float64,
                   Ptr{UInt64}, Int32, Float64, UInt32, Ref{Ptr{Cvoid}}),
                  prob[], t0, x0v, θ0v, T, cv, σv, F.λref, sd, adapt, factor, flags, run)
        elseif kind === :boomerang        # spdmp with F::FactBoomerang, src/sfact.jl:29-48,73-145
            ccall((:zzb_spdmp_boomerang_run, libzzb200), Int32,
                  (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}, Float64, Float64,
                   Ptr{UInt64}, Int32, Float64, UInt32, Ref{Ptr{Cvoid}}),
                  prob[], t0, x0v, θ0v, T, cv, σv, F.λref, F.ρ, sd, adapt, factor, flags, run)
        else                              # spdmp / pdmp with F::ZigZag, src/sfact.jl:162-214,236 (LocalBound: src/local.jl:95-149)
            ccall((:zzb_spdmp_run, libzzb200), Int32,
                  (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{UInt64}, Int32, Float64,
                   UInt32, Ref{Ptr{Cvoid}}), prob[], t0, x0v, θ0v, T, cv, sd, adapt, factor, flags, run)
        end
        if st != 0
            # ZZB_E_BOUND still hands back a run (partial trace, error details): free it too before throwing
            run[] != C_NULL && ccall((:zzb_run_free, libzzb200), Int32, (Ptr{Cvoid},), run[])
            ccall((:zzb_problem_free, libzzb200), Int32, (Ptr{Cvoid},), prob[])
            check(st)
        end
    end
    try
        n = Ref{Int64}(0)
        check(ccall((:zzb_trace_len, libzzb200), Int32, (Ptr{Cvoid}, Ref{Int64}), run[], n))
        # the library writes 32-byte (Float64, Int64, Float64, Float64) records: build the trace with exactly that element type,
        # whatever the types of t0 (e.g. an Int) and of x0 / θ0 (e.g. Float32) are
        Ξ = ZigZagBoomerang.FactTrace(F, Float64(t0), x0v, θ0v, Vector{Tuple{Float64,Int,Float64,Float64}}(undef, n[]))
        check(ccall((:zzb_trace_copy, libzzb200), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64), run[], Ξ.events, 0, n[]))
        t, x, θ = similar(x0v), similar(x0v), similar(x0v)
        check(ccall((:zzb_run_final_state, libzzb200), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    run[], t, x, θ, cv))
        acc = Vector{Int}(undef, d); num = Ref{Int64}(0)
        check(ccall((:zzb_run_counts, libzzb200), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ref{Int64}), run[], acc, num))
        return Ξ, (t, x, θ), (acc, num[]), cv
    finally
        ccall((:zzb_run_free, libzzb200), Int32, (Ptr{Cvoid},), run[])
        ccall((:zzb_problem_free, libzzb200), Int32, (Ptr{Cvoid},), prob[])
    end
end

const Nbhd = Union{ZigZagBoomerang.All,ZigZagBoomerang.Matched}

# spdmp / pdmp, F::ZigZag (src/sfact.jl:162-214,236).  Matched() and All() select the same kernel (INTEGRATION.md).
function spdmp(∇ϕ::GaussianPotential, t0, x0, θ0, T, c, G::Nbhd, F::ZigZag, args...;
               factor = 1.8, adapt = false, adaptscale = false, progress = false, progress_stops = 20, seed = Seed())
    adaptscale && error("adaptscale = true is not implemented on the device path")
    Ξ, u, an, cv = device_run(F.λref == 0 ? :zigzag : :refresh, ∇ϕ, t0, x0, θ0, T, c, F; factor = factor, adapt = adapt, seed = seed)
    c .= cv                                            # adapted bounds, like the in-place `adapt!` of the reference
    Ξ, u, an, c
end

# spdmp / pdmp, F::FactBoomerang (same generic function upstream; flow src/sfact.jl:29-48, rate/bound src/fact_samplers.jl:37-39,58-65)
function spdmp(∇ϕ::GaussianPotential, t0, x0, θ0, T, c, G::Nbhd, F::FactBoomerang, args...;
               factor = 1.8, adapt = false, adaptscale = false, progress = false, progress_stops = 20, seed = Seed())
    adaptscale && error("adaptscale = true is not implemented on the device path")
    Ξ, u, an, cv = device_run(:boomerang, ∇ϕ, t0, x0, θ0, T, c, F; factor = factor, adapt = adapt, seed = seed)
    c .= cv
    Ξ, u, an, c
end

# spdmp(∇ϕ, t0, x0, θ0, T, C::LocalBound, F, ...) (src/local.jl:95-149)
function spdmp(∇ϕ::GaussianPotential, t0, x0, θ0, T, C::LocalBound, G::Nbhd, F::ZigZag, args...;
               factor = 1.8, adapt = false, progress = false, progress_stops = 20, seed = Seed())
    Ξ, u, an, cv = device_run(:zigzag, ∇ϕ, t0, x0, θ0, T, C.c, F; flags = ZZB_FLAG_LOCAL_BOUND, factor = factor, adapt = adapt, seed = seed)
    C.c .= cv
    Ξ, u, an, C
end

# Without G: the generated spdmp / pdmp(∇ϕ::GaussianPotential, ..., c, F::ZigZag, ...) methods below are more specific than the
# reference's spdmp / pdmp(∇ϕ, ..., C::LocalBound, F::Union{ZigZag,FactBoomerang,JointFlow}, ...) (src/local.jl:148-149) in
# arguments 1 and 7 but less specific in argument 6: state the intersection explicitly so that the documented drop-in call
# spdmp(GaussianPotential(Γ), t0, x0, θ0, T, LocalBound(c), Z) is not ambiguous.
spdmp(∇ϕ::GaussianPotential, t0, x0, θ0, T, C::LocalBound, F::ZigZag, args...; kargs...) =
    spdmp(∇ϕ, t0, x0, θ0, T, C, ZigZagBoomerang.Matched(), F, args...; kargs...)
pdmp(∇ϕ::GaussianPotential, t0, x0, θ0, T, C::LocalBound, F::ZigZag, args...; kargs...) =
    spdmp(∇ϕ, t0, x0, θ0, T, C, ZigZagBoomerang.All(), F, args...; kargs...)
spdmp(∇ϕ::GaussianPotential, t0, x0, θ0, T, C::LocalBound, F::FactBoomerang, args...; kargs...) =
    error("LocalBound with FactBoomerang is not implemented on the device path")
pdmp(∇ϕ::GaussianPotential, t0, x0, θ0, T, C::LocalBound, F::FactBoomerang, args...; kargs...) =
    error("LocalBound with FactBoomerang is not implemented on the device path")

# sspdmp(∇ϕ, t0, x0, θ0, T, c, [G,] F::ZigZag, κ, ...) (src/ss_fact.jl:159-217); acc is the scalar count of reflections
function sspdmp(∇ϕ::GaussianPotential, t0, x0, θ0, T, c, G, F::ZigZag, κ, args...;
                strong_upperbounds = false, factor = 1.5, adapt = false, reversible = false, seed = Seed())
    flags = (reversible ? ZZB_FLAG_STICKY_REVERSIBLE : UInt32(0)) | (strong_upperbounds ? ZZB_FLAG_STICKY_STRONG_UB : UInt32(0))
    Ξ, u, (acc, num), cv = device_run(:sticky, ∇ϕ, t0, x0, θ0, T, c, F; κ = κ, seed = seed, flags = flags, adapt = adapt, factor = factor)
    Ξ, u, (sum(acc), num), cv    # (adapt: the reference resets acc and num at every adaptation, src/ss_fact.jl:134; these are the totals)
end

# sspdmp2(∇ϕ, t, x0, v0, T, c, nothing, Z, κ; strong_upperbounds, adapt, factor) (src/stickyzz.jl:322-338): the dense sticky ZigZag
# `stickyzz` = the loop of sspdmp with proposal times at rate 0.01 + (a + b t)^+ and coordinates that start at 0 starting frozen.
# Returns trace, (acc, num) -- the reference returns (trace, acc::AcceptanceDiagnostics).
function sspdmp2(∇ϕ::GaussianPotential, t, x0, v0, T, c, ::Nothing, Z::ZigZag, κ, args...; strong_upperbounds = false, progress = false,
                 adapt = false, factor = 1.5, seed = Seed())
    flags = ZZB_FLAG_STICKY_ZZ | (strong_upperbounds ? ZZB_FLAG_STICKY_STRONG_UB : UInt32(0))
    Ξ, u, (acc, num), cv = device_run(:sticky, ∇ϕ, t, x0, v0, T, c, Z; κ = κ, seed = seed, flags = flags, adapt = adapt, factor = factor)
    Ξ, (sum(acc), num)
end
sspdmp(∇ϕ::GaussianPotential, t0, x0, θ0, T, c, F::ZigZag, κ, args...; kargs...) = sspdmp(∇ϕ, t0, x0, θ0, T, c, nothing, F, κ, args...; kargs...)

# sspdmp3(∇ϕ, u0, T, c, nothing, Z, κ, ...; rule) (src/sparsestickyzz.jl:405-422): the strong-bound sparse sticky ZigZag.  u0 is the
# reference's sparse state (`sparsestickystate(x0)`); its dense image (x0, θ0) is what the library takes -- coordinates at 0
# start frozen.  Returns trace, (acc, num), (t, x, θ).
function sspdmp3(∇ϕ::GaussianPotential, u0, T, c, ::Nothing, Z::ZigZag, κ, args...; adapt = false, factor = 1.5, rule = :reversible,
                 progress = false, progress_stops = 20, clusterα = 1.0, seed = Seed())
    (adapt || clusterα != 1.0) && error("sspdmp3 options adapt / clusterα < 1 are not implemented on the device path")
    init()
    d = length(u0)
    x0v, θ0v = zeros(d), zeros(d)
    for (i, ui) in pairs(u0.u)            # SparseState: i => (t, x, θ) for the active coordinates (src/sparsestickyzz.jl:3-16)
        x0v[i], θ0v[i] = ui[2], ui[3]
    end
    Γ = ∇ϕ.Γ
    prob = Ref{Ptr{Cvoid}}(C_NULL); run = Ref{Ptr{Cvoid}}(C_NULL)
    h = ∇ϕ.h === nothing ? Ptr{Float64}(C_NULL) : pointer(∇ϕ.h)
    sd = UInt64[seed[1], seed[2]]
    GC.@preserve Γ x0v θ0v sd ∇ϕ begin
        check(ccall((:zzb_problem_create_gaussian, libzzb200), Int32,
                    (Ref{Ptr{Cvoid}}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
                    prob, d, Γ.colptr, Γ.rowval, Γ.nzval, h, C_NULL, C_NULL, C_NULL, C_NULL))
        st = ccall((:zzb_sspdmp3_run, libzzb200), Int32,
                   (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Float64, Int32, Ptr{UInt64}, UInt32, Ref{Ptr{Cvoid}}),
                   prob[], x0v, θ0v, T, c, κ, rule === :sticky ? 0 : 1, sd, UInt32(0), run)
        if st != 0
            run[] != C_NULL && ccall((:zzb_run_free, libzzb200), Int32, (Ptr{Cvoid},), run[])
            ccall((:zzb_problem_free, libzzb200), Int32, (Ptr{Cvoid},), prob[])
            check(st)
        end
    end
    try
        n = Ref{Int64}(0)
        check(ccall((:zzb_trace_len, libzzb200), Int32, (Ptr{Cvoid}, Ref{Int64}), run[], n))
        Ξ = ZigZagBoomerang.FactTrace(Z, 0.0, x0v, θ0v, Vector{Tuple{Float64,Int,Float64,Float64}}(undef, n[]))
        check(ccall((:zzb_trace_copy, libzzb200), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64), run[], Ξ.events, 0, n[]))
        t, x, θ, cv = zeros(d), zeros(d), zeros(d), zeros(d)
        check(ccall((:zzb_run_final_state, libzzb200), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), run[], t, x, θ, cv))
        acc = Vector{Int}(undef, d); num = Ref{Int64}(0)
        check(ccall((:zzb_run_counts, libzzb200), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ref{Int64}), run[], acc, num))
        return Ξ, (sum(acc), num[]), (t, x, θ)
    finally
        ccall((:zzb_run_free, libzzb200), Int32, (Ptr{Cvoid},), run[])
        ccall((:zzb_problem_free, libzzb200), Int32, (Ptr{Cvoid},), prob[])
    end
end

# sspdmp4(Coloring, ∇ϕ, t, x0, v0, T, c, nothing, Z, κ; adapt, factor) (src/asynchzz.jl:250-265): the strong-bound sticky ZigZag of
# `asynchzz` with a bound constant c[i] and a thaw rate κ[i] per coordinate; coordinates with x0 == 0 start frozen and continue with
# v0 once thawed.  The reference's thread schedule (and its `Coloring`) is replaced by the device's own; same law.  Returns
# trace, (acc, num) -- the reference returns (trace, acc::AcceptanceDiagnostics).
function sspdmp4(Coloring, ∇ϕ::GaussianPotential, t, x0, v0, T, c, ::Nothing, Z::ZigZag, κ, args...; progress = false, adapt = false,
                 factor = 1.5, seed = Seed())
    adapt && error("sspdmp4(...; adapt = true) is not implemented on the device path")
    init()
    d = length(x0)
    x0v, θ0v = Vector{Float64}(x0), Vector{Float64}(v0)
    cv = c isa Number ? fill(Float64(c), d) : Vector{Float64}(c)
    κv = κ isa Number ? fill(Float64(κ), d) : Vector{Float64}(κ)
    Γ = ∇ϕ.Γ
    prob = Ref{Ptr{Cvoid}}(C_NULL); run = Ref{Ptr{Cvoid}}(C_NULL)
    h = ∇ϕ.h === nothing ? Ptr{Float64}(C_NULL) : pointer(∇ϕ.h)
    sd = UInt64[seed[1], seed[2]]
    GC.@preserve Γ x0v θ0v cv κv sd ∇ϕ begin
        check(ccall((:zzb_problem_create_gaussian, libzzb200), Int32,
                    (Ref{Ptr{Cvoid}}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
                    prob, d, Γ.colptr, Γ.rowval, Γ.nzval, h, C_NULL, C_NULL, C_NULL, C_NULL))
        st = ccall((:zzb_sspdmp4_run, libzzb200), Int32,
                   (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{UInt64}, UInt32, Ref{Ptr{Cvoid}}),
                   prob[], Float64(t), x0v, θ0v, T, cv, κv, sd, UInt32(0), run)
        if st != 0
            run[] != C_NULL && ccall((:zzb_run_free, libzzb200), Int32, (Ptr{Cvoid},), run[])
            ccall((:zzb_problem_free, libzzb200), Int32, (Ptr{Cvoid},), prob[])
            check(st)
        end
    end
    try
        n = Ref{Int64}(0)
        check(ccall((:zzb_trace_len, libzzb200), Int32, (Ptr{Cvoid}, Ref{Int64}), run[], n))
        θstart = [x0v[i] == 0 ? 0.0 : θ0v[i] for i in 1:d]
        Ξ = ZigZagBoomerang.FactTrace(Z, Float64(t), x0v, θstart, Vector{Tuple{Float64,Int,Float64,Float64}}(undef, n[]))
        check(ccall((:zzb_trace_copy, libzzb200), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64), run[], Ξ.events, 0, n[]))
        acc = Vector{Int}(undef, d); num = Ref{Int64}(0)
        check(ccall((:zzb_run_counts, libzzb200), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ref{Int64}), run[], acc, num))
        return Ξ, (sum(acc), num[])
    finally
        ccall((:zzb_run_free, libzzb200), Int32, (Ptr{Cvoid},), run[])
        ccall((:zzb_problem_free, libzzb200), Int32, (Ptr{Cvoid},), prob[])
    end
end

for FT in (:ZigZag, :FactBoomerang)
    @eval begin
        spdmp(∇ϕ::GaussianPotential, t0, x0, θ0, T, c, F::$FT, args...; kargs...) =
            spdmp(∇ϕ, t0, x0, θ0, T, c, ZigZagBoomerang.Matched(), F, args...; kargs...)
        pdmp(∇ϕ::GaussianPotential, t0, x0, θ0, T, c, F::$FT, args...; kargs...) =
            spdmp(∇ϕ, t0, x0, θ0, T, c, ZigZagBoomerang.All(), F, args...; kargs...)
    end
end

export GaussianPotential, LogisticSubsampled
end # module
