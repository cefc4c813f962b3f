# make_reference_golden.jl -- run ONCE on any machine with Julia and ZigZagBoomerang.jl 0.13.2 (commit 691afe2):
#
#     julia --project -e 'using Pkg; Pkg.add(name="ZigZagBoomerang", version="0.13.2")'
#     julia --project tests/golden/make_reference_golden.jl > tests/golden/reference_golden.txt
#
# It dumps, with every Float64 as its bit pattern, (i) the head of the reference's uniform stream
# `rand(Rng(seed))` (src/ZigZagBoomerang.jl:9-10: Rng = RandomNumbers.Xorshifts.Xoroshiro128Plus), (ii) `poisson_time` and
# `log` on fixed arguments and (iii) the complete `spdmp` traces of three small cases whose inputs are closed-form (no input
# RNG), with a fixed `seed=`.  tests/test_oracle.py::test_reference_golden_vectors compares the oracle in its faithful mode
# (single xoroshiro128+ stream consumed in event order, in-place moves: `seq | inplace`) with this file when it is present:
# that is what turns "parity unpinned" into a pinned oracle.  Neither this container nor the GPU box has Julia, so the file
# is not part of the repository yet.
using ZigZagBoomerang, SparseArrays, LinearAlgebra
const ZZB = ZigZagBoomerang
bits(x::Float64) = string(reinterpret(UInt64, x), base = 16, pad = 16)

# scripts/gridlaplace.jl:4-21
function gridlaplacian(T, m, n)
    S = sparse(T(0.0) * I, n * m, n * m)
    linear = LinearIndices((1:m, 1:n))
    for i in 1:m, j in 1:n
        for (i2, j2) in ((i + 1, j), (i, j + 1))
            if i2 <= m && j2 <= n
                S[linear[i, j], linear[i2, j2]] -= 1.0
                S[linear[i2, j2], linear[i, j]] -= 1.0
                S[linear[i, j], linear[i, j]] += 1.0
                S[linear[i2, j2], linear[i2, j2]] += 1.0
            end
        end
    end
    S
end

seed = (UInt64(0x0123456789abcdef), UInt64(0xfedcba9876543210))
println("# ZigZagBoomerang reference golden vectors; Float64 values are UInt64 bit patterns in hex")
rng = ZZB.Rng(seed)
println("rand ", join((bits(rand(rng)) for _ in 1:16), " "))
println("poisson_time ", join((bits(ZZB.poisson_time(a, b, u)) for (a, b, u) in
        ((1.5, 0.7, 0.3), (-0.4, 0.9, 0.8), (2.0, 0.0, 0.5), (0.8, -0.6, 0.9), (0.8, -0.6, 0.1), (-1.0, -1.0, 0.5))), " "))
println("log ", join((bits(log(u)) for u in (0.1, 0.25, 0.5, 0.75, 0.9999999, 1.0e-10)), " "))

function dump_case(name, Γt, Γb, μ, x0, θ0, T, c; kw...)
    ∇ϕ(x, i, Γ) = ZZB.idot(Γ, i, x)            # scripts/gaussianrandomfield.jl:25
    Z = ZigZag(Γb, μ)
    Ξ, (t, x, θ), (acc, num), cout = spdmp(∇ϕ, 0.0, copy(x0), copy(θ0), T, copy(c), Z, Γt; seed = seed, kw...)
    println("case ", name, " d ", length(x0), " T ", bits(T), " num ", num, " events ", length(Ξ.events))
    println("acc ", join(acc, " "))
    println("c ", join(bits.(cout), " "))
    println("final_t ", join(bits.(t), " "))
    println("final_x ", join(bits.(x), " "))
    println("final_theta ", join(bits.(θ), " "))
    for (te, i, xe, θe) in Ξ.events
        println("e ", bits(te), " ", i, " ", bits(xe), " ", bits(θe))
    end
end

# case 1: the 4 x 4 lattice GMRF of scripts/gaussianrandomfield.jl with closed-form initial state, c = column norms
n = 4
Γ = 0.01I + gridlaplacian(Float64, n, n)
d = n * n
x0 = [sin(Float64(i)) for i in 1:d]
θ0 = [isodd(i) ? 1.0 : -1.0 for i in 1:d]
c = [norm(Γ[:, i], 2) for i in 1:d]
dump_case("grid4", Γ, Γ, zeros(d), x0, θ0, 5.0, c)

# case 2: sampler matrix 0.9 Γ with Z.μ != 0 and adaptation of c (test/maintest.jl:23 uses 0.9 Γ too)
μ = [0.1 * cos(Float64(i)) for i in 1:d]
dump_case("grid4_scaled_mu_adapt", Γ, 0.9 * Γ, μ, x0, θ0, 5.0, 0.2 .* c; adapt = true)

# case 3: a 3 x 5 lattice with the tight bound c = sqrt(eps) of scripts/example.jl:39 (acceptance ~ 1)
Γ2 = 0.01I + gridlaplacian(Float64, 3, 5)
d2 = 15
dump_case("grid3x5_tight", Γ2, Γ2, zeros(d2), [cos(0.7 * i) for i in 1:d2], [i % 3 == 0 ? -1.0 : 1.0 for i in 1:d2], 8.0,
          fill(sqrt(eps()), d2))
