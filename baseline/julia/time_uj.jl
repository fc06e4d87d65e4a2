#=
time_uj.jl -- time the TRUE reference (FLOWVPM.jl + FastMultipole.jl) on the bench workload.

Julia is not available in this repository's build/bench environment, so bench.py's
cpu_baseline is the op-for-op C restatement in oracle/.  Anyone with Julia >= 1.6 can time
the real thing beside it:

    julia --threads=auto --project=/path/to/FLOWVPM.jl baseline/julia/time_uj.jl 65536

It builds the same jittered-lattice cloud recipe as bench.py (scripts/benchmark_fmm2.jl:11-34
of the reference; the random numbers differ, the work per interaction does not) and prints
interactions/s of UJ_direct (winckelmans, U+J) and of UJ_direct with the SFS sweep.
=#
import FLOWVPM
const vpm = FLOWVPM
using Random

function cloud(n; seed=20240607)
    Random.seed!(seed)
    d = (7.0 / n)^(1 / 3)
    nx = ny = max(1, ceil(Int, 1 / d))
    pfield = vpm.ParticleField(n; formulation=vpm.rVPM, kernel=vpm.winckelmans, UJ=vpm.UJ_direct)
    for i in 0:n-1
        iz, rem = divrem(i, nx * ny)
        iy, ix = divrem(rem, nx)
        X = ((ix + 0.5) * d, (iy + 0.5) * d, (iz + 0.5) * d) .+ (rand(3) .- 0.5) .* (0.5 * d)
        Gamma = [0, 0, 1.0 / n] .+ (rand(3) .- 0.5) ./ 10
        sigma = d / 2 * 1.3 * (1 + (rand() - 0.5) / 10)
        vpm.add_particle(pfield, X, Gamma, sigma)
    end
    return pfield
end

n = length(ARGS) > 0 ? parse(Int, ARGS[1]) : 32768
pfield = cloud(n)
pfield.UJ(pfield)                        # compile
t = @elapsed pfield.UJ(pfield; reset=true)
println("UJ_direct  U+J : $(n)^2 interactions in $(t) s = $(n^2 / t) interactions/s on $(Threads.nthreads()) threads")
t = @elapsed pfield.UJ(pfield; reset=true, reset_sfs=true, sfs=true)
println("UJ_direct + SFS: $(t) s")
