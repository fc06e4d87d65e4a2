#=
make_golden.jl -- run the TRUE reference (FLOWVPM.jl v4.0.x + FastMultipole.jl) on the committed
golden inputs and dump what it computes, so that parity can be pinned to the reference itself.

    julia --project=/path/to/FLOWVPM.jl baseline/julia/make_golden.jl [repo_root]

Reads  tests/golden/ref_inputs.f64 / .txt   (written by tests/golden/make_ref_inputs.py)
Writes tests/golden/ref_outputs.f64 / .txt  (read by tests/test_reference_golden.py)

For every input case (8 x N: X, Gamma, sigma, static) and every kernel family it builds a
ParticleField with `add_particle` (src/FLOWVPM_particlefield.jl:167-194), calls the reference's
own `UJ_direct(pfield; sfs=true, reset=true, reset_sfs=true)` (src/FLOWVPM_UJ.jl:21-37) with the
transposed and the classic SFS scheme, and stores rows 10:12 (U), 16:24 (J) and 40:42 (SFS) as a
15 x N array named `<case>/<kernel>/<T|C>`.  A second call without reset is stored as
`<case>/<kernel>/accumulate` (rows after UJ_direct(pfield; reset=false, sfs=false) on top of the
first result: the reset / static-particle rules of src/FLOWVPM_particlefield.jl:464-511).
=#
import FLOWVPM
const vpm = FLOWVPM

root = length(ARGS) > 0 ? ARGS[1] : normpath(joinpath(@__DIR__, "..", ".."))
gold = joinpath(root, "tests", "golden")

function read_arrays(path)
    raw = reinterpret(Float64, read(path * ".f64"))
    out = Dict{String,Matrix{Float64}}()
    for line in eachline(path * ".txt")
        name, r, c, off = split(line)
        r, c, off = parse(Int, r), parse(Int, c), parse(Int, off)
        out[name] = reshape(collect(raw[off+1:off+r*c]), r, c)
    end
    return out
end

kernels = Dict("singular" => vpm.kernel_singular, "gaussian" => vpm.kernel_gaussian,
               "gaussianerf" => vpm.kernel_gaussianerf, "winckelmans" => vpm.kernel_winckelmans)
ROWS = vcat(10:12, 16:24, 40:42)

function field(a, kernel, transposed)
    n = size(a, 2)
    pfield = vpm.ParticleField(n; kernel=kernel, UJ=vpm.UJ_direct, transposed=transposed)
    for i in 1:n
        vpm.add_particle(pfield, a[1:3, i], a[4:6, i], a[7, i]; static=(a[8, i] != 0))
    end
    return pfield
end

inputs = read_arrays(joinpath(gold, "ref_inputs"))
results = Pair{String,Matrix{Float64}}[]
for (case, a) in sort(collect(inputs); by=first), (kname, kernel) in sort(collect(kernels); by=first)
    for (tag, transposed) in (("T", true), ("C", false))
        pfield = field(a, kernel, transposed)
        vpm.UJ_direct(pfield; sfs=true, reset=true, reset_sfs=true)
        push!(results, "$case/$kname/$tag" => pfield.particles[ROWS, 1:pfield.np])
        if transposed
            vpm.UJ_direct(pfield; sfs=false, reset=false, reset_sfs=false)
            push!(results, "$case/$kname/accumulate" => pfield.particles[ROWS, 1:pfield.np])
        end
    end
end

open(joinpath(gold, "ref_outputs.f64"), "w") do fb
    open(joinpath(gold, "ref_outputs.txt"), "w") do ft
        off = 0
        for (name, m) in results
            write(fb, m)                                   # column-major Float64, little endian on x86/ARM
            println(ft, "$name $(size(m, 1)) $(size(m, 2)) $off")
            off += length(m)
        end
    end
end
println("wrote $(length(results)) arrays to ", joinpath(gold, "ref_outputs.f64"),
        "  (FLOWVPM ", pkgversion(vpm), ", ", Threads.nthreads(), " threads)")
