#=##############################################################################
# FLOWVPMCuda.jl -- Julia binding of libvpm_cuda.so for FLOWVPM.jl v4.0.x
#
# The reference keeps its whole host side (ParticleField, formulations, RK3,
# relaxation, dynamic SFS, I/O).  This file is the only thing a maintainer adds:
# a callable for the `UJ` slot of the ParticleField plus, optionally, overloads of
# the two FastMultipole hooks, all of which `ccall` the C ABI of
# include/vpm_cuda.h.  It cannot be executed in the build environment of this
# repository (no Julia there); tests/ exercise the identical ABI through ctypes.
#
# Usage
#     import FLOWVPM; const vpm = FLOWVPM
#     include("FLOWVPMCuda.jl"); using .FLOWVPMCuda
#     FLOWVPMCuda.init!("/path/to/libvpm_cuda.so"; n_gpus=1)
#     pfield = vpm.ParticleField(maxp; UJ=FLOWVPMCuda.UJ_cuda, kernel=vpm.winckelmans, ...)
#     FLOWVPMCuda.pin!(pfield)                 # optional: page-lock pfield.particles once
#     vpm.run_vpm!(pfield, dt, nsteps)         # every pfield.UJ(...) call now runs on the GPU(s)
#
# Install it in the slot, not only as `custom_UJ`: the dynamic-SFS test filter and the
# relaxation step call `pfield.UJ` directly (src/FLOWVPM_subfilterscale.jl:478,
# src/FLOWVPM_timeintegration.jl:244,443).
=###############################################################################
module FLOWVPMCuda

import FLOWVPM
const vpm = FLOWVPM
const fmm = FLOWVPM.fmm

const VPM_OK = Cint(0)
const FLAG_RESET = Cint(1)
const FLAG_RESET_SFS = Cint(2)
const FLAG_SFS = Cint(4)
const FLAG_TRANSPOSED = Cint(8)
const FLAG_NO_FARFIELD_SHORTCUT = Cint(16)
const FLAG_FP32 = Cint(32)   # library-only: FP32-arithmetic U/J sweep (1e-5 bar); default is FP64

const lib = Ref{String}("libvpm_cuda.so")
const handle = Ref{Ptr{Cvoid}}(C_NULL)

"Kernel family id of include/vpm_cuda.h from the reference's singletons (src/FLOWVPM.jl:129-133)"
function kernel_id(k::vpm.Kernel)
    k === vpm.kernel_singular    && return Cint(0)
    k === vpm.kernel_gaussian    && return Cint(1)
    k === vpm.kernel_gaussianerf && return Cint(2)
    k === vpm.kernel_winckelmans && return Cint(3)
    error("FLOWVPMCuda: kernel $(k) has no CUDA implementation (user-defined kernels stay on UJ_direct/UJ_fmm)")
end

function check(rc::Cint)
    rc == VPM_OK && return nothing
    msg = unsafe_string(ccall((:vpm_last_error, lib[]), Cstring, (Ptr{Cvoid},), handle[]))
    error("libvpm_cuda error $(rc): $(msg)")   # the reference reports failures with error(...)
end

"Open the library and create the handle (device buffers, streams, NCCL communicators)."
function init!(path::AbstractString="libvpm_cuda.so"; n_gpus::Integer=1, devices=nothing)
    lib[] = String(path)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    ids = devices === nothing ? C_NULL : Cint.(collect(devices))
    n = devices === nothing ? Cint(n_gpus) : Cint(length(devices))
    rc = ccall((:vpm_create, lib[]), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cint}), h, n, ids)
    if rc != VPM_OK
        msg = unsafe_string(ccall((:vpm_last_error, lib[]), Cstring, (Ptr{Cvoid},), C_NULL))
        error("libvpm_cuda: vpm_create failed ($(rc)): $(msg)")
    end
    handle[] = h[]
    atexit(shutdown!)
    return nothing
end

function shutdown!()
    if handle[] != C_NULL
        ccall((:vpm_destroy, lib[]), Cint, (Ptr{Cvoid},), handle[])
        handle[] = C_NULL
    end
end

"Page-lock pfield.particles (allocated once at maxparticles, src/FLOWVPM_particlefield.jl:134)."
function pin!(pfield::vpm.ParticleField)
    P = pfield.particles
    GC.@preserve P check(ccall((:vpm_pin_host, lib[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t),
                               handle[], pointer(P), sizeof(P)))
end

function flags(pfield; sfs, reset, reset_sfs)
    f = Cint(0)
    reset && (f |= FLAG_RESET)
    reset_sfs && (f |= FLAG_RESET_SFS)
    sfs && (f |= FLAG_SFS)
    pfield.transposed && (f |= FLAG_TRANSPOSED)
    return f
end

"""
    UJ_cuda(pfield; rbf=false, sfs=false, reset=true, reset_sfs=false, optargs...)

Drop-in for `UJ_direct(pfield; ...)` (src/FLOWVPM_UJ.jl:21-37): same keywords, same
effects on rows 10:27 and 40:42 of `pfield.particles` (reset-then-accumulate, static
particles never reset, SFS sweep after the final J).  `rbf` is accepted and ignored,
as in the reference.
"""
function UJ_cuda(pfield::vpm.ParticleField{Float64}; rbf::Bool=false, sfs::Bool=false,
                 reset::Bool=true, reset_sfs::Bool=false, fp32::Bool=false, optargs...)
    P = pfield.particles
    # fp32=true (library-only keyword): U/J sweep in FP32 arithmetic, 1e-5 instead of 1e-12
    GC.@preserve P check(ccall((:vpm_uj_direct, lib[]), Cint,
                               (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Cint, Cint),
                               handle[], P, size(P, 1), pfield.np, kernel_id(pfield.kernel),
                               flags(pfield; sfs, reset, reset_sfs) | (fp32 ? FLAG_FP32 : Cint(0))))
    return nothing
end

function UJ_cuda(pfield::vpm.ParticleField{Float32}; rbf::Bool=false, sfs::Bool=false,
                 reset::Bool=true, reset_sfs::Bool=false, optargs...)
    P = pfield.particles
    GC.@preserve P check(ccall((:vpm_uj_direct_f32, lib[]), Cint,
                               (Ptr{Cvoid}, Ptr{Float32}, Int64, Int64, Cint, Cint),
                               handle[], P, size(P, 1), pfield.np, kernel_id(pfield.kernel),
                               flags(pfield; sfs, reset, reset_sfs)))
    return nothing
end

"`UJ_direct(source, target)` (src/FLOWVPM_UJ.jl:48-50): probes / fluid-domain evaluation."
function UJ_cuda(source::vpm.ParticleField{Float64}, target::vpm.ParticleField{Float64})
    S, T = source.particles, target.particles
    GC.@preserve S T check(ccall((:vpm_uj_direct_st, lib[]), Cint,
                                 (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Int64, Int64, Cint),
                                 handle[], S, size(S, 1), source.np, T, size(T, 1), target.np,
                                 kernel_id(source.kernel)))
    return nothing
end

# ------------------------------------------------------------------------------
# Hook 2 -- the FastMultipole pair-loop overload (src/FLOWVPM_fmm.jl:102-168).
# Replaces the body of FLOWVPM's own `fmm.direct!` method: FastMultipole calls it for
# `direct!(system)` and for every near-field leaf pair on the CPU path.  Row offsets of
# the target buffer are FastMultipole's (position 1:3, scalar potential 4, gradient
# 5:7, hessian 8:16) and therefore arguments of the ABI.
# ------------------------------------------------------------------------------
function direct_cuda!(target_buffer::Matrix{Float64}, target_index::UnitRange,
                      ::fmm.DerivativesSwitch{PS,VS,GS}, source_system::vpm.ParticleField,
                      source_buffer::Matrix{Float64}, source_index::UnitRange) where {PS,VS,GS}
    GC.@preserve target_buffer source_buffer check(ccall((:vpm_p2p_buffers, lib[]), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Cint, Cint, Cint, Ptr{Float64}, Int64, Int64, Cint, Cint, Cint),
        handle[], target_buffer, size(target_buffer, 1), first(target_index) - 1, last(target_index),
        0, 4, 7, source_buffer, first(source_index) - 1, last(source_index),
        kernel_id(source_system.kernel), VS, GS))
    return nothing
end

# ------------------------------------------------------------------------------
# Hook 3 -- the FMM near-field device hook.  With `useGPU>0`, UJ_fmm passes
# `nearfield_device=true` (src/FLOWVPM_UJ.jl:97) and FastMultipole calls
# `fmm.nearfield_device!`.  The method below has the only call shape the reference shows
# (src/FLOWVPM_gpu.jl:637-643): both systems are ParticleFields, `target_indices[k]` is the body
# range of a target leaf and `source_indices[k]` the source range -- or the vector of source
# ranges gathered for that leaf (combine_source_indices, :554-580).  The whole list goes to the
# GPU(s) in ONE call (vpm_nearfield_ranges) instead of one launch per target leaf.
# ------------------------------------------------------------------------------
function fmm.nearfield_device!(target_system::vpm.ParticleField{Float64}, target_indices::AbstractVector{<:UnitRange},
                               ::fmm.DerivativesSwitch{PS,VS,GS}, source_system::vpm.ParticleField{Float64},
                               source_indices::AbstractVector) where {PS,VS,GS}
    ntr = length(target_indices)
    length(source_indices) == ntr || error("nearfield_device!: one source entry per target leaf expected")
    tb = Int64[first(r) - 1 for r in target_indices]     # 0-based, half-open on the C side
    te = Int64[last(r) for r in target_indices]
    soff = zeros(Int64, ntr + 1)
    sb, se = Int64[], Int64[]
    for k in 1:ntr
        g = source_indices[k]
        for r in (g isa UnitRange ? (g,) : g)
            push!(sb, first(r) - 1); push!(se, last(r))
        end
        soff[k + 1] = length(sb)
    end
    TP, SP = target_system.particles, source_system.particles
    GC.@preserve TP SP tb te sb se soff check(ccall((:vpm_nearfield_ranges, lib[]), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int64,
         Ptr{Float64}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Cint, Cint, Cint),
        handle[], TP, size(TP, 1), target_system.np, tb, te, ntr,
        SP, size(SP, 1), source_system.np, sb, se, soff, kernel_id(source_system.kernel), VS, GS))
    return nothing
end

"""
    UJ_fmm_cuda(pfield; verbose, rbf, sfs, reset, reset_sfs, autotune)

`UJ_fmm` (src/FLOWVPM_UJ.jl:62-129) with the near field and the SFS term on the GPU(s) and
NO edit of the reference: the reset rules (:75-80), the `fmm.fmm!` call with
`nearfield_device=true` (:90-101; FastMultipole then calls `fmm.nearfield_device!` above), the
autotune bookkeeping (:104-118) are restated; `Estr_fmm!` (:125) becomes `Estr_cuda!`.
Install it in the slot: `ParticleField(maxp; UJ=FLOWVPMCuda.UJ_fmm_cuda, ...)`.
"""
function UJ_fmm_cuda(pfield::vpm.ParticleField{Float64}; verbose::Bool=false, rbf::Bool=false, sfs::Bool=false,
                     sfs_type::Int=-1, transposed_sfs::Bool=true, reset::Bool=true, reset_sfs::Bool=false,
                     autotune::Bool=true, optargs...)
    reset && vpm._reset_particles(pfield)
    (reset_sfs || sfs) && vpm._reset_particles_sfs(pfield)
    o = pfield.fmm
    if rbf
        vpm.zeta_fmm(pfield)
    else
        args = fmm.fmm!(pfield; expansion_order=o.p - 1, leaf_size_source=max(o.ncrit, o.min_ncrit),
                        multipole_acceptance=o.theta,
                        error_tolerance=fmm.PowerRelativeGradient{o.relative_tolerance, o.absolute_tolerance, true}(),
                        tune=true, shrink_recenter=o.shrink_recenter, nearfield_device=true,
                        scalar_potential=false, hessian=true, silence_warnings=!verbose)
        tuned, cache, target_tree, source_tree, m2l_list, direct_list, _ = args
        if autotune
            new_p = o.autotune_p ? tuned.expansion_order + 1 : o.p
            new_ncrit = o.autotune_ncrit ? tuned.leaf_size_source[1] : o.ncrit
            pfield.fmm = vpm.FMM(new_p, new_ncrit, o.theta, o.shrink_recenter, o.relative_tolerance,
                                 o.absolute_tolerance, o.autotune_p, o.autotune_ncrit, o.autotune_reg_error,
                                 o.default_rho_over_sigma, o.min_ncrit)
        end
        sfs && Estr_cuda!(pfield, target_tree, source_tree, direct_list)
    end
    return nothing
end

"The whole direct_list on FastMultipole's own (tree-sorted) buffers in one call (vpm_p2p_leafpairs)."
function nearfield_cuda!(target_buffer::Matrix{Float64}, target_branches, source_system::vpm.ParticleField,
                         source_buffer::Matrix{Float64}, source_branches, direct_list;
                         velocity::Bool=true, velocity_gradient::Bool=true)
    tb = Int64[first(b.bodies_index[1]) - 1 for b in target_branches]
    te = Int64[last(b.bodies_index[1]) for b in target_branches]
    sb = Int64[first(b.bodies_index[1]) - 1 for b in source_branches]
    se = Int64[last(b.bodies_index[1]) for b in source_branches]
    pt = Int32[p[1] - 1 for p in direct_list]
    ps = Int32[p[2] - 1 for p in direct_list]
    GC.@preserve target_buffer source_buffer tb te sb se pt ps check(ccall((:vpm_p2p_leafpairs, lib[]), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Cint, Cint, Cint, Ptr{Float64}, Int64,
         Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Int32}, Ptr{Int32}, Int64, Cint, Cint, Cint),
        handle[], target_buffer, size(target_buffer, 1), size(target_buffer, 2), 0, 4, 7,
        source_buffer, size(source_buffer, 2), tb, te, length(tb), sb, se, length(sb), pt, ps, length(pt),
        kernel_id(source_system.kernel), velocity, velocity_gradient))
    return nothing
end

"`Estr_fmm!` over the near-field list (src/FLOWVPM_subfilterscale_models.jl:94-188)."
function Estr_cuda!(pfield::vpm.ParticleField{Float64}, target_tree, source_tree, direct_list)
    P = pfield.particles
    ts = Int64.(target_tree.sort_index_list[1]) .- 1
    ss = Int64.(source_tree.sort_index_list[1]) .- 1
    tb = Int64[first(b.bodies_index[1]) - 1 for b in target_tree.branches]
    te = Int64[last(b.bodies_index[1]) for b in target_tree.branches]
    sb = Int64[first(b.bodies_index[1]) - 1 for b in source_tree.branches]
    se = Int64[last(b.bodies_index[1]) for b in source_tree.branches]
    pt = Int32[p[1] - 1 for p in direct_list]
    ps = Int32[p[2] - 1 for p in direct_list]
    GC.@preserve P ts ss tb te sb se pt ps check(ccall((:vpm_estr_leafpairs, lib[]), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64,
         Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Int32}, Ptr{Int32}, Int64, Cint, Cint),
        handle[], P, size(P, 1), pfield.np, ts, ss, tb, te, length(tb), sb, se, length(sb), pt, ps, length(pt),
        kernel_id(pfield.kernel), flags(pfield; sfs=true, reset=false, reset_sfs=false)))
    return nothing
end

"`zeta_direct(pfield)` (src/FLOWVPM_viscous.jl:488-515): J[1:3] <- sum_j Gamma_j zeta_sigma_j."
function zeta_cuda(pfield::vpm.ParticleField{Float64})
    P = pfield.particles
    GC.@preserve P check(ccall((:vpm_zeta_direct, lib[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Cint),
                               handle[], P, size(P, 1), pfield.np, kernel_id(pfield.kernel)))
    return nothing
end

# ------------------------------------------------------------------------------
# Optional: leaf lists built on the device (include/vpm_cuda.h vpm_leaflists_*, SURVEY 8 f-3).
# `leaflists_cuda!` replaces the host tree of the near field when FastMultipole's own tree is
# not needed (near-field-only evaluations, or as the source of `direct_list` for the hooks
# above); `UJ_nearfield_cuda!` is the near-field half of UJ_fmm (src/FLOWVPM_UJ.jl:62-129)
# over those resident lists.
# ------------------------------------------------------------------------------
function leaflists_cuda!(pfield::vpm.ParticleField{Float64}; ncrit::Integer=50, theta::Real=pfield.fmm.theta,
                         fetch::Bool=false)
    P = pfield.particles
    nl, npairs = Ref{Int64}(0), Ref{Int64}(0)
    GC.@preserve P check(ccall((:vpm_leaflists_build, lib[]), Cint,
                               (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Cdouble, Ref{Int64}, Ref{Int64}),
                               handle[], P, size(P, 1), pfield.np, ncrit, theta, nl, npairs))
    fetch || return (n_leaves=nl[], n_pairs=npairs[])
    sort_index = Vector{Int64}(undef, pfield.np)
    lb, le = Vector{Int64}(undef, nl[]), Vector{Int64}(undef, nl[])
    pt, ps = Vector{Int32}(undef, npairs[]), Vector{Int32}(undef, npairs[])
    GC.@preserve sort_index lb le pt ps check(ccall((:vpm_leaflists_get, lib[]), Cint,
        (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}), handle[], sort_index, lb, le, pt, ps))
    # 0-based, half-open on the C side -> 1-based Julia ranges
    return (sort_index=sort_index .+ 1, leaves=[(lb[i]+1):le[i] for i in 1:nl[]],
            direct_list=[(pt[k] + 1, ps[k] + 1) for k in 1:npairs[]])
end

"Run the U/J leaf-list kernels (Hook 3, `UJ_nearfield_cuda!`) in FP32 arithmetic (errors ~1e-6, below FMM truncation)."
function nearfield_fp32!(on::Bool=true)
    check(ccall((:vpm_set_option, lib[]), Cint, (Ptr{Cvoid}, Cint, Cint), handle[], Cint(1), Cint(on)))
end

"""
The lists built by `leaflists_cuda!` are valid for one particle configuration: after positions or core sizes
change (every RK substep) call `leaflists_cuda!` again; the library checks a fingerprint and errors otherwise.
"""
function UJ_nearfield_cuda!(pfield::vpm.ParticleField{Float64}; reset::Bool=true)
    P = pfield.particles
    GC.@preserve P check(ccall((:vpm_uj_nearfield, lib[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Cint, Cint),
                               handle[], P, size(P, 1), pfield.np, kernel_id(pfield.kernel), reset ? FLAG_RESET : Cint(0)))
    return nothing
end

# ------------------------------------------------------------------------------
# Optional: the integrator on the device (include/vpm_cuda.h vpm_field_*).  Mirrors
# `nextstep` (src/FLOWVPM_particlefield.jl:435-460) for ReformulatedVPM{f,g}, NoSFS /
# ConstantSFS / DynamicSFS (pseudo3level), Inviscid or CoreSpreading, relaxation pedrizzetti /
# correctedpedrizzetti.
# ------------------------------------------------------------------------------
struct StepParams
    dt::Cdouble; f::Cdouble; g::Cdouble; Uinf::NTuple{3,Cdouble}; Cs::Cdouble; rlxf::Cdouble
    alpha::Cdouble; sfs_rlxf::Cdouble; minC::Cdouble; maxC::Cdouble; deltat::Cdouble
    nu::Cdouble; sgm0::Cdouble; cs_beta::Cdouble; cs_tol::Cdouble
    kernel_id::Int32; integration::Int32; relaxation::Int32; relax::Int32
    sfs::Int32; clip_backscatter::Int32; transposed::Int32; force_positive::Int32
    controls::Int32; viscous::Int32; cs_itmax::Int32; cs_iterror::Int32
end

function upload!(pfield::vpm.ParticleField{Float64})
    P = pfield.particles
    GC.@preserve P check(ccall((:vpm_field_upload, lib[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64),
                               handle[], P, size(P, 1), pfield.np))
end

function download!(pfield::vpm.ParticleField{Float64})
    P = pfield.particles
    GC.@preserve P check(ccall((:vpm_field_download, lib[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64),
                               handle[], P, size(P, 1), pfield.np))
end

function nextstep_cuda!(pfield::vpm.ParticleField{Float64}, dt::Real; relax::Bool=false,
                        clip_backscatter::Bool=false, force_positive::Bool=false,
                        control_directional::Bool=false, control_magnitude::Bool=false)
    form = pfield.formulation
    # ClassicVPM{R} is an empty struct (src/FLOWVPM_formulation.jl:23) and its integrators pass
    # reset_sfs=true on every substep: only the reformulated family is mirrored on the device
    # (cVPM as ReformulatedVPM(0, 0), the reference's own `formulation_cVPM`)
    form isa vpm.ReformulatedVPM || error("nextstep_cuda!: use ReformulatedVPM{R}(f, g) (cVPM = ReformulatedVPM(0, 0)); ClassicVPM stays on the host integrators")
    S = pfield.SFS
    sfs = S isa vpm.DynamicSFS ? 2 : (vpm.isSFSenabled(S) ? 1 : 0)
    Cs = S isa vpm.ConstantSFS ? S.Cs : 1.0
    alpha, sfs_rlxf, minC, maxC = S isa vpm.DynamicSFS ? (S.alpha, S.rlxf, S.minC, S.maxC) : (0.667, 0.005, 0.0, 1.0)
    rlx = pfield.relaxation
    relaxation = rlx.relax === vpm.relax_pedrizzetti ? 1 : rlx.relax === vpm.relax_correctedpedrizzetti ? 2 : 0
    integration = pfield.integration === vpm.rungekutta3 ? 1 : 0
    Uinf = pfield.Uinf(pfield.t)
    deltat = pfield.nt > 0 ? pfield.t / pfield.nt : 0.0
    V = pfield.viscous
    cs = V isa vpm.CoreSpreading
    pse = V isa vpm.ParticleStrengthExchange   # per-particle part only (src/FLOWVPM_viscous.jl:257-298)
    nu, sgm0, cs_beta, cs_tol = cs ? (V.nu, V.sgm0, V.beta, V.tol) : (pse ? V.nu : 0.0, 1.0, 1.5, 1e-3)
    if cs
        # the scheme's `zeta` argument (src/FLOWVPM_viscous.jl:63-141): zeta_fmm -> near field of device-built
        # leaf lists with pfield.fmm's leaf size and acceptance, accumulating on J[1:3] exactly as the
        # reference's zeta_fmm does; anything else -> zeta_direct
        zfmm = V.zeta === vpm.zeta_fmm
        check(ccall((:vpm_field_zeta_method, lib[]), Cint, (Ptr{Cvoid}, Cint, Int64, Cdouble), handle[],
                    zfmm ? 1 : 0, pfield.fmm.ncrit, pfield.fmm.theta))
        t_sgm = Ref{Cdouble}(V.t_sgm)
        check(ccall((:vpm_field_tsgm, lib[]), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Cint), handle[], t_sgm, 1))
    end
    sp = Ref(StepParams(dt, form.f, form.g, (Uinf[1], Uinf[2], Uinf[3]), Cs, rlx.rlxf, alpha, sfs_rlxf, minC, maxC,
                        deltat, nu, sgm0, cs_beta, cs_tol, kernel_id(pfield.kernel), integration, relaxation, relax,
                        sfs, clip_backscatter, pfield.transposed, force_positive,
                        Int32(control_directional) | (Int32(control_magnitude) << 1),
                        cs ? 1 : (pse ? (V.recalculate_vols ? 2 : 3) : 0), cs ? V.itmax : 15, cs ? V.iterror : true))
    check(ccall((:vpm_field_step, lib[]), Cint, (Ptr{Cvoid}, Ref{StepParams}), handle[], sp))
    if cs
        check(ccall((:vpm_field_tsgm, lib[]), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Cint), handle[], t_sgm, 0))
        V.t_sgm = t_sgm[]
    end
    pfield.t += dt
    pfield.nt += 1
    return nothing
end

export UJ_cuda, UJ_fmm_cuda, zeta_cuda

end # module
