"""The UJ slot of the reference, backed by libvpm_cuda.so.

Mirrors src/FLOWVPM_UJ.jl: `UJ_direct(pfield; rbf, sfs, reset, reset_sfs)`
(:21-37), `UJ_direct(source, target)` (:48-50) and the near-field half of
`UJ_fmm` (:62-129): FastMultipole's 6-argument `fmm.direct!` overload on
buffers (src/FLOWVPM_fmm.jl:102-168), the `nearfield_device!` hook over a
direct_list, and `Estr_fmm!` (src/FLOWVPM_subfilterscale_models.jl:94-188).
Every function calls the C ABI; nothing here computes pair arithmetic.
"""
import ctypes as C

import numpy as np

from . import _cabi
from .particlefield import ParticleField, NFIELDS

_default_handle = None


def get_handle():
    """Process-wide single-GPU handle (created on first use; raises without a GPU)."""
    global _default_handle
    if _default_handle is None:
        _default_handle = _cabi.Handle(1)
    return _default_handle


def set_handle(handle):
    global _default_handle
    _default_handle = handle


def _flags(pfield, sfs, reset, reset_sfs, no_shortcut=False, fp32=False):
    f = 0
    if reset:
        f |= _cabi.FLAG_RESET
    if reset_sfs:
        f |= _cabi.FLAG_RESET_SFS
    if sfs:
        f |= _cabi.FLAG_SFS
    if pfield.transposed:
        f |= _cabi.FLAG_TRANSPOSED
    if no_shortcut:
        f |= _cabi.FLAG_NO_FARFIELD_SHORTCUT
    if fp32:
        f |= _cabi.FLAG_FP32
    return f


def _check_matrix(P):
    if not (isinstance(P, np.ndarray) and P.ndim == 2 and P.flags.f_contiguous):
        raise ValueError("particles must be a Fortran-ordered 2-D array (Julia Matrix layout)")


def UJ_direct(pfield, target=None, *, rbf=False, sfs=False, reset=True, reset_sfs=False,
              handle=None, no_farfield_shortcut=False, fp32=False, **optargs):
    """UJ_direct(pfield; rbf, sfs, reset=true, reset_sfs=false)  -- src/FLOWVPM_UJ.jl:21-37
    UJ_direct(source, target)                                    -- src/FLOWVPM_UJ.jl:48-50

    `rbf` is accepted and ignored, as in the reference.  `fp32=True` (library-only) runs
    the U/J sweep in FP32 arithmetic (VPM_FLAG_FP32, 1e-5 bar); the default is FP64."""
    h = handle or get_handle()
    if target is not None:
        src = pfield
        _check_matrix(src.particles)
        _check_matrix(target.particles)
        if src.particles.dtype != np.float64 or target.particles.dtype != np.float64:
            raise TypeError("UJ_direct(source, target) is FP64 only")
        h.check(h.lib.vpm_uj_direct_st(h.ptr, src.particles.ctypes.data, src.particles.shape[0],
                                       src.np, target.particles.ctypes.data,
                                       target.particles.shape[0], target.np, src.kernel.id))
        return None
    P = pfield.particles
    _check_matrix(P)
    flags = _flags(pfield, sfs, reset, reset_sfs, no_farfield_shortcut, fp32)
    if P.dtype == np.float64:
        fn = h.lib.vpm_uj_direct
    elif P.dtype == np.float32:
        fn = h.lib.vpm_uj_direct_f32
    else:
        raise TypeError(f"unsupported element type {P.dtype}")
    h.check(fn(h.ptr, P.ctypes.data, P.shape[0], pfield.np, pfield.kernel.id, flags))
    return None


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _pairs(direct_list):
    """(pair_tgt, pair_src) as contiguous int32 arrays from either an (n, 2) array of leaf index
    pairs (the reference's direct_list) or a 2-TUPLE of 1-D arrays (no copy when they already
    are contiguous int32: a 9e7-entry list costs 0.2 s to de-interleave on the host)."""
    if isinstance(direct_list, tuple) and len(direct_list) == 2 and np.ndim(direct_list[0]) == 1:
        pt, ps = _i32(direct_list[0]), _i32(direct_list[1])
        if len(pt) != len(ps):
            raise ValueError("pair_tgt and pair_src must have the same length")
        return pt, ps
    dl = np.asarray(direct_list)
    if dl.ndim != 2 or dl.shape[1] != 2:
        raise ValueError("direct_list must be (n_pairs, 2) or a (pair_tgt, pair_src) tuple")
    return _i32(dl[:, 0]), _i32(dl[:, 1])


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# FastMultipole's target-buffer convention as used through its accessors:
# position rows 1:3, scalar potential 4, gradient 5:7, hessian 8:16 (0-based below)
ROW_POS, ROW_GRAD, ROW_HESS = 0, 4, 7


def direct_buffers(target_buffer, target_index, source_buffer, source_index, kernel, *,
                   want_U=True, want_J=True, row_pos=ROW_POS, row_grad=ROW_GRAD,
                   row_hess=ROW_HESS, handle=None):
    """fmm.direct!(target_buffer, target_index, switch, source_system, source_buffer,
    source_index) -- src/FLOWVPM_fmm.jl:102-168.  Index arguments are half-open 0-based
    (start, stop) ranges."""
    h = handle or get_handle()
    _check_matrix(target_buffer)
    _check_matrix(source_buffer)
    if source_buffer.shape[0] != 8:
        raise ValueError("source buffer must have 8 rows [x y z rho Gx Gy Gz sigma]")
    t0, t1 = target_index
    s0, s1 = source_index
    if t1 > target_buffer.shape[1] or s1 > source_buffer.shape[1]:
        raise IndexError("index range outside the buffer")
    h.check(h.lib.vpm_p2p_buffers(h.ptr, target_buffer.ctypes.data, target_buffer.shape[0], t0, t1,
                                  row_pos, row_grad, row_hess, source_buffer.ctypes.data, s0, s1,
                                  kernel.id, int(want_U), int(want_J)))


def nearfield_device(target_buffer, target_leaves, source_buffer, source_leaves, direct_list,
                     kernel, *, want_U=True, want_J=True, row_pos=ROW_POS, row_grad=ROW_GRAD,
                     row_hess=ROW_HESS, handle=None):
    """The FMM near field: nearfield_device!(...) reached from UJ_fmm when useGPU>0
    (src/FLOWVPM_UJ.jl:97; shape src/FLOWVPM_gpu.jl:637-643).

    target_leaves / source_leaves: (begin, end) int64 arrays of half-open body
    ranges in the tree-sorted buffers; direct_list: (n_pairs, 2) leaf index pairs."""
    h = handle or get_handle()
    _check_matrix(target_buffer)
    _check_matrix(source_buffer)
    tb, te = _i64(target_leaves[0]), _i64(target_leaves[1])
    sb, se = _i64(source_leaves[0]), _i64(source_leaves[1])
    pt, ps = _pairs(direct_list)
    h.check(h.lib.vpm_p2p_leafpairs(
        h.ptr, target_buffer.ctypes.data, target_buffer.shape[0], target_buffer.shape[1],
        row_pos, row_grad, row_hess, source_buffer.ctypes.data, source_buffer.shape[1],
        tb.ctypes.data, te.ctypes.data, len(tb), sb.ctypes.data, se.ctypes.data, len(sb),
        pt.ctypes.data, ps.ctypes.data, len(pt), kernel.id, int(want_U), int(want_J)))


def fmm_nearfield_device(target_system, target_indices, derivatives_switch, source_system, source_indices,
                         *, handle=None):
    """fmm.nearfield_device!(target_system, target_indices, switch, source_system, source_indices) in the
    call shape the reference shows (src/FLOWVPM_gpu.jl:637-643; reached from UJ_fmm with useGPU > 0,
    src/FLOWVPM_UJ.jl:97).  Both systems are ParticleFields; `target_indices[k]` is a `range` of particle
    columns (0-based here, 1-based UnitRange in Julia), `source_indices[k]` the source range -- or the list
    of source ranges that combine_source_indices (src/FLOWVPM_gpu.jl:554-580) gathered for that target
    leaf.  `derivatives_switch` = (PS, VS, GS) as in fmm.DerivativesSwitch{PS,VS,GS}.  Accumulates U and J
    of the target particles (no reset)."""
    h = handle or get_handle()
    TP, SP = target_system.particles, source_system.particles
    _check_matrix(TP)
    _check_matrix(SP)
    if len(target_indices) != len(source_indices):
        raise ValueError("target_indices and source_indices must have one entry per target leaf")
    _, VS, GS = derivatives_switch
    tb = _i64([r.start for r in target_indices])
    te = _i64([r.stop for r in target_indices])
    groups = [[g] if isinstance(g, range) else list(g) for g in source_indices]
    for r in list(target_indices) + [r for g in groups for r in g]:
        if r.step != 1:
            raise ValueError("index ranges must have unit step")
    soff = _i64(np.concatenate([[0], np.cumsum([len(g) for g in groups])]))
    sb = _i64([r.start for g in groups for r in g])
    se = _i64([r.stop for g in groups for r in g])
    h.check(h.lib.vpm_nearfield_ranges(
        h.ptr, TP.ctypes.data, TP.shape[0], target_system.np, tb.ctypes.data, te.ctypes.data, len(tb),
        SP.ctypes.data, SP.shape[0], source_system.np, sb.ctypes.data, se.ctypes.data, soff.ctypes.data,
        source_system.kernel.id, int(bool(VS)), int(bool(GS))))


def Estr_fmm(pfield, target_sort_index, source_sort_index, target_leaves, source_leaves,
             direct_list, *, handle=None, no_farfield_shortcut=False):
    """Estr_fmm!(target_pfield, source_pfield, target_tree, source_tree, direct_list)
    -- src/FLOWVPM_subfilterscale_models.jl:94-188 with target == source field."""
    h = handle or get_handle()
    P = pfield.particles
    _check_matrix(P)
    ts, ss = _i64(target_sort_index), _i64(source_sort_index)
    if len(ts) != pfield.np or len(ss) != pfield.np:
        raise ValueError("sort index length must equal pfield.np")
    tb, te = _i64(target_leaves[0]), _i64(target_leaves[1])
    sb, se = _i64(source_leaves[0]), _i64(source_leaves[1])
    pt, ps = _pairs(direct_list)
    flags = _flags(pfield, True, False, False, no_farfield_shortcut)
    h.check(h.lib.vpm_estr_leafpairs(
        h.ptr, P.ctypes.data, P.shape[0], pfield.np, ts.ctypes.data, ss.ctypes.data,
        tb.ctypes.data, te.ctypes.data, len(tb), sb.ctypes.data, se.ctypes.data, len(sb),
        pt.ctypes.data, ps.ctypes.data, len(pt), pfield.kernel.id, flags))


def leaf_lists(pfield, ncrit=64, theta=0.4, *, handle=None, fetch=True):
    """Device-built leaf lists of `pfield` (vpm_leaflists_build): sort index, leaf ranges and the
    near-field direct_list by the theta-MAC; they stay resident for `UJ_nearfield`.  With
    fetch=True they are also returned as dict(sort_index, leaf_begin, leaf_end, direct_list,
    pair_tgt, pair_src) -- the last two are the columns of direct_list as contiguous arrays."""
    h = handle or get_handle()
    P = pfield.particles
    _check_matrix(P)
    nl, npairs = C.c_int64(), C.c_int64()
    h.check(h.lib.vpm_leaflists_build(h.ptr, P.ctypes.data, P.shape[0], pfield.np, int(ncrit), float(theta),
                                      C.byref(nl), C.byref(npairs)))
    if not fetch:
        return dict(n_leaves=nl.value, n_pairs=npairs.value)
    sort_index = np.empty(pfield.np, dtype=np.int64)
    lb, le = np.empty(nl.value, dtype=np.int64), np.empty(nl.value, dtype=np.int64)
    pt, ps = np.empty(npairs.value, dtype=np.int32), np.empty(npairs.value, dtype=np.int32)
    h.check(h.lib.vpm_leaflists_get(h.ptr, sort_index.ctypes.data, lb.ctypes.data, le.ctypes.data,
                                    pt.ctypes.data, ps.ctypes.data))
    return dict(sort_index=sort_index, leaf_begin=lb, leaf_end=le,
                direct_list=np.ascontiguousarray(np.stack([pt, ps], axis=1)), pair_tgt=pt, pair_src=ps)


def UJ_nearfield(pfield, *, reset=True, handle=None, no_farfield_shortcut=False):
    """Near-field half of UJ_fmm (src/FLOWVPM_UJ.jl:62-129) over the leaf lists last built by
    `leaf_lists(pfield, ...)`, entirely on the device(s); adds to rows U, J (after
    _reset_particles when reset=True)."""
    h = handle or get_handle()
    P = pfield.particles
    _check_matrix(P)
    flags = (_cabi.FLAG_RESET if reset else 0) | (_cabi.FLAG_NO_FARFIELD_SHORTCUT if no_farfield_shortcut else 0)
    h.check(h.lib.vpm_uj_nearfield(h.ptr, P.ctypes.data, P.shape[0], pfield.np, pfield.kernel.id, flags))


def zeta_direct(pfield, *, handle=None):
    """zeta_direct(pfield) -- src/FLOWVPM_viscous.jl:488-515: J[1:3] of every particle <-
    sum_j Gamma_j zeta_sigma_j(x_i - x_j) (the vorticity the particle field represents)."""
    h = handle or get_handle()
    P = pfield.particles
    _check_matrix(P)
    h.check(h.lib.vpm_zeta_direct(h.ptr, P.ctypes.data, P.shape[0], pfield.np, pfield.kernel.id))


def zeta_fmm(pfield, sort_index, leaves, direct_list, *, handle=None):
    """zeta_fmm(pfield) -- src/FLOWVPM_viscous.jl:523-558 given the tree's sort index, leaf
    body ranges and near-field list (the tree itself is FastMultipole's, external)."""
    h = handle or get_handle()
    P = pfield.particles
    _check_matrix(P)
    si = _i64(sort_index)
    lb, le = _i64(leaves[0]), _i64(leaves[1])
    pa, pb = _pairs(direct_list)
    h.check(h.lib.vpm_zeta_leafpairs(h.ptr, P.ctypes.data, P.shape[0], pfield.np, si.ctypes.data,
                                     lb.ctypes.data, le.ctypes.data, len(lb), pa.ctypes.data,
                                     pb.ctypes.data, len(pa), pfield.kernel.id))


class ResidentField:
    """Device-resident mirror of pfield.particles: the reference's `nextstep`
    integration call (`euler` / `rungekutta3` for ReformulatedVPM{f,g}, NoSFS or ConstantSFS,
    Pedrizzetti relaxations, Inviscid) runs on the GPU; the matrix crosses PCIe only on
    upload() / download().  DynamicSFS and viscous schemes are not covered: use UJ_direct
    in the UJ slot for those."""

    INTEGRATIONS = {"euler": 0, "rungekutta3": 1}
    RELAXATIONS = {None: 0, "none": 0, "pedrizzetti": 1, "correctedpedrizzetti": 2}

    def __init__(self, pfield, handle=None):
        self.pfield = pfield
        self.h = handle or get_handle()
        P = pfield.particles
        _check_matrix(P)
        if P.dtype != np.float64:
            raise TypeError("ResidentField is FP64 only")
        self.upload()

    def upload(self):
        P = self.pfield.particles
        self.h.check(self.h.lib.vpm_field_upload(self.h.ptr, P.ctypes.data, P.shape[0], self.pfield.np))

    def download(self):
        P = self.pfield.particles
        self.h.check(self.h.lib.vpm_field_download(self.h.ptr, P.ctypes.data, P.shape[0], self.pfield.np))

    def UJ(self, *, sfs=False, reset=True, reset_sfs=False):
        self.h.check(self.h.lib.vpm_field_uj(self.h.ptr, self.pfield.kernel.id,
                                             _flags(self.pfield, sfs, reset, reset_sfs)))

    SFS_SCHEMES = {False: 0, None: 0, "none": 0, True: 1, "constant": 1, "dynamic": 2}
    ZETA_METHODS = {"direct": 0, "zeta_direct": 0, "fmm": 1, "zeta_fmm": 1, "fmm_reset": 2}

    def zeta_method(self, zeta="direct", ncrit=50, theta=0.4):
        """CoreSpreading's `zeta` argument (src/FLOWVPM_viscous.jl:63-141) for this field: "direct" (zeta_direct),
        "fmm" (zeta_fmm, :523-558: near field of leaf lists with pfield.fmm's ncrit / theta, J[1:3] accumulated
        on as the reference does) or "fmm_reset" (same sums, J[1:3] zeroed first).  Used by rbf_conjugategradient,
        zeta() and the CoreSpreading branch of nextstep until changed."""
        self.h.check(self.h.lib.vpm_field_zeta_method(self.h.ptr, self.ZETA_METHODS[zeta], int(ncrit), float(theta)))

    def zeta(self):
        """cs.zeta(pfield) on the resident matrix (results in J[1:3])"""
        self.h.check(self.h.lib.vpm_field_zeta(self.h.ptr, self.pfield.kernel.id))

    def rbf_conjugategradient(self, itmax=15, tol=1e-3, iterror=True):
        """rbf_conjugategradient(pfield, cs) with cs.zeta = zeta_direct (src/FLOWVPM_viscous.jl:309-478):
        target vorticity in M[7:9]; returns (iterations, final relative residuals)"""
        it = C.c_int32(0)
        res = (C.c_double * 3)()
        self.h.check(self.h.lib.vpm_field_rbf(self.h.ptr, self.pfield.kernel.id, int(itmax), float(tol), int(iterror),
                                              C.byref(it), res))
        return it.value, np.array(res[:])

    @property
    def t_sgm(self):
        v = C.c_double(0.0)
        self.h.check(self.h.lib.vpm_field_tsgm(self.h.ptr, C.byref(v), 0))
        return v.value

    @t_sgm.setter
    def t_sgm(self, value):
        v = C.c_double(float(value))
        self.h.check(self.h.lib.vpm_field_tsgm(self.h.ptr, C.byref(v), 1))

    def nextstep(self, dt, *, integration="rungekutta3", f=0.0, g=0.2, Uinf=(0.0, 0.0, 0.0), sfs=False, Cs=1.0,
                 clip_backscatter=False, relaxation="pedrizzetti", relax=True, rlxf=0.3, alpha=0.667,
                 sfs_rlxf=0.005, minC=0.0, maxC=1.0, force_positive=False, control_directional=False,
                 control_magnitude=False, viscous=None):
        """sfs: False | "constant" (ConstantSFS, coefficient Cs) | "dynamic" (DynamicSFS with the
        pseudo-3-level procedure: alpha, sfs_rlxf, minC, maxC, force_positive);
        viscous: None (Inviscid) | dict(nu=, sgm0=, beta=1.5, itmax=15, tol=1e-3, iterror=True[, zeta="direct" |
        "fmm" | "fmm_reset", ncrit=50, theta=0.4]) for CoreSpreading(nu, sgm0, zeta) (src/FLOWVPM_viscous.jl:63-223;
        without `zeta` the method set by zeta_method() stays) |
        dict(scheme="pse", nu=, recalculate_vols=True) for ParticleStrengthExchange (:228-298)"""
        if minC < 0 or maxC < 0 or minC > maxC:
            raise ValueError(f"Invalid C bounds: minC={minC}, maxC={maxC}")  # subfilterscale.jl:456-462
        sp = _cabi.VpmStepParams()
        sp.dt, sp.f, sp.g, sp.Cs, sp.rlxf = dt, f, g, Cs, rlxf
        sp.alpha, sp.sfs_rlxf, sp.minC, sp.maxC = alpha, sfs_rlxf, minC, maxC
        sp.force_positive = int(force_positive)
        sp.controls = int(control_directional) | (int(control_magnitude) << 1)
        sp.deltat = self.pfield.t / self.pfield.nt if self.pfield.nt > 0 else 0.0   # subfilterscale.jl:344-347
        sp.Uinf[0], sp.Uinf[1], sp.Uinf[2] = Uinf
        sp.kernel_id = self.pfield.kernel.id
        sp.integration = self.INTEGRATIONS[integration]
        sp.relaxation = self.RELAXATIONS[relaxation]
        sp.relax, sp.sfs, sp.clip_backscatter = int(relax), self.SFS_SCHEMES[sfs], int(clip_backscatter)
        sp.transposed = int(self.pfield.transposed)
        if viscous is not None and viscous.get("scheme", "corespreading") == "pse":
            # ParticleStrengthExchange(nu; recalculate_vols): the per-particle part, src/FLOWVPM_viscous.jl:257-298
            sp.viscous = 2 if viscous.get("recalculate_vols", True) else 3
            sp.nu = viscous["nu"]
        elif viscous is not None:
            sp.viscous = 1
            sp.nu, sp.sgm0 = viscous["nu"], viscous["sgm0"]
            sp.cs_beta, sp.cs_tol = viscous.get("beta", 1.5), viscous.get("tol", 1e-3)
            sp.cs_itmax, sp.cs_iterror = int(viscous.get("itmax", 15)), int(viscous.get("iterror", True))
            if "zeta" in viscous:
                self.zeta_method(viscous["zeta"], viscous.get("ncrit", 50), viscous.get("theta", 0.4))
        self.h.check(self.h.lib.vpm_field_step(self.h.ptr, C.byref(sp)))
        self.pfield.t += dt
        self.pfield.nt += 1


def source_system_to_buffer(pfield):
    """fmm.source_system_to_buffer! for every particle (src/FLOWVPM_fmm.jl:62-71) with the
    default rho/sigma = 1 (autotune_reg_error off); returns the 8 x np buffer."""
    P = pfield.live()
    buf = np.zeros((8, pfield.np), dtype=np.float64, order="F")
    buf[0:3] = P[0:3]
    buf[3] = P[6]
    buf[4:7] = P[3:6]
    buf[7] = P[6]
    return buf


def buffer_to_target_system(pfield, target_buffer, *, row_grad=ROW_GRAD, row_hess=ROW_HESS):
    """fmm.buffer_to_target_system! (src/FLOWVPM_fmm.jl:170-176): U += gradient, J += hessian"""
    P = pfield.live()
    P[9:12] += target_buffer[row_grad:row_grad + 3, : pfield.np]
    P[15:24] += target_buffer[row_hess:row_hess + 9, : pfield.np]
