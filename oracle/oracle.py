"""ctypes wrapper of oracle/libvpm_oracle.so (the C restatement of the reference).

TEST INFRASTRUCTURE ONLY -- see oracle/vpm_oracle.c for what is restated and how
the oracle is pinned.  Builds the shared object on first use with oracle/Makefile.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvpm_oracle.so")

KERNEL_IDS = {"singular": 0, "gaussian": 1, "gaussianerf": 2, "winckelmans": 3}
FLAG_RESET, FLAG_RESET_SFS, FLAG_SFS, FLAG_TRANSPOSED = 1, 2, 4, 8

_lib = None


def build(force=False):
    src = os.path.join(HERE, "vpm_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(LIB_PATH):
        res = subprocess.run(["make", "-C", HERE, "-B", "libvpm_oracle.so"], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        p, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
        L.vpm_oracle_erf64.argtypes = [dbl]
        L.vpm_oracle_erf64.restype = dbl
        L.vpm_oracle_g_dgdr.argtypes = [i32, dbl, C.POINTER(dbl), C.POINTER(dbl)]
        L.vpm_oracle_g_dgdr.restype = None
        L.vpm_oracle_zeta.argtypes = [i32, dbl]
        L.vpm_oracle_zeta.restype = dbl
        L.vpm_oracle_direct_buffers.argtypes = [p, i64, i64, i64, p, i64, i64, i32, i32, i32]
        L.vpm_oracle_direct_buffers.restype = None
        L.vpm_oracle_erf32.argtypes = [C.c_float]
        L.vpm_oracle_erf32.restype = dbl
        L.vpm_oracle_direct_buffers_f32.argtypes = [p, i64, i64, i64, p, i64, i64, i32, i32, i32]
        L.vpm_oracle_direct_buffers_f32.restype = None
        L.vpm_oracle_direct_buffers_mt.argtypes = [p, i64, i64, i64, p, i64, i64, i32, i32, i32, i32]
        L.vpm_oracle_direct_buffers_mt.restype = None
        L.vpm_oracle_reset_particles.argtypes = [p, i64, i64]
        L.vpm_oracle_reset_particles_sfs.argtypes = [p, i64, i64]
        L.vpm_oracle_estr_direct.argtypes = [p, i64, i64, i32, i32, i32]
        L.vpm_oracle_estr_direct.restype = None
        L.vpm_oracle_estr_direct_targets.argtypes = [p, i64, i64, p, i64, i32, i32, i32]
        L.vpm_oracle_estr_direct_targets.restype = None
        L.vpm_oracle_uj_direct.argtypes = [p, i64, i64, i32, i32, i32]
        L.vpm_oracle_uj_direct.restype = i32
        L.vpm_oracle_direct_leafpairs.argtypes = [p, i64, p, p, p, p, p, p, p, i64, i32, i32, i32]
        L.vpm_oracle_direct_leafpairs.restype = None
        L.vpm_oracle_estr_leafpairs.argtypes = [p, i64, p, p, p, p, p, p, p, p, i64, i32, i32]
        L.vpm_oracle_estr_leafpairs.restype = None
        L.vpm_oracle_zeta_direct.argtypes = [p, i64, i64, i32, i32]
        L.vpm_oracle_zeta_direct.restype = None
        L.vpm_oracle_zeta_leafpairs.argtypes = [p, i64, p, p, p, p, p, i64, i32]
        L.vpm_oracle_zeta_leafpairs.restype = None
        L.vpm_oracle_field_step.argtypes = [p, i64, i64, p, p, i32]
        L.vpm_oracle_field_step.restype = i32
        L.vpm_oracle_rbf_cg.argtypes = [p, i64, i64, i32, i32, dbl, i32, p, i32]
        L.vpm_oracle_rbf_cg.restype = i32
        L.vpm_oracle_set_cs_zeta.argtypes = [p]
        L.vpm_oracle_set_cs_zeta.restype = None
        L.vpm_oracle_max_threads.restype = i32
        L.vpm_oracle_num_procs.restype = i32
        _lib = L
    return _lib


def _kid(kernel):
    if isinstance(kernel, str):
        return KERNEL_IDS[kernel]
    return int(getattr(kernel, "id", kernel))


def _f(P):
    assert isinstance(P, np.ndarray) and P.dtype == np.float64 and P.flags.f_contiguous, \
        "oracle works on Fortran-ordered float64 matrices"
    return P


def max_threads():
    return int(lib().vpm_oracle_max_threads())


def num_procs():
    """host cores available to this process, ignoring OMP_NUM_THREADS (bench.py's CPU legs)"""
    return int(lib().vpm_oracle_num_procs())


def erf64(x):
    return float(lib().vpm_oracle_erf64(float(x)))


def erf32(x):
    """custom_erf32 (Float32 in; the last branch returns a Float64, as in the reference)"""
    return float(lib().vpm_oracle_erf32(C.c_float(float(x))))


def direct_buffers_f32(tgt, t0, t1, src, s0, s1, kernel, want_U=True, want_J=True):
    """fmm.direct! on Float32 buffers with Julia's promotion rules (a ParticleField{Float32}), in place"""
    assert tgt.dtype == np.float32 and src.dtype == np.float32 and tgt.flags.f_contiguous and src.flags.f_contiguous
    assert src.shape[0] == 8 and tgt.shape[0] >= 16
    lib().vpm_oracle_direct_buffers_f32(tgt.ctypes.data, tgt.shape[0], int(t0), int(t1), src.ctypes.data,
                                        int(s0), int(s1), _kid(kernel), int(want_U), int(want_J))


def g_dgdr(kernel, s):
    g, dg = C.c_double(), C.c_double()
    lib().vpm_oracle_g_dgdr(_kid(kernel), float(s), C.byref(g), C.byref(dg))
    return g.value, dg.value


def zeta(kernel, s):
    return float(lib().vpm_oracle_zeta(_kid(kernel), float(s)))


def flags(sfs=False, reset=True, reset_sfs=False, transposed=True):
    return (FLAG_RESET * bool(reset)) | (FLAG_RESET_SFS * bool(reset_sfs)) | (FLAG_SFS * bool(sfs)) | \
        (FLAG_TRANSPOSED * bool(transposed))


def uj_direct(P, np_, kernel, *, sfs=False, reset=True, reset_sfs=False, transposed=True, nthreads=0):
    """UJ_direct(pfield; sfs, reset, reset_sfs) on the 46 x N matrix P, in place."""
    _f(P)
    nthreads = nthreads or max_threads()
    rc = lib().vpm_oracle_uj_direct(P.ctypes.data, P.shape[0], int(np_), _kid(kernel),
                                    flags(sfs, reset, reset_sfs, transposed), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle uj_direct failed ({rc})")


def estr_direct(P, np_, kernel, transposed=True, nthreads=0):
    _f(P)
    lib().vpm_oracle_estr_direct(P.ctypes.data, P.shape[0], int(np_), _kid(kernel), int(transposed),
                                 int(nthreads or max_threads()))


def estr_direct_targets(P, np_, targets, kernel, transposed=True, nthreads=0):
    """Estr_direct for the listed target particles only (all non-static sources); adds into P's SFS rows."""
    _f(P)
    t = np.ascontiguousarray(targets, dtype=np.int64)
    lib().vpm_oracle_estr_direct_targets(P.ctypes.data, P.shape[0], int(np_), t.ctypes.data, t.size, _kid(kernel),
                                         int(transposed), int(nthreads or max_threads()))


def uj_slice(P, np_, targets, kernel, *, sfs=True, transposed=True, nthreads=0):
    """Reference values of U, J (and SFS) for the listed targets of the field P, every particle of P being a
    source: U, J by fmm.direct! on buffers; SFS by Estr_direct over P AS IT STANDS (the J rows of P are the
    velocity gradients the sweep reads -- the caller passes the field whose J the GPU just computed, so that
    the SFS comparison is on identical inputs).  Returns (U[3,k], J[9,k], SFS[3,k])."""
    _f(P)
    t = np.ascontiguousarray(targets, dtype=np.int64)
    nthreads = nthreads or num_procs()
    src = np.zeros((8, np_), order="F")
    src[0:3], src[4:7], src[7] = P[0:3, :np_], P[3:6, :np_], P[6, :np_]
    src[3] = P[6, :np_]
    tb = np.zeros((16, t.size), order="F")
    tb[0:3] = P[0:3, t]
    direct_buffers(tb, 0, t.size, src, 0, np_, kernel, True, True, nthreads)
    sfs_out = None
    if sfs:
        Q = P  # in place on the SFS rows of the targets only (restored below)
        before = Q[39:42, t].copy()
        Q[39:42, t] = 0.0
        estr_direct_targets(Q, np_, t, kernel, transposed, nthreads)
        sfs_out = Q[39:42, t].copy()
        Q[39:42, t] = before
    return tb[4:7].copy(), tb[7:16].copy(), sfs_out


def direct_buffers(tgt, t0, t1, src, s0, s1, kernel, want_U=True, want_J=True, nthreads=1):
    """fmm.direct! on FastMultipole-style buffers (16-row target, 8-row source), in place."""
    _f(tgt)
    _f(src)
    assert src.shape[0] == 8 and tgt.shape[0] >= 16
    lib().vpm_oracle_direct_buffers_mt(tgt.ctypes.data, tgt.shape[0], int(t0), int(t1), src.ctypes.data,
                                       int(s0), int(s1), _kid(kernel), int(want_U), int(want_J),
                                       int(nthreads))


def _viscous_id(viscous):
    """None -> 0 Inviscid; dict without "scheme" or scheme="corespreading" -> 1; scheme="pse" -> 2 (3 when
    recalculate_vols=False)"""
    if viscous is None:
        return 0
    if viscous.get("scheme", "corespreading") == "pse":
        return 2 if viscous.get("recalculate_vols", True) else 3
    return 1


def field_step(P, np_, kernel, dt, *, integration="rungekutta3", f=0.0, g=0.2, Uinf=(0.0, 0.0, 0.0), sfs=False,
               Cs=1.0, clip_backscatter=False, relaxation="pedrizzetti", relax=True, rlxf=0.3, transposed=True,
               alpha=0.667, sfs_rlxf=0.005, minC=0.0, maxC=1.0, force_positive=False, control_directional=False,
               control_magnitude=False, deltat=0.0, viscous=None, nthreads=0):
    """one euler / rungekutta3 step of ReformulatedVPM{f,g} on the 46 x N matrix, in place;
    sfs: False | "constant" | "dynamic" (pseudo-3-level procedure)"""
    _f(P)
    v = viscous or {}
    dp = np.array([dt, f, g, Uinf[0], Uinf[1], Uinf[2], Cs, rlxf, alpha, sfs_rlxf, minC, maxC, deltat,
                   v.get("nu", 0.0), v.get("sgm0", 1.0), v.get("beta", 1.5), v.get("tol", 1e-3), v.get("t_sgm", 0.0)],
                  dtype=np.float64)
    ip = np.array([_kid(kernel), {"euler": 0, "rungekutta3": 1}[integration],
                   {None: 0, "none": 0, "pedrizzetti": 1, "correctedpedrizzetti": 2}[relaxation], int(relax),
                   {False: 0, None: 0, "none": 0, True: 1, "constant": 1, "dynamic": 2}[sfs],
                   int(clip_backscatter), int(transposed), int(force_positive),
                   int(control_directional) | (int(control_magnitude) << 1),
                   _viscous_id(viscous), int(v.get("itmax", 15)), int(v.get("iterror", True))], dtype=np.int32)
    rc = lib().vpm_oracle_field_step(P.ctypes.data, P.shape[0], int(np_), dp.ctypes.data, ip.ctypes.data,
                                     int(nthreads or max_threads()))
    if rc != 0:
        raise RuntimeError(f"oracle field_step failed ({rc})")
    if viscous is not None and _viscous_id(viscous) == 1:
        viscous["t_sgm"] = float(dp[17])   # CoreSpreading.t_sgm: time since the last core reset


def rbf_conjugategradient(P, np_, kernel, itmax=15, tol=1e-3, iterror=True, nthreads=0):
    """rbf_conjugategradient with cs.zeta = zeta_direct; returns (iterations, final relative residuals)"""
    _f(P)
    info = np.zeros(3)
    rc = lib().vpm_oracle_rbf_cg(P.ctypes.data, P.shape[0], int(np_), _kid(kernel), int(itmax), float(tol),
                                 int(iterror), info.ctypes.data, int(nthreads or max_threads()))
    if rc < 0:
        raise RuntimeError("Maximum number of iterations reached before convergence")
    return rc, info


_ZETA_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int64, C.c_int)
_zeta_cb_keepalive = None


class cs_zeta_fmm:
    """Context manager: while active, rbf_conjugategradient and CoreSpreading (field_step) evaluate
    cs.zeta = zeta_fmm (src/FLOWVPM_viscous.jl:523-558) instead of zeta_direct: leaf lists for the CURRENT
    X and sigma from oracle/leaflists.py (the recipe the device builder restates; FastMultipole's tree is
    not in the reference repo), then the list loop.  reset=False is the reference's behaviour (J[1:3] is
    accumulated on); reset=True zeroes J[1:3] first."""

    def __init__(self, ncrit=50, theta=0.4, reset=False):
        self.ncrit, self.theta, self.reset = ncrit, theta, reset
        self.calls = 0

    def _eval(self, ptr, nf, np_, kernel_id):
        from . import leaflists
        P = np.ctypeslib.as_array((C.c_double * (nf * np_)).from_address(ptr)).reshape((nf, np_), order="F")
        ll = leaflists.build_leaf_lists(P[0:3].copy(), P[6].copy(), ncrit=self.ncrit, theta=self.theta)
        if self.reset:
            P[15:18] = 0.0
        si, lb, le = _i64(ll["sort_index"]), _i64(ll["leaf_begin"]), _i64(ll["leaf_end"])
        dl = np.asarray(ll["direct_list"])
        pt, ps = _i32(dl[:, 0]), _i32(dl[:, 1])
        lib().vpm_oracle_zeta_leafpairs(ptr, nf, si.ctypes.data, lb.ctypes.data, le.ctypes.data, pt.ctypes.data,
                                        ps.ctypes.data, len(pt), kernel_id)
        self.calls += 1

    def __enter__(self):
        global _zeta_cb_keepalive
        _zeta_cb_keepalive = _ZETA_CB(self._eval)
        lib().vpm_oracle_set_cs_zeta(C.cast(_zeta_cb_keepalive, C.c_void_p))
        return self

    def __exit__(self, *exc):
        global _zeta_cb_keepalive
        lib().vpm_oracle_set_cs_zeta(None)
        _zeta_cb_keepalive = None
        return False


def zeta_direct(P, np_, kernel, nthreads=0):
    _f(P)
    lib().vpm_oracle_zeta_direct(P.ctypes.data, P.shape[0], int(np_), _kid(kernel), int(nthreads or max_threads()))


def zeta_leafpairs(P, sort, leaves, direct_list, kernel):
    _f(P)
    si, lb, le = _i64(sort), _i64(leaves[0]), _i64(leaves[1])
    dl = np.asarray(direct_list)
    pt, ps = _i32(dl[:, 0]), _i32(dl[:, 1])
    lib().vpm_oracle_zeta_leafpairs(P.ctypes.data, P.shape[0], si.ctypes.data, lb.ctypes.data, le.ctypes.data,
                                    pt.ctypes.data, ps.ctypes.data, len(pt), _kid(kernel))


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def direct_leafpairs(tgt, src, tleaves, sleaves, direct_list, kernel, want_U=True, want_J=True):
    _f(tgt)
    _f(src)
    tb, te, sb, se = _i64(tleaves[0]), _i64(tleaves[1]), _i64(sleaves[0]), _i64(sleaves[1])
    dl = np.asarray(direct_list)
    pt, ps = _i32(dl[:, 0]), _i32(dl[:, 1])
    lib().vpm_oracle_direct_leafpairs(tgt.ctypes.data, tgt.shape[0], src.ctypes.data, tb.ctypes.data,
                                      te.ctypes.data, sb.ctypes.data, se.ctypes.data, pt.ctypes.data,
                                      ps.ctypes.data, len(pt), _kid(kernel), int(want_U), int(want_J))


def estr_leafpairs(P, tsort, ssort, tleaves, sleaves, direct_list, kernel, transposed=True):
    _f(P)
    ts, ss = _i64(tsort), _i64(ssort)
    tb, te, sb, se = _i64(tleaves[0]), _i64(tleaves[1]), _i64(sleaves[0]), _i64(sleaves[1])
    dl = np.asarray(direct_list)
    pt, ps = _i32(dl[:, 0]), _i32(dl[:, 1])
    lib().vpm_oracle_estr_leafpairs(P.ctypes.data, P.shape[0], ts.ctypes.data, ss.ctypes.data,
                                    tb.ctypes.data, te.ctypes.data, sb.ctypes.data, se.ctypes.data,
                                    pt.ctypes.data, ps.ctypes.data, len(pt), _kid(kernel), int(transposed))
