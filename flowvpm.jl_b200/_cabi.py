"""ctypes binding of libvpm_cuda.so (include/vpm_cuda.h).

This is the same boundary the Julia shim (julia/FLOWVPMCuda.jl) reaches through
`ccall`.  There is no fallback: if the library or a CUDA device is missing the
calls raise.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libvpm_cuda.so")

# every symbol include/vpm_cuda.h declares (tests check the .so exports them all)
SYMBOLS = [
    "vpm_create", "vpm_destroy", "vpm_last_error", "vpm_abi_version", "vpm_num_devices",
    "vpm_uj_direct", "vpm_uj_direct_f32", "vpm_uj_direct_st",
    "vpm_upload_state", "vpm_eval", "vpm_download_results",
    "vpm_pin_host", "vpm_unpin_host", "vpm_set_option", "vpm_nearfield_ranges", "vpm_field_zeta_method",
    "vpm_field_zeta",
    "vpm_p2p_buffers", "vpm_p2p_leafpairs", "vpm_estr_leafpairs",
    "vpm_zeta_direct", "vpm_zeta_leafpairs",
    "vpm_leaflists_build", "vpm_leaflists_get", "vpm_uj_nearfield",
    "vpm_field_upload", "vpm_field_download", "vpm_field_uj", "vpm_field_step", "vpm_field_rbf",
    "vpm_field_tsgm",
    "vpm_uj_device", "vpm_sfs_device",
    "vpm_get_timing", "vpm_measure_dfma_peak", "vpm_measure_ffma_peak", "vpm_test_math", "vpm_plan_query",
]

VPM_OK = 0
KERNEL_SINGULAR, KERNEL_GAUSSIAN, KERNEL_GAUSSIANERF, KERNEL_WINCKELMANS = 0, 1, 2, 3
FLAG_RESET, FLAG_RESET_SFS, FLAG_SFS, FLAG_TRANSPOSED, FLAG_NO_FARFIELD_SHORTCUT = 1, 2, 4, 8, 16
FLAG_FP32 = 32
OPT_NEARFIELD_FP32 = 1
OPT_UJ_VARIANT, OPT_SFS_VARIANT, OPT_UJ_CONST, OPT_UJ_TABLE, OPT_SMALL_GRAPH = 2, 3, 4, 5, 6


class VpmTiming(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("h2d_ms", "prep_ms", "uj_ms", "sfs_ms", "finish_ms", "d2h_ms", "total_ms")] + \
               [("uj_pairs", C.c_int64), ("sfs_pairs", C.c_int64),
                ("kernel_launches", C.c_int32), ("n_gpus", C.c_int32)]


class VpmStepParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("f", C.c_double), ("g", C.c_double), ("Uinf", C.c_double * 3),
                ("Cs", C.c_double), ("rlxf", C.c_double),
                ("alpha", C.c_double), ("sfs_rlxf", C.c_double), ("minC", C.c_double), ("maxC", C.c_double),
                ("deltat", C.c_double), ("nu", C.c_double), ("sgm0", C.c_double), ("cs_beta", C.c_double),
                ("cs_tol", C.c_double),
                ("kernel_id", C.c_int32), ("integration", C.c_int32), ("relaxation", C.c_int32),
                ("relax", C.c_int32), ("sfs", C.c_int32), ("clip_backscatter", C.c_int32),
                ("transposed", C.c_int32), ("force_positive", C.c_int32), ("controls", C.c_int32),
                ("viscous", C.c_int32), ("cs_itmax", C.c_int32), ("cs_iterror", C.c_int32)]


class VpmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libvpm_cuda error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """dlopen libvpm_cuda.so and declare the prototypes; raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VpmError(-4, f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                           "(the CUDA extension is the only implementation; there is no CPU path)")
    lib = C.CDLL(LIB_PATH)
    p, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    P = C.POINTER
    lib.vpm_create.argtypes = [P(p), i32, P(i32)]
    lib.vpm_destroy.argtypes = [p]
    lib.vpm_last_error.argtypes = [p]
    lib.vpm_last_error.restype = C.c_char_p
    lib.vpm_abi_version.argtypes = []
    lib.vpm_num_devices.argtypes = [p]
    lib.vpm_uj_direct.argtypes = [p, p, i64, i64, i32, i32]
    lib.vpm_uj_direct_f32.argtypes = [p, p, i64, i64, i32, i32]
    lib.vpm_uj_direct_st.argtypes = [p, p, i64, i64, p, i64, i64, i32]
    lib.vpm_upload_state.argtypes = [p, p, i64, i64]
    lib.vpm_eval.argtypes = [p, i32, i32]
    lib.vpm_download_results.argtypes = [p, p, i64, i64, i32]
    lib.vpm_set_option.argtypes = [p, i32, i32]
    lib.vpm_nearfield_ranges.argtypes = [p, p, i64, i64, p, p, i64, p, i64, i64, p, p, p, i32, i32, i32]
    lib.vpm_pin_host.argtypes = [p, p, C.c_size_t]
    lib.vpm_unpin_host.argtypes = [p, p]
    lib.vpm_p2p_buffers.argtypes = [p, p, i64, i64, i64, i32, i32, i32, p, i64, i64, i32, i32, i32]
    lib.vpm_p2p_leafpairs.argtypes = [p, p, i64, i64, i32, i32, i32, p, i64, p, p, i64, p, p, i64,
                                      p, p, i64, i32, i32, i32]
    lib.vpm_estr_leafpairs.argtypes = [p, p, i64, i64, p, p, p, p, i64, p, p, i64, p, p, i64, i32, i32]
    lib.vpm_zeta_direct.argtypes = [p, p, i64, i64, i32]
    lib.vpm_zeta_leafpairs.argtypes = [p, p, i64, i64, p, p, p, i64, p, p, i64, i32]
    lib.vpm_leaflists_build.argtypes = [p, p, i64, i64, i64, dbl, P(i64), P(i64)]
    lib.vpm_leaflists_get.argtypes = [p, p, p, p, p, p]
    lib.vpm_uj_nearfield.argtypes = [p, p, i64, i64, i32, i32]
    lib.vpm_field_upload.argtypes = [p, p, i64, i64]
    lib.vpm_field_download.argtypes = [p, p, i64, i64]
    lib.vpm_field_uj.argtypes = [p, i32, i32]
    lib.vpm_field_step.argtypes = [p, P(VpmStepParams)]
    lib.vpm_field_rbf.argtypes = [p, i32, i32, dbl, i32, P(i32), P(dbl)]
    lib.vpm_field_tsgm.argtypes = [p, P(dbl), i32]
    lib.vpm_field_zeta_method.argtypes = [p, i32, i64, dbl]
    lib.vpm_field_zeta.argtypes = [p, i32]
    lib.vpm_uj_device.argtypes = [p, p, i64, i64, i64, p, i32, i32, p]
    lib.vpm_sfs_device.argtypes = [p, p, p, p, i64, i64, i64, p, i32, i32, p]
    lib.vpm_get_timing.argtypes = [p, P(VpmTiming)]
    lib.vpm_measure_dfma_peak.argtypes = [p, P(dbl), P(dbl)]
    lib.vpm_measure_ffma_peak.argtypes = [p, i32, P(dbl), P(dbl)]
    lib.vpm_test_math.argtypes = [p, i32, i32, p, p, p, i64]
    lib.vpm_plan_query.argtypes = [i64, i64, i32, i32, p]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name != "vpm_last_error":
            fn.restype = C.c_int
    _lib = lib
    return lib


class Handle:
    """Owner of one vpm_handle (device buffers, streams, NCCL communicators)."""

    def __init__(self, n_gpus=1, device_ids=None):
        self.lib = load()
        self._h = C.c_void_p()
        ids = None
        if device_ids is not None:
            ids = (C.c_int * len(device_ids))(*device_ids)
            n_gpus = len(device_ids)
        rc = self.lib.vpm_create(C.byref(self._h), int(n_gpus), ids)
        if rc != VPM_OK:
            msg = self.lib.vpm_last_error(None).decode()
            self._h = C.c_void_p()
            raise VpmError(rc, msg)

    @property
    def ptr(self):
        return self._h

    def check(self, rc):
        if rc != VPM_OK:
            raise VpmError(rc, self.lib.vpm_last_error(self._h).decode())

    def set_option(self, option, value):
        self.check(self.lib.vpm_set_option(self._h, int(option), int(value)))

    def timing(self):
        t = VpmTiming()
        self.check(self.lib.vpm_get_timing(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in VpmTiming._fields_}

    def close(self):
        if self._h:
            self.lib.vpm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
