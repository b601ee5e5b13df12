"""ctypes binding of libdegnorm_b200.so (include/degnorm_b200.h).  There is no CPU fallback: if the library
is missing or no CUDA device is usable, importing callers get a loud error."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DEGNORM_B200_LIB: tuning aid, another build of the same library (build.py variants); still no fallback
LIB_PATH = os.environ.get("DEGNORM_B200_LIB") or os.path.join(HERE, "libdegnorm_b200.so")

ABI_VERSION = 6
DN_NCOUNTERS = 8
DN_MAX_BINS = 64
DN_MAX_SAMPLES = 256
CNT_EXIT, CNT_N_HICOV, CNT_NMF_CALLS, CNT_SUM_COLS, CNT_EIG_STEPS, CNT_DROPS_LO, CNT_DROPS_HI, CNT_RESIDENT = range(8)
EXIT_NAMES = {0: "none", 1: "few_hicov", 2: "empty_sample", 3: "median", 4: "no_selection", 5: "refined",
              6: "fallback_high", 7: "fallback", -1: "plan_error"}

DN_FLAG_PLAIN_NMF, DN_FLAG_RAW_RHO = 1, 2
(DN_EXIT_NONE, DN_EXIT_FEW_HICOV, DN_EXIT_EMPTY_SAMPLE, DN_EXIT_MEDIAN, DN_EXIT_NO_SELECTION, DN_EXIT_REFINED,
 DN_EXIT_FALLBACK_HIGH, DN_EXIT_FALLBACK) = range(8)
DN_ERR_INVALID, DN_ERR_CUDA, DN_ERR_UNSUPPORTED, DN_ERR_WORKSPACE = -1, -2, -3, -4


class DnParams(C.Structure):
    _fields_ = [("p", C.c_int32), ("nmf_iter", C.c_int32), ("bins", C.c_int32), ("min_bins", C.c_int32),
                ("min_high_coverage", C.c_int32), ("downsample_rate", C.c_int32), ("min_gene_len", C.c_int32),
                ("skip_baseline_selection", C.c_int32), ("flags", C.c_int32)]


class DnPlan(C.Structure):
    _fields_ = [("tile", C.c_int32), ("threads", C.c_int32), ("ctas", C.c_int32), ("resident_cols", C.c_int32),
                ("chunk_cols", C.c_int32), ("smem_bytes", C.c_int32), ("cluster", C.c_int32), ("reserved", C.c_int32),
                ("ws_cols", C.c_int64), ("ws_bytes", C.c_int64)]


class DegnormCudaError(RuntimeError):
    pass


_lib = None

_P = C.c_void_p
_SIGS = {
    "dn_abi_version": (C.c_int, []),
    "dn_last_error": (C.c_char_p, []),
    "dn_device_info": (C.c_int, [C.POINTER(C.c_int32)] * 3),
    "dn_make_plan": (C.c_int, [C.POINTER(DnParams), C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                               C.c_int32, C.c_int32, C.POINTER(DnPlan)]),
    "dn_init_ratio_svd": (C.c_int, [_P, _P, _P, C.c_int32, C.POINTER(DnParams), C.POINTER(DnPlan), _P, _P, _P, _P, _P,
                                    C.c_int64, _P]),
    "dn_baseline_selection": (C.c_int, [_P, _P, _P, C.c_int32, C.POINTER(DnParams), C.POINTER(DnPlan), _P, _P, _P, _P,
                                        _P, _P, _P, _P, _P, _P, _P, C.c_int64, _P]),
    "dn_estimates": (C.c_int, [_P, _P, _P, C.c_int32, C.POINTER(DnParams), _P, _P, _P, _P, _P, _P, _P]),
    "dn_outer_sums": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, C.c_int64, _P]),
    "dn_outer_apply": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "dn_init_sums": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P, C.c_int64, _P]),
    "dn_init_apply": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "dn_sums_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "dn_probe_fp64": (C.c_int64, [C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "dn_probe_lds": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
}
EXPORTS = sorted(_SIGS)


def lib():
    """The loaded library (loads it on first use)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("degnorm_b200: %s is missing -- build it with `python -m degnorm_b200.build` "
                              "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        if l.dn_abi_version() != ABI_VERSION:
            raise ImportError("degnorm_b200: ABI version mismatch")
        _lib = l
    return _lib


def check(rc):
    if rc == 0:
        return
    msg = lib().dn_last_error().decode("utf-8", "replace")
    if rc == DN_ERR_INVALID:
        raise ValueError(msg)
    if rc == DN_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise DegnormCudaError("degnorm_b200 status %d: %s" % (rc, msg))


def device_info():
    sm, smem, cc = C.c_int32(), C.c_int32(), C.c_int32()
    check(lib().dn_device_info(C.byref(sm), C.byref(smem), C.byref(cc)))
    return sm.value, smem.value, cc.value
