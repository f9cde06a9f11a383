"""ctypes binding of include/tofrt.h.  No torch types cross this boundary."""
import ctypes
import os

from . import _build

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)

TOFRT_OK, E_BADARG, E_TOOFEW, E_CUDA, E_TOOLARGE, E_NOSETUP = 0, -1, -2, -3, -4, -5


class TofrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tofrt error {code}: {msg}")
        self.code = code


class Params(ctypes.Structure):
    """tofrt_params: the module globals of main_rt.py:449-455."""
    _fields_ = [("c1", ctypes.c_double), ("c2", ctypes.c_double), ("l0", ctypes.c_double),
                ("h0", ctypes.c_double), ("d", ctypes.c_double)]


class RayRecords(ctypes.Structure):
    """tofrt_ray_records: optional [G][K] per-ray outputs (NULL = not wanted)."""
    _fields_ = [("pipe_x", ctypes.c_void_p), ("pipe_z", ctypes.c_void_p), ("lens2_x", ctypes.c_void_p),
                ("lens2_z", ctypes.c_void_p), ("target_x", ctypes.c_void_p), ("idx", ctypes.c_void_p),
                ("legs", ctypes.c_void_p)]


EXPORTS = ["tofrt_version", "tofrt_last_error", "tofrt_device_count", "tofrt_create", "tofrt_destroy",
           "tofrt_setup", "tofrt_setup_shooters", "tofrt_set_option", "tofrt_get_info", "tofrt_last_stats", "tofrt_lens_host", "tofrt_table_dev", "tofrt_table_f32_dev", "tofrt_table_host", "tofrt_host_alloc", "tofrt_host_free", "tofrt_hits_host", "tofrt_debug_search_host", "tofrt_fermat_dev", "tofrt_fermat_host", "tofrt_fermat3d_setup", "tofrt_fermat3d_dev", "tofrt_fermat3d_host", "tofrt_kernel_info", "tofrt_measure_fp64_peak", "tofrt_measure_fp32_peak"]

_lib = None


def load():
    """Load libtofrt.so.  Fails loudly when the CUDA extension is missing: there
    is no CPU fallback in this package."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = ctypes.CDLL(path)
    lib.tofrt_version.restype = ctypes.c_int
    lib.tofrt_last_error.restype = ctypes.c_char_p
    lib.tofrt_device_count.restype = ctypes.c_int
    lib.tofrt_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
    lib.tofrt_destroy.argtypes = [ctypes.c_void_p]
    lib.tofrt_setup.argtypes = [ctypes.c_void_p, ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_int,
                                ctypes.c_double, ctypes.c_double]
    lib.tofrt_setup_shooters.argtypes = [ctypes.c_void_p, ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_double, ctypes.c_double]
    lib.tofrt_set_option.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.tofrt_get_info.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.tofrt_last_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    lib.tofrt_lens_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.tofrt_table_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(RayRecords), ctypes.c_void_p]
    lib.tofrt_table_f32_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.tofrt_table_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(RayRecords)]
    lib.tofrt_host_alloc.argtypes = [ctypes.c_size_t]
    lib.tofrt_host_alloc.restype = ctypes.c_void_p
    lib.tofrt_host_free.argtypes = [ctypes.c_void_p]
    lib.tofrt_host_free.restype = None
    lib.tofrt_hits_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                    ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.POINTER(ctypes.c_longlong)]
    lib.tofrt_debug_search_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.tofrt_fermat_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.tofrt_fermat_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.tofrt_fermat3d_setup.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_int]
    lib.tofrt_fermat3d_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]
    lib.tofrt_fermat3d_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.tofrt_kernel_info.argtypes = [ctypes.c_void_p] + [c_int_p] * 4
    lib.tofrt_measure_fp64_peak.argtypes = [ctypes.c_void_p, c_double_p, c_double_p]
    lib.tofrt_measure_fp32_peak.argtypes = [ctypes.c_void_p, c_double_p, c_double_p]
    for name in EXPORTS:
        getattr(lib, name)
    _lib = lib
    return lib


def check(rc):
    if rc != TOFRT_OK:
        msg = load().tofrt_last_error().decode("utf-8", "replace")
        if rc == E_TOOFEW:
            raise ValueError(msg)  # the reference raises ValueError (main_rt.py:24-29)
        if rc == E_BADARG:
            raise ValueError(msg)
        raise TofrtError(rc, msg)
