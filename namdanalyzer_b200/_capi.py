"""ctypes binding of libnda_b200.so (include/nda_b200.h).

The library is built in-tree by ``namdanalyzer_b200.build`` (nvcc, sm_100a) and loaded from
``namdanalyzer_b200/lib/libnda_b200.so``.  There is no CPU fallback: if the library is missing
or no B200 is visible, every compute entry point raises ``RuntimeError``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnda_b200.so")

c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_i32p = ctypes.POINTER(ctypes.c_int)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_int, c_float, c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_void_p

# name -> (restype, argtypes); every symbol include/nda_b200.h declares
SIGNATURES = {
    "nda_last_error": (ctypes.c_char_p, []),
    "nda_get_parallel_backend": (c_int, []),
    "nda_device_count": (c_int, []),
    "nda_set_device": (c_int, [c_int]),
    "nda_get_device": (c_int, []),
    "nda_set_devices": (c_int, [c_i32p, c_int]),
    "nda_get_devices": (c_int, [c_i32p, c_int]),
    "nda_last_collective": (c_int, []),
    "nda_set_option": (c_int, [ctypes.c_char_p, ctypes.c_char_p]),
    "nda_shutdown": (c_int, []),
    "nda_debug_tc_probe": (c_int, [ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_uint16), c_u32p]),
    "nda_measure_h2d": (c_int, [ctypes.c_size_t, c_int, c_f64p]),
    "nda_host_register": (c_int, [c_void_p, ctypes.c_size_t]),
    "nda_host_unregister": (c_int, [c_void_p]),
    "nda_launch_count": (ctypes.c_uint64, []),
    "nda_launch_count_reset": (None, []),
    "nda_set_within_pure_predicate": (c_int, [c_int]),
    "nda_debug_tc_trace": (c_int, [ctypes.POINTER(ctypes.c_longlong)]),
    "nda_last_filter": (c_int, []),
    "nda_get_within": (c_int, [c_f32p, c_int, c_int, c_i32p, c_int, c_i32p, c_int, c_i32p, c_f32p, c_float]),
    "nda_get_distances": (c_int, [c_f32p, c_int, c_f32p, c_int, c_f32p, c_f32p, c_int, c_int]),
    "nda_get_radial_nbr_density": (c_int, [c_f32p, c_int, c_f32p, c_int, c_f32p, c_int, c_f32p, c_int, c_int,
                                           c_float, c_float]),
    "nda_get_radial_nbr_counts": (c_int, [c_f32p, c_int, c_f32p, c_int, c_i32p, c_int, c_u64p, c_f32p, c_int,
                                          c_f32p, c_int, c_int, c_int, c_int, c_float]),
    "nda_get_distances_sum": (c_int, [c_f32p, c_int, c_f32p, c_int, c_f64p, c_f32p, c_int, c_int, c_int, c_int]),
    "nda_get_within_bits": (c_int, [c_f32p, c_int, c_int, c_i32p, c_int, c_i32p, c_int, c_u32p, c_i32p, c_f32p,
                                    c_float, c_int, c_int]),
    "nda_set_water_dist_pbc": (c_int, [c_f32p, c_int, c_f32p, c_int, c_f32p, c_int, c_int]),
    "nda_water_orient_at_surface": (c_int, [c_f32p, c_int, c_f32p, c_f32p, c_int, c_f32p, c_f32p, c_int,
                                            c_float, c_float, c_int]),
    "nda_get_hb_corr": (c_int, [c_f32p, c_int, c_int, c_f32p, c_int, c_f32p, c_int, c_f32p, c_f32p, c_int, c_int, c_int,
                                c_float, c_float, c_int]),
    "nda_get_hb_counts": (c_int, [c_f32p, c_int, c_int, c_f32p, c_int, c_f32p, c_int, c_f32p, c_u64p, c_float, c_float,
                                  c_int]),
    "nda_dcd_get_coor": (c_int, [ctypes.c_char_p, c_i32p, c_int, c_int, c_i32p, c_int, c_i32p, c_int, c_int,
                                 ctypes.POINTER(ctypes.c_longlong), c_f32p, ctypes.c_char]),
    "nda_dcd_get_cell": (c_int, [ctypes.c_char_p, c_i32p, c_int, ctypes.POINTER(ctypes.c_longlong), c_f64p,
                                 ctypes.c_char]),
    "nda_dev_radial_nbr_counts": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                                          c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "nda_dev_within": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                               c_void_p, c_float, c_int, c_int, c_void_p]),
    "nda_dev_distances_sum": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                      c_int, c_void_p]),
    "nda_dev_radial_nbr_counts_sel": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                                              c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "nda_dev_distances_sum_sel": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                          c_int, c_int, c_int, c_void_p]),
    "nda_dev_synth_water_box": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float,
                                        ctypes.c_uint32, c_void_p]),
    "nda_dev_synth_water_box_slice": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float,
                                              ctypes.c_uint32, c_void_p]),
    "nda_measure_fp32_peak": (c_int, [c_int, c_f64p]),
    "nda_last_stats": (c_int, [c_f64p]),
    "nda_last_kernel_ms": (c_int, [c_f64p, c_i32p]),
}

_lib = None


def load():
    """Load the CUDA extension; raise RuntimeError (never fall back) if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "namdanalyzer_b200: CUDA extension %s is missing -- build it with "
                "`python -m namdanalyzer_b200.build` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def set_option(name, value=None):
    """nda_set_option: ``value`` None restores the default.  The library reads its NDA_* environment
    variables once, at first use; this is the only way to change a switch afterwards."""
    check(load().nda_set_option(name.encode(), None if value is None else str(value).encode()))


class options:
    """``with _capi.options(NDA_FILTER="fold2"): ...`` -- set switches for a block, restore the defaults after."""

    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        for k, v in self.kw.items():
            set_option(k, v)
        return self

    def __exit__(self, *exc):
        for k in self.kw:
            set_option(k, os.environ.get(k))      # back to what the process was started with
        return False


def set_devices(devices):
    """Shard the host entry points' frames over these devices (one process, one host thread per device)."""
    arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
    check(load().nda_set_devices(arr, len(devices)))


def check(rc):
    """CUDA / argument failures -> RuntimeError with the library's message (the reference's CUDA
    build calls exit() instead, lib/cuda/src/getWithin.cu:9-17)."""
    if rc != 0:
        msg = load().nda_last_error()
        raise RuntimeError("libnda_b200: " + (msg.decode() if msg else "error %d" % rc))
