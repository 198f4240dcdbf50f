"""ctypes binding of libb200math.so (the C ABI declared in include/b200.h).

The library is the product; this module only loads it and declares the prototypes.  There is no
CPU fallback: if the shared object is missing, or no CUDA device is visible, the first call
raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB", os.path.join(_HERE, "libb200math.so"))     # B200_LIB: development override

# flags (include/b200.h)
FEXP = 0x1
IN_MONT = 0x2
OUT_MONT = 0x4
OUT_UNITY_ONLY = 0x8
DEVICE_PTRS = 0x10
BASES_TABLES = 0x20
NO_SUBGROUP_CHECK = 0x40

ERR_CUDA, ERR_ARG, ERR_ENCODING, ERR_NOGPU = -1, -2, -3, -4

# every symbol include/b200.h declares: name -> (restype, argtypes)
_vp, _sz, _u32, _int, _u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint64
PROTOTYPES = {
    "b200_init": (_int, [_u32]),
    "b200_shutdown": (None, []),
    "b200_last_error": (ctypes.c_char_p, []),
    "b200_device_count": (_int, []),
    "b200_set_device": (_int, [_int]),
    "b200_set_stream": (_int, [_vp]),
    "b200_take_error": (_int, [ctypes.POINTER(_int)]),
    "b200_fp_bytes": (_int, [_int]),
    "b200_pairing_batch": (_int, [_int, _sz, _vp, _vp, _vp, _u32]),
    "b200_pairing2_batch": (_int, [_int, _sz, _vp, _vp, _vp, _vp, _vp, _u32]),
    "b200_fexp_batch": (_int, [_int, _sz, _vp, _vp, _u32]),
    "b200_g1_mul_batch": (_int, [_int, _sz, _vp, _vp, _vp, _u32]),
    "b200_g1_mul2_batch": (_int, [_int, _sz, _vp, _vp, _vp, _vp, _vp, _u32]),
    "b200_g1_msm": (_int, [_int, _sz, _vp, _vp, _vp, _u32]),
    "b200_bases_upload": (_int, [_int, _sz, _vp, _u32, ctypes.POINTER(_u64)]),
    "b200_g1_msm_resident": (_int, [_u64, _sz, _vp, _vp, _u32]),
    "b200_bases_free": (_int, [_u64]),
    "b200_g1_sum": (_int, [_int, _sz, _vp, _vp, _u32]),
    "b200_g2_mul_batch": (_int, [_int, _sz, _vp, _vp, _vp, _u32]),
    "b200_g2_sum": (_int, [_int, _sz, _vp, _vp, _u32]),
    "b200_gt_mul_batch": (_int, [_int, _sz, _vp, _vp, _vp, _u32]),
    "b200_gt_inv_batch": (_int, [_int, _sz, _vp, _vp, _u32]),
    "b200_gt_exp_batch": (_int, [_int, _sz, _vp, _vp, _vp, _u32]),
    "b200_g1_decompress_batch": (_int, [_int, _sz, _vp, _vp, _u32]),
    "b200_g2_decompress_batch": (_int, [_int, _sz, _vp, _vp, _u32]),
    "b200_g1_compress_batch": (_int, [_int, _sz, _vp, _vp, _u32]),
    "b200_g2_compress_batch": (_int, [_int, _sz, _vp, _vp, _u32]),
    "b200_g1_validate_batch": (_int, [_int, _sz, _vp, _vp, _u32]),
    "b200_g2_validate_batch": (_int, [_int, _sz, _vp, _vp, _u32]),
    "b200_g2_lines_upload": (_int, [_int, _sz, _vp, _u32, ctypes.POINTER(_u64)]),
    "b200_g2_lines_free": (_int, [_u64]),
    "b200_pairing_fixed_batch": (_int, [_u64, _sz, _vp, _vp, _vp, _u32]),
    "b200_pairing2_fixed_batch": (_int, [_u64, _sz, _vp, _vp, _vp, _vp, _vp, _u32]),
    "b200_g1_normalize_batch": (_int, [_int, _sz, _vp, _vp, _u32]),
    "b200_g2_msm": (_int, [_int, _sz, _vp, _vp, _vp, _u32]),
    "b200_hash_to_g1_batch": (_int, [_int, _sz, _vp, _vp, _vp, _sz, _vp, _u32]),
    "b200_launch_count": (_u64, []),
}


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libb200math error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Load libb200math.so and bind every prototype.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "%s not found: build it with `make lib` (or __graft_entry__.build()). "
            "mathlib_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise B200Error(rc, load().b200_last_error().decode("utf-8", "replace"))


def buf_ptr(b):
    """void* of a bytes / bytearray / numpy array / int (device pointer)."""
    if b is None:
        return None
    if isinstance(b, int):
        return ctypes.c_void_p(b)
    if isinstance(b, (bytes, bytearray)):
        return ctypes.cast((ctypes.c_char * len(b)).from_buffer(b) if isinstance(b, bytearray)
                           else ctypes.c_char_p(b), ctypes.c_void_p)
    if hasattr(b, "ctypes"):             # numpy
        return ctypes.c_void_p(b.ctypes.data)
    raise TypeError("unsupported buffer type %r" % type(b))
