"""ctypes binding to cuttlefish_b200/lib/libcfx.so (include/cfx.h). No CPU fallback exists:
if the library is missing and cannot be built, importing raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("CFX_BUILD_DIR", "lib"), "libcfx.so")      # (developer: a variant build)


class SurfaceDesc(ctypes.Structure):
    """cfx_surface_desc (include/cfx.h)."""
    _fields_ = [("format", ctypes.c_uint32), ("type", ctypes.c_uint32), ("quality", ctypes.c_uint32),
                ("alpha_type", ctypes.c_uint32), ("color_mask", ctypes.c_uint32),
                ("color_space", ctypes.c_uint32), ("width", ctypes.c_uint32), ("height", ctypes.c_uint32),
                ("src_format", ctypes.c_uint32), ("flags", ctypes.c_uint32),
                ("src_row_pitch", ctypes.c_uint64)]


# every symbol include/cfx.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("cfx_init", ctypes.c_int, [ctypes.c_int]),
    ("cfx_init_devices", ctypes.c_int, [ctypes.c_int]),
    ("cfx_set_devices", ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    ("cfx_device_count", ctypes.c_int, []),
    ("cfx_shutdown", None, []),
    ("cfx_format_supported", ctypes.c_int, [ctypes.c_uint32, ctypes.c_uint32]),
    ("cfx_format_is_exact", ctypes.c_int, [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]),
    ("cfx_block_info", ctypes.c_int, [ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32),
                                      ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]),
    ("cfx_encoded_size", ctypes.c_size_t, [ctypes.POINTER(SurfaceDesc)]),
    ("cfx_encode", ctypes.c_int, [ctypes.POINTER(SurfaceDesc), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    ("cfx_encode_batch", ctypes.c_int, [ctypes.c_int, ctypes.POINTER(SurfaceDesc), ctypes.POINTER(ctypes.c_void_p),
                                        ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]),
    ("cfx_encode_device", ctypes.c_int, [ctypes.POINTER(SurfaceDesc), ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_size_t, ctypes.c_void_p]),
    ("cfx_resize", ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_void_p,
                                  ctypes.c_uint32, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint32]),
    ("cfx_mip_levels", ctypes.c_uint32, [ctypes.c_uint32, ctypes.c_uint32]),
    ("cfx_encode_mip_chain", ctypes.c_int, [ctypes.POINTER(SurfaceDesc), ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
                                            ctypes.POINTER(ctypes.c_void_p)]),
    ("cfx_encode_mip_chain_device", ctypes.c_int, [ctypes.POINTER(SurfaceDesc), ctypes.c_void_p, ctypes.c_uint32,
                                                   ctypes.c_uint32, ctypes.POINTER(ctypes.c_void_p),
                                                   ctypes.POINTER(ctypes.c_size_t), ctypes.c_void_p]),
    ("cfx_dds_header", ctypes.c_size_t, [ctypes.POINTER(SurfaceDesc), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]),
    ("cfx_ktx_header", ctypes.c_size_t, [ctypes.POINTER(SurfaceDesc), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]),
    ("cfx_encode_mip_chain_to_file", ctypes.c_int, [ctypes.POINTER(SurfaceDesc), ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                                    ctypes.c_uint32, ctypes.c_char_p]),
    ("cfx_ipc_export", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    ("cfx_ipc_open", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    ("cfx_ipc_close", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    ("cfx_host_alloc", ctypes.c_void_p, [ctypes.c_size_t]),
    ("cfx_host_free", None, [ctypes.c_void_p]),
    ("cfx_kernel_launches", ctypes.c_uint64, []),
    ("cfx_last_error", ctypes.c_char_p, []),
    ("cfx_version", ctypes.c_char_p, []),
]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
