"""ctypes binding of libnpde_b200.so -- the C ABI declared in include/npde.h.

There is exactly one implementation of every operator: the CUDA library.  If it
has not been built, importing any operator raises; there is no CPU or PyTorch
fallback and no hook to install one.  (The CPU test-suite binds the
host-compiled SIMT emulator build of the same sources by swapping the module
state from tests/emu_backend.py -- the package itself never does.)
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# NPDE_B200_LIB: development hook to A/B-test differently compiled builds of the SAME CUDA library
LIB_PATH = os.environ.get("NPDE_B200_LIB") or os.path.join(HERE, "libnpde_b200.so")

OK = 0
BC_SQUARE, BC_MASK = 0, 1
ACT = {"none": 0, "clamp": 1, "sigmoid": 2}
AGG = {"max": 0, "mean": 1}
INIT = {"zero": 0, "random": 1, "avg": 2}
IT_JACOBI, IT_CONV, IT_MULTIGRID, IT_UNET = 0, 1, 2, 3
WS_REDUCE, WS_CONV_FWD, WS_CONV_BWD, WS_MG, WS_UNET_FWD, WS_UNET_BWD, WS_CG, WS_CONV_BPTT, WS_ITERATE = 0, 1, 2, 3, 4, 5, 6, 7, 16
MAX_CONV_LAYERS = 8

_c_float_p = ctypes.c_void_p  # raw device pointers
_int = ctypes.c_int
_size = ctypes.c_size_t
_vp = ctypes.c_void_p


class IterateDesc(ctypes.Structure):
    """struct npde_iterate_desc (include/npde.h)."""
    _fields_ = [
        ("iterator", _int), ("B", _int), ("N", _int), ("bc_kind", _int), ("n_layers", _int),
        ("pre", _int), ("post", _int), ("act", _int), ("k", _int), ("temporal_depth", _int),
        ("x_in", _vp), ("x_out", _vp), ("bc", _vp), ("f", _vp), ("w", _vp), ("w_host", _vp),
        ("gt", _vp), ("err_l2", _vp), ("err_fd", _vp), ("workspace", _vp), ("workspace_bytes", _size),
    ]


# name -> (restype, argtypes); must list every symbol of include/npde.h
PROTOTYPES = {
    "npde_version": (_int, []),
    "npde_debug_launch_count": (ctypes.c_ulonglong, []),
    "npde_strerror": (ctypes.c_char_p, [_int]),
    "npde_workspace_bytes": (_size, [_int] * 8),
    "npde_set_boundary": (_int, [_vp, _vp, _int, _int, _int, _vp]),
    "npde_initialize": (_int, [_vp, _vp, _int, _int, _vp, _int, _int, _vp]),
    "npde_pad_boundary": (_int, [_vp, _vp, _int, _vp, _int, _int, _vp]),
    "npde_fd_step": (_int, [_vp, _vp, _int, _vp, _vp, _int, _int, _vp]),
    "npde_fd_error": (_int, [_vp, _vp, _int, _vp, _int, _vp, _int, _int, _vp, _size, _vp]),
    "npde_l2_error": (_int, [_vp, _vp, _vp, _int, _int, _vp, _size, _vp]),
    "npde_cg_step": (_int, [_vp, _int, _vp, _int, _int, _vp, _size, _vp]),
    "npde_conv_H": (_int, [_vp, _vp, _int, _vp, _int, _vp, _int, _int, _vp, _size, _vp]),
    "npde_conv_step_fwd": (_int, [_vp, _vp, _int, _vp, _vp, _vp, _int, _int, _vp, _int, _int, _vp, _size, _vp]),
    "npde_conv_step_bwd": (_int, [_vp, _vp, _int, _vp, _vp, _int, _int, _vp, _vp, _vp, _int, _int, _vp, _size, _vp]),
    "npde_conv_iterate_tape": (_int, [_vp, _vp, _int, _vp, _vp, _vp, _int, _int, _int, _vp, _vp, _int, _int, _vp,
                                      _size, _vp]),
    "npde_conv_iterate_bwd": (_int, [_vp, _vp, _int, _vp, _vp, _int, _int, _int, _vp, _vp, _vp, _int, _int, _vp,
                                     _size, _vp]),
    "npde_subsample": (_int, [_vp, _vp, _int, _int, _vp]),
    "npde_restrict": (_int, [_vp, _vp, _int, _vp, _int, _int, _vp]),
    "npde_prolong": (_int, [_vp, _vp, _int, _vp, _int, _int, _vp]),
    "npde_mg_step": (_int, [_vp, _vp, _int, _vp, _int, _int, _int, _vp, _int, _int, _vp, _size, _vp]),
    "npde_unet_H": (_int, [_vp, _vp, _int, _vp, _vp, _vp, _int, _int, _int, _vp, _int, _int, _vp, _size, _vp]),
    "npde_unet_step_fwd": (_int, [_vp, _vp, _int, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp, _int, _int,
                                  _vp, _size, _vp]),
    "npde_unet_step_fwd_hw": (_int, [_vp, _vp, _int, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp, _int,
                                     _int, _vp, _size, _vp]),
    "npde_unet_step_bwd": (_int, [_vp, _vp, _int, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp, _vp, _vp, _vp,
                                  _vp, _int, _int, _vp, _size, _vp]),
    "npde_iterate": (_int, [ctypes.POINTER(IterateDesc), _vp]),
    "npde_solve": (_int, [ctypes.POINTER(IterateDesc), ctypes.c_float, _int, _int, _vp, _vp]),
    "npde_packed_floats": (_size, [_int, _int]),
    "npde_pack": (_int, [_vp, _vp, _int, _int, _vp]),
    "npde_unpack": (_int, [_vp, _vp, _int, _int, _vp]),
    "npde_conv_step_packed": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp]),
    "npde_conv_iterate_packed": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _vp]),
    "npde_fd_error_packed": (_int, [_vp, _vp, _vp, _int, _int, _vp]),
}

_lib = None
_emulated = False


class NpdeError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib().npde_strerror(code)
        super().__init__("%s failed: %s (code %d)" % (where, msg.decode() if msg else "?", code))


def _bind(path):
    handle = ctypes.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(handle, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return handle


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libnpde_b200.so is not built (%s). Build it with `python neural-pde-solver_b200/build.py` "
                "(needs nvcc); there is no CPU fallback." % LIB_PATH)
        _lib = _bind(LIB_PATH)
    return _lib


def host_pointers():
    """True iff the bound library takes HOST pointers -- only the emulator build that tests/emu_backend.py binds;
    the product (libnpde_b200.so) never does, so every wrapper insists on CUDA tensors."""
    return _emulated


def check(code, where):
    if code != OK:
        raise NpdeError(code, where)
