"""ctypes binding of libtgcn_b200.so (the C-ABI declared in include/tgcn_b200.h).

There is deliberately no fallback: if the library is missing and cannot be built, or a call
fails, a RuntimeError is raised.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtgcn_b200.so")

OK = 0
ABI_VERSION = 200      # tgcn_version(): bumped whenever a C-ABI signature changes (include/tgcn_b200.h)
BIAS_NONE, BIAS_PER_VERTEX, BIAS_PER_FILTER = 0, 1, 2
RECURSION_REFERENCE, RECURSION_CHEBYSHEV = 0, 1
ENGINE_AUTO, ENGINE_FFMA, ENGINE_TCGEN05, ENGINE_RESIDENT = 0, 1, 2, 3

_p = ctypes.c_void_p
_i = ctypes.c_int
_l = ctypes.c_int64
_f = ctypes.c_float

# name -> (restype, argtypes); must list every symbol include/tgcn_b200.h declares
SIGNATURES = {
    "tgcn_version": (_i, []),
    "tgcn_last_error": (ctypes.c_char_p, []),
    "tgcn_device_supported": (_i, []),
    "tgcn_launch_count": (ctypes.c_longlong, []),
    "tgcn_set_tuning": (_i, [ctypes.c_char_p, _i]),
    "tgcn_rowtile_plan_host": (_l, [_p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "tgcn_rowtile_plan_create": (_l, [_p, _i, _i, _i, _p, _p, _p]),
    "tgcn_rowtile_plan_destroy": (_i, [_l]),
    "tgcn_to_slab": (_i, [_p, _p, _i, _i, _i, _p]),
    "tgcn_from_slab": (_i, [_p, _p, _i, _i, _i, _p]),
    "tgcn_spmm_step": (_i, [_p, _p, _p, _i, _p, _p, _p, _l, _f, _f, _p]),
    "tgcn_cheb_basis": (_i, [_p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _p]),
    "tgcn_basis_to_reference": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "tgcn_mix_weights": (_i, [_p, _p, _i, _l, _i, _i, _p]),
    "tgcn_contract_fwd_scratch": (_l, [_i, _i, _i, _i, _i]),
    "tgcn_contract_fwd": (_i, [_p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "tgcn_contract_bwd_w_workspace": (_l, [_i, _i, _i, _i, _i]),
    "tgcn_contract_bwd_w": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "tgcn_contract_bwd_x": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "tgcn_cheb_adjoint": (_i, [_p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _p]),
    "tgcn_bias_grad": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "tgcn_pool_max_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "tgcn_pool_max_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "tgcn_layer_fwd_workspace": (_l, [_i, _i, _i, _i, _i]),
    "tgcn_layer_slab_width": (_i, [_i, _i, _i, _i, _i, _i]),
    "tgcn_layer_fwd": (_i, [_p, _p, _p, _i, _p, _p, _p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "tgcn_layer_bwd_workspace": (_l, [_i, _i, _i, _i, _i]),
    "tgcn_layer_bwd": (_i, [_p, _p, _p, _i, _p, _p, _p, _p, _p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "tgcn_resident_supported": (_i, [_i, _i, _i, _i, _l]),
    "tgcn_resident_stack_bytes": (_l, [_i, _i, _i, _i]),
    "tgcn_resident_bwd_workspace": (_l, [_i, _i, _i, _i, _i]),
    "tgcn_pack_csr_host": (_l, [_p, _p, _p, _i, _i, _p, _p]),
    "tgcn_resident_pack_classes": (_i, [_i, _i, _i, _i]),
    "tgcn_resident_weights_bytes": (_l, [_i, _i, _i]),
    "tgcn_resident_layer_fwd": (_i, [_p, _p, _i, _l, _p, _p, _p, _i, _p, _p, _p, _i, _i, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "tgcn_resident_layer_bwd": (_i, [_p, _p, _i, _l, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _i, _p, _p,
                                     _i, _i, _i, _i, _i, _p]),
    "tgcn_head_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _f, _f, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "tgcn_head_workspace": (_l, [_i, _i, _i]),
    "tgcn_head_fused_update_supported": (_i, [_i, _i, _i]),
    "tgcn_head_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "tgcn_peer_alloc": (_i, [_l, _p]),
    "tgcn_peer_free": (_i, [_p]),
    "tgcn_peer_export": (_i, [_p, _p]),
    "tgcn_peer_import": (_i, [_p, _p]),
    "tgcn_peer_close": (_i, [_p]),
    "tgcn_peer_region_bytes": (_l, [_l, _i]),
    "tgcn_peer_allreduce_sgd": (_i, [_p, _i, _i, _p, _p, _p, _p, _i, _f, _f, _p, _p]),
    "tgcn_halo_signal": (_i, [_p, _p, _i, _i, _p, _p]),
    "tgcn_halo_pull": (_i, [_p, _p, _p, _i, _i, _p, _p, _i, _l, _p, _p, _p]),
    "tgcn_pair_one_level_f32": (_i, [_p, _p, _p, _l, _p, _p, _l, _p]),
    "tgcn_pair_one_level_f64": (_i, [_p, _p, _p, _l, _p, _p, _l, _p]),
}



class Dropout(ctypes.Structure):
    """tgcn_dropout_t (include/tgcn_b200.h)."""
    _fields_ = [("p", ctypes.c_float), ("seed", ctypes.c_uint32), ("step", ctypes.c_void_p)]


class Fc1Update(ctypes.Structure):
    """tgcn_fc1_update_t (include/tgcn_b200.h)."""
    _fields_ = [("lr", ctypes.c_float), ("momentum", ctypes.c_float), ("mom", ctypes.c_void_p), ("world", ctypes.c_int),
                ("rank", ctypes.c_int), ("regions", ctypes.c_void_p), ("state", ctypes.c_void_p), ("gather", ctypes.c_void_p)]


def dropout_arg(p, seed=0, step_ptr=None):
    """byref(tgcn_dropout_t) for the C calls, or None when p == 0 (the struct must outlive the call only)."""
    if not p:
        return None
    return ctypes.byref(Dropout(float(p), int(seed) & 0xFFFFFFFF, step_ptr))


_lib = None
_lock = threading.Lock()


def load():
    """Load (building first if the sources changed and nvcc is present) and return the CDLL."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        if os.path.exists(_build.nvcc_path()) or not os.path.exists(LIB_PATH):
            # the fingerprint check inside build() is cheap: a stale, git-ignored .so must never be bound to newer
            # ctypes signatures.  Without nvcc a prebuilt library is used as it is (and must pass the version check).
            _build.build()
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libtgcn_b200.so is missing and could not be built; there is no fallback path")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.tgcn_version() != ABI_VERSION:
            raise RuntimeError("libtgcn_b200.so reports ABI version %d, the Python side expects %d: rebuild it "
                               "(python -m tgcn_b200.build --force)" % (lib.tgcn_version(), ABI_VERSION))
        _lib = lib
    return _lib


def last_error():
    return load().tgcn_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != OK:
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, last_error()))
