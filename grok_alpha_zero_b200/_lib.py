"""ctypes binding of libgaz_b200.so (the C ABI in include/gaz_b200.h).

The library is the product: there is no Python or CPU fallback.  If the shared object is
missing, or no CUDA device is present when an engine is created, the call fails loudly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libgaz_b200.so"


class GazConfig(C.Structure):
    _fields_ = [("game", C.c_int32), ("mode", C.c_int32), ("n_games", C.c_int32), ("trees_per_game", C.c_int32),
                ("node_cap", C.c_int32), ("slot_cap", C.c_int32), ("device", C.c_int32), ("lut_n", C.c_int32),
                ("c_puct_init", C.c_float), ("c_puct_base", C.c_float), ("gumbel_m", C.c_int32),
                ("use_softmax", C.c_int32), ("c_visit", C.c_double), ("c_scale", C.c_double), ("slot_pool", C.c_int64)]


# every symbol include/gaz_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "gaz_last_error": (C.c_char_p, []),
    "gaz_abi_version": (C.c_int, []),
    "gaz_create": (C.c_int, [C.POINTER(GazConfig), C.POINTER(_P)]),
    "gaz_destroy": (None, [_P]),
    "gaz_set_puct_params": (C.c_int, [_P, C.c_float, C.c_float]),
    "gaz_set_gumbel_params": (C.c_int, [_P, C.c_int, C.c_double, C.c_double, C.c_int]),
    "gaz_set_game": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, C.c_int]),
    "gaz_set_games": (C.c_int, [_P, _P, _P]),
    "gaz_reset_games": (C.c_int, [_P]),
    "gaz_apply_actions": (C.c_int, [_P, _P, _P]),
    "gaz_get_game": (C.c_int, [_P, C.c_int, _P, _P]),
    "gaz_get_states": (C.c_int, [_P, _P, _P]),
    "gaz_set_noise": (C.c_int, [_P, C.c_float, C.c_float, C.c_uint64]),
    "gaz_set_tree_keys": (C.c_int, [_P, _P]),
    "gaz_gumbel_pi_dense": (C.c_int, [_P, _P]),
    "gaz_new_roots": (C.c_int, [_P, _P]),
    "gaz_run_begin": (C.c_int, [_P, _P]),
    "gaz_select": (C.c_int, [_P]),
    "gaz_get_leaves": (C.c_int, [_P, _P, _P]),
    "gaz_get_leaf_depths": (C.c_int, [_P, _P]),
    "gaz_put_evals": (C.c_int, [_P, _P, _P, C.c_int]),
    "gaz_eval_hash": (C.c_int, [_P, C.c_uint64, C.c_int]),
    "gaz_expand": (C.c_int, [_P]),
    "gaz_remaining": (C.c_int, [_P]),
    "gaz_rounds_hash": (C.c_int, [_P, C.c_int, C.c_uint64, C.c_int]),
    "gaz_prune": (C.c_int, [_P, _P, C.c_int]),
    "gaz_root_stats": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gaz_gumbel_pi": (C.c_int, [_P, C.c_int, _P]),
    "gaz_set_gumbel_noise": (C.c_int, [_P, _P]),
    "gaz_root_dense": (C.c_int, [_P, _P, _P, _P]),
    "gaz_timer_begin": (C.c_int, [_P]),
    "gaz_timer_end": (C.c_int, [_P, _P]),
    "gaz_sync": (C.c_int, [_P]),
    "gaz_status": (C.c_int, [_P]),
    "gaz_tree_sizes": (C.c_int, [_P, _P]),
    "gaz_bytes_allocated": (C.c_int64, [_P]),
    "gaz_pool_info": (C.c_int, [_P, _P]),
    "gaz_eval_cache_enable": (C.c_int, [_P, C.c_int64, C.c_int]),
    "gaz_eval_cache_stats": (C.c_int, [_P, _P]),
    "gaz_augment": (C.c_int, [C.c_int, _P, _P, C.c_int64, C.c_int, C.c_int, _P, _P, C.c_int, _P, _P]),
}


def bind(path, symbols=None):
    """dlopen `path` and attach prototypes; raises if a declared symbol is missing."""
    lib = C.CDLL(path)
    for name, (res, args) in (symbols or SYMBOLS).items():
        fn = getattr(lib, name)  # AttributeError -> loud failure
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib_path():
    return os.path.join(HERE, LIB_NAME)


def load():
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise RuntimeError(
                "%s is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % p)
        syms = dict(SYMBOLS)
        try:
            from . import _net_symbols
            syms.update(_net_symbols.SYMBOLS)
        except ImportError:
            pass
        _lib = bind(p, syms)
    return _lib
