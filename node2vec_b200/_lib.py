"""ctypes binding of libn2v_b200.so (the C ABI in include/n2v_b200.h).

The library is the product: if it is missing or a call fails this module raises -- there
is no fallback path of any kind.
"""
import ctypes as C
import os

from . import build as _build

N2V_ABI_VERSION = 7          # include/n2v_b200.h
N2V_MAX_PARTS = 16
OK, ERR_INVALID, ERR_CUDA, ERR_SCRATCH, ERR_ZERO_WEIGHT = 0, 1, 2, 3, 4
SUM_MODE = {"naive": 0, "neumaier": 1}
GRAPH_UNIT_WEIGHT, GRAPH_SYMMETRIC, GRAPH_SIMPLE = 1, 2, 4


class GraphPart(C.Structure):
    _fields_ = [("vtx", C.c_void_p), ("arcs", C.c_void_p), ("col", C.c_void_p), ("weight", C.c_void_p),
                ("hash", C.c_void_p), ("ratio", C.c_void_p)]


class Graph(C.Structure):
    _fields_ = [("n_vertices", C.c_int64), ("n_arcs", C.c_int64), ("flags", C.c_uint32),
                ("n_parts", C.c_int32), ("part_size", C.c_int64), ("parts", GraphPart * N2V_MAX_PARTS)]


class WalkConsts(C.Structure):
    _fields_ = [("t_ret", C.c_uint64), ("t_nbr", C.c_uint64), ("t_far", C.c_uint64),
                ("fold_gain", C.c_float), ("fold_mode", C.c_int32), ("max_trials", C.c_int32),
                ("mix_qm1", C.c_float)]


class SgnsParams(C.Structure):
    _fields_ = [("dim", C.c_int32), ("window", C.c_int32), ("negative", C.c_int32), ("epochs", C.c_int32),
                ("epoch", C.c_int32), ("batch_words", C.c_int32), ("atomic_updates", C.c_int32),
                ("reserved", C.c_int32), ("alpha", C.c_float), ("min_alpha", C.c_float), ("seed", C.c_uint64),
                ("walk_offset", C.c_int64), ("total_walks", C.c_int64)]


SGNS_STAT_NAMES = ("pairs", "tokens_kept", "negatives_skipped", "targets_clipped")
WALK_STAT_NAMES = ("steps", "trials", "probes", "searches", "fold_hits", "fallbacks", "dead", "reserved")

_P = C.c_void_p
_SIGNATURES = {
    "n2v_abi_version": (C.c_int, []),
    "n2v_last_error": (C.c_char_p, []),
    "n2v_csr_scratch_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "n2v_csr_build": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, C.c_size_t,
                                C.POINTER(C.c_uint32), _P]),
    "n2v_ipc_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "n2v_ipc_free": (C.c_int, [_P]),
    "n2v_ipc_export": (C.c_int, [_P, C.c_char_p]),
    "n2v_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "n2v_ipc_close": (C.c_int, [_P]),
    "n2v_hash_buckets_bound": (C.c_int64, [C.c_int64, C.c_int64]),
    "n2v_hash_build": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, C.c_int64, _P]),
    "n2v_alias_build": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int, _P, _P, _P, _P,
                                  C.POINTER(C.c_int64), _P]),
    "n2v_edge_alias_build": (C.c_int, [C.POINTER(Graph), _P, _P, C.c_int64, C.c_double, C.c_double, C.c_int,
                                       _P, _P, _P, _P, _P]),
    "n2v_alias_draw": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, _P, _P]),
    "n2v_walk": (C.c_int, [C.POINTER(Graph), _P, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_double,
                           C.c_uint64, _P, C.c_int64, _P, _P, _P]),
    "n2v_walk_consts": (C.c_int, [C.c_double, C.c_double, C.c_uint32, C.c_int, C.POINTER(WalkConsts)]),
    "n2v_ratio_build": (C.c_int, [C.POINTER(Graph), _P, _P]),
    "n2v_trim_sample": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_uint32, _P, _P, _P]),
    "n2v_set_l2_fetch_granularity": (C.c_int, [C.c_int]),
    "n2v_get_l2_fetch_granularity": (C.c_int, []),
    "n2v_first_occurrence_slots": (C.c_int64, [C.c_int64]),
    "n2v_first_occurrence": (C.c_int, [_P, _P, _P, C.c_int64, _P, C.c_int64, _P, _P]),
    "n2v_first_occurrence_bytes": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int64, _P, _P]),
    "n2v_vocab_count": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P]),
    "n2v_sgns_prepare": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_double, C.c_double, _P, _P, _P,
                                   C.POINTER(C.c_int64), _P]),
    "n2v_sgns_init": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_uint64, _P]),
    "n2v_sgns_exp_table": (C.c_int, [C.POINTER(C.c_float)]),
    "n2v_sgns_train": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int64, _P, _P, C.c_int64, _P, _P, _P,
                                 C.POINTER(SgnsParams), _P, _P, _P, C.c_int64, _P]),
    "n2v_sgns_train_shared": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int64, _P, _P, C.c_int64, _P, _P, _P,
                                        C.POINTER(SgnsParams), _P, _P, _P, C.c_int64, _P]),
    "n2v_scale": (C.c_int, [_P, C.c_int64, C.c_float, _P]),
}

_lib = None


class N2VError(RuntimeError):
    pass


def library_path() -> str:
    return _build.SO


def load(build_if_missing: bool = True):
    """Load libn2v_b200.so, building it with nvcc when absent.  Raises if it cannot."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("N2V_B200_LIB", _build.SO)   # override = kernel-tuning builds only
    if path == _build.SO and build_if_missing and _build.is_stale() and _build.have_nvcc():
        _build.build_library()          # sources newer than the .so (or no .so): rebuild in-tree
    if not os.path.exists(path):
        raise N2VError(f"{path} is missing: run `python -m node2vec_b200.build`")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = ABI mismatch, loudly
        fn.restype = res
        fn.argtypes = args
    if lib.n2v_abi_version() != N2V_ABI_VERSION:
        raise N2VError(f"{path} speaks ABI {lib.n2v_abi_version()}, this package binds ABI {N2V_ABI_VERSION}: rebuild it")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    """Map a status code to the exception the reference's callers expect."""
    if rc == OK:
        return
    msg = load().n2v_last_error().decode("utf-8", "replace")
    if rc in (ERR_INVALID, ERR_ZERO_WEIGHT):
        raise ValueError(msg)
    raise N2VError(f"{what}: {msg} (code {rc})")


def current_stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_tuned = set()


def tune_device(device_index=None):
    """Optional per-device tuning knob: N2V_L2_FETCH=32|64|128 sets cudaLimitMaxL2FetchGranularity
    before the first walk.  Measured on B200 (profiles/r02_gather_granule_l2fetch.txt): the random
    32-byte gather rate is 42.6 G sectors/s at 32, 64 (driver default) and 128 alike, so the default
    is to leave the driver's value alone."""
    import torch
    idx = torch.cuda.current_device() if device_index is None else int(device_index)
    if idx in _tuned:
        return
    _tuned.add(idx)
    want = int(os.environ.get("N2V_L2_FETCH", "0"))
    if want:
        with torch.cuda.device(idx):
            check(load().n2v_set_l2_fetch_granularity(want), "n2v_set_l2_fetch_granularity")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise N2VError("node2vec_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def ptr(t):
    """Device (or host) address of a torch tensor / None."""
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())
