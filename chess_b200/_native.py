"""ctypes binding of libchess_b200.so (the C ABI in include/chess_b200.h).

There is no CPU fallback: importing this module without the shared library
raises, and creating a context without a B200 raises.  The library is built
in-tree by `python -m chess_b200.build` (nvcc, sm_100a).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libchess_b200.so")


class ChessNativeError(RuntimeError):
    pass


def _load():
    # (Re)build when the sources changed since the library was made; the stamp
    # check is a hash of the few .cu files.  Never fall back to CPU code.
    try:
        from . import build as _build
        _build.build()
    except Exception as exc:  # pragma: no cover - depends on the toolchain
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "chess_b200: native library {} is missing and could not be built ({}). "
                "Run `python -m chess_b200.build`; there is no CPU fallback.".format(LIB_PATH, exc))
        # an older library is there.  Using it is only right when it was built from these very sources
        # (a box without nvcc running the shipped .so); otherwise results would come from code that is not
        # the checked-in source, so refuse unless the caller explicitly accepts the stale library.
        stale = True
        try:
            with open(LIB_PATH + ".stamp") as fh:
                stale = fh.read().strip() != _build._stamp()
        except Exception:
            pass
        if stale:
            if os.environ.get("CHESS_B200_ALLOW_STALE_LIB") != "1":
                raise ImportError(
                    "chess_b200: the sources changed since {} was built and the rebuild failed ({}). Fix the "
                    "build, or set CHESS_B200_ALLOW_STALE_LIB=1 to load the stale library.".format(LIB_PATH, exc))
            import warnings
            warnings.warn("chess_b200: loading a STALE {} (rebuild failed: {})".format(LIB_PATH, exc))
    return C.CDLL(LIB_PATH)


lib = _load()


class chess_pair(C.Structure):
    _fields_ = [("ref_off", C.c_int64), ("qry_off", C.c_int64),
                ("ref_pitch", C.c_int32), ("qry_pitch", C.c_int32),
                ("ref_n", C.c_int32), ("qry_n", C.c_int32),
                ("flip", C.c_int32), ("reserved", C.c_int32)]


class chess_sample(C.Structure):
    _fields_ = [("band", C.c_void_p), ("nd_lo", C.c_void_p), ("nd_hi", C.c_void_p),
                ("pmax", C.c_void_p), ("pmin", C.c_void_p),
                ("n_bins", C.c_int64), ("W", C.c_int32), ("reserved", C.c_int32), ("rowpre", C.c_void_p)]


class chess_params(C.Structure):
    _fields_ = [("min_bins", C.c_int32), ("keep_unmappable", C.c_int32),
                ("abs_window", C.c_int32), ("compute_sn", C.c_int32),
                ("rel_window", C.c_double), ("mappability_cutoff", C.c_double),
                ("default_value", C.c_double), ("kernel", C.c_int32), ("flags", C.c_int32),
                ("data_range", C.c_double)]


PAIR_DTYPE = np.dtype([("ref_off", "<i8"), ("qry_off", "<i8"), ("ref_pitch", "<i4"),
                       ("qry_pitch", "<i4"), ("ref_n", "<i4"), ("qry_n", "<i4"),
                       ("flip", "<i4"), ("reserved", "<i4")], align=True)
assert PAIR_DTYPE.itemsize == C.sizeof(chess_pair) == 40

PAIR_OK, PAIR_NOT_FOUND, PAIR_TOO_SMALL, PAIR_UNMAPPABLE, PAIR_WINDOW_EXCEEDS = range(5)
FLAG_SYMMETRIC, FLAG_EQUAL_SIZES, FLAG_MOSTLY_EQUAL_SIZES = 1, 2, 4

_vp, _i32, _i64, _dbl, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_size_t
_PP = C.POINTER(chess_params)
_SP = C.POINTER(chess_sample)

# name -> (restype, argtypes); every symbol include/chess_b200.h declares
SIGNATURES = {
    "chess_abi_version": (C.c_int, []),
    "chess_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "chess_ctx_destroy": (None, [_vp]),
    "chess_last_error": (C.c_char_p, [_vp]),
    "chess_launch_count": (_i64, [_vp]),
    "chess_expected_workspace_bytes": (_sz, [_i64]),
    "chess_expected_from_coo": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _i64,
                                          _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "chess_possible_contacts": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _vp, _vp]),
    "chess_oe_values": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "chess_band_build": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _dbl, _vp, _vp]),
    "chess_band_build_range": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _dbl, _vp, _vp]),
    "chess_coo_check": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "chess_symmetric_check": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _vp]),
    "chess_gather_submatrix": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "chess_compare_batch": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _PP, _vp, _vp, _vp, _vp]),
    "chess_compare_batch_host": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _PP, _vp, _vp, _vp, _vp]),
    "chess_band_index_build": (C.c_int, [_vp, _vp, _i64, _i32, _dbl, _vp, _vp, _vp]),
    "chess_band_extrema_build": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "chess_compare_band_batch": (C.c_int, [_vp, _SP, _SP, _vp, _i64, _i32, _PP, _vp, _vp, _vp, _vp]),
    "chess_compare_band_batch_host": (C.c_int, [_vp, _SP, _SP, _vp, _i64, _i32, _PP, _vp, _vp, _vp, _vp]),
    "chess_sim_band_batch_host": (C.c_int, [_vp, _SP, _SP, _vp, _i64, _i32, _PP, _vp, _vp, _vp, _vp, _vp]),
    "chess_sn_dense": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "chess_zscore": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "chess_band_prefix_build": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "chess_sparse_text_lines": (C.c_int, [_vp, _vp, _i64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _vp]),
    "chess_sparse_text_parse": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                                          C.POINTER(C.c_int64), _vp]),
    "chess_sparse_text_compact": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                                            C.POINTER(C.c_int64), _vp]),
    "chess_log_ratio_batch": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "chess_background_regions": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "chess_background_batch": (C.c_int, [_vp, _SP, _SP, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _PP,
                                         _vp, _vp, _vp, _vp]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args

# the struct layouts above follow include/chess_b200.h at this ABI version (chess_sample grew `rowpre` in 2,
# chess_params grew `data_range` in 3; 4 added chess_band_build_range and chess_coo_check)
ABI_VERSION = 4
if lib.chess_abi_version() != ABI_VERSION:
    raise ChessNativeError("libchess_b200.so has ABI version {}, this binding expects {}: rebuild with "
                           "python -m chess_b200.build".format(lib.chess_abi_version(), ABI_VERSION))


def last_error(ctx=None):
    msg = lib.chess_last_error(ctx)
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, ctx=None, what=""):
    if rc != 0:
        raise ChessNativeError("{} failed (code {}): {}".format(what or "chess_b200 call", rc, last_error(ctx)))


class Context(object):
    """One library context = one GPU.  Not thread-safe."""

    def __init__(self, device=0):
        handle = _vp()
        rc = lib.chess_ctx_create(int(device), C.byref(handle))
        if rc != 0:
            raise ChessNativeError("chess_ctx_create(device={}) failed (code {}): {}".format(
                device, rc, last_error(None)))
        self.handle = handle
        self.device = int(device)

    def launches(self):
        return int(lib.chess_launch_count(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            lib.chess_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_params(min_bins=20, keep_unmappable_bins=False, absolute_window_size=None,
                relative_window_size=1.0, mappability_cutoff=0.1, default_value=1,
                compute_sn=True, kernel=0, flags=0, data_range=None):
    p = chess_params()
    p.data_range = float(data_range) if data_range else 0.0
    p.min_bins = int(min_bins)
    p.keep_unmappable = 1 if keep_unmappable_bins else 0
    p.abs_window = int(absolute_window_size) if absolute_window_size else 0
    p.compute_sn = 1 if compute_sn else 0
    p.rel_window = float(relative_window_size)
    p.mappability_cutoff = float(mappability_cutoff)
    p.default_value = float(default_value)
    p.kernel = int(kernel)
    p.flags = int(flags)
    return p
