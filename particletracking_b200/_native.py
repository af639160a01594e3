"""ctypes binding of ``libptk.so`` (C ABI declared in ``include/ptk.h``).

There is no CPU fallback: importing this module without the built library, or calling
it without a B200-class device, raises.  Build with ``python -c "import __graft_entry__ as
g; g.build()"`` or ``make -C particletracking_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os

from .rig import PtkRig

_HERE = os.path.dirname(os.path.abspath(__file__))
# PTK_LIB selects another build of the same ABI (the bounds-checked libptk_dbg.so); default: the release library
LIB_PATH = os.environ.get("PTK_LIB") or os.path.join(_HERE, "libptk.so")

PTK_OK = 0
PTK_ERR_NUMERIC = -5
PROBLEM_OK, PROBLEM_INVALID, PROBLEM_INFEASIBLE, PROBLEM_EMPTY = 0, 1, 2, 3
WS_COST, WS_MATCH, WS_TRACK, WS_WEIGHTS, WS_AP3 = 1, 2, 3, 4, 5
AP3_OK, AP3_GAP, AP3_INVALID, AP3_EMPTY = 0, 1, 2, 3
AP3_MAXN = 64
ND_OK, ND_INVALID, ND_INFEASIBLE, ND_EMPTY, ND_UNBALANCED, ND_GAP, ND_ABORTED = 0, 1, 2, 3, 6, 7, 8
ROW_COLS = 18
MAX_RODS = 256

# every symbol include/ptk.h declares (tests/test_abi.py checks header <-> library <-> this list)
SYMBOLS = [
    "ptk_version", "ptk_strerror", "ptk_last_cuda_error", "ptk_device_count",
    "ptk_undistort_batch", "ptk_cost_batch", "ptk_lsap_batch", "ptk_match_batch",
    "ptk_rows_batch", "ptk_rows_to_ends", "ptk_track_batch", "ptk_chain_tracks", "ptk_chain_rebase", "ptk_create_weights", "ptk_reorder_endpoints", "ptk_triangulate_points", "ptk_project_points", "ptk_npartite3_batch", "ptk_matchnd_sequence", "ptk_matchnd_sequence_workspace_bytes",
    "ptk_csv_write_rods", "ptk_format_double", "ptk_workspace_bytes",
    "ptk_csv_open", "ptk_csv_close", "ptk_csv_error", "ptk_csv_n_rows", "ptk_csv_n_cols", "ptk_csv_col_name",
    "ptk_csv_col_kind", "ptk_csv_col_copy", "ptk_csv_col_dict_size", "ptk_csv_col_dict", "ptk_parse_double",
    "ptk_ctx_create", "ptk_ctx_destroy", "ptk_reconstruct_host", "ptk_reconstruct_host_compact", "ptk_expand_rows_host",
    "ptk_reconstruct_device", "ptk_reconstruct_device_workspace_bytes",
    "ptk_reconstruct_shard", "ptk_reconstruct_shard_workspace_bytes", "ptk_chain_rebase_gathered",
    "ptk_measure_fp64_peak", "ptk_launch_count", "ptk_bounds_report", "ptk_timing_enable", "ptk_timing_read",
]


class NativeLibraryMissing(ImportError):
    pass


class PtkError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        msg = lib().ptk_strerror(status).decode()
        cuda = lib().ptk_last_cuda_error().decode()
        super().__init__(f"{where}: {msg}" + (f" [{cuda}]" if cuda and status == -2 else ""))


_lib = None


def lib() -> ctypes.CDLL:
    """The loaded library; raises NativeLibraryMissing if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: the CUDA library has not been built "
            "(run __graft_entry__.build() or `make -C particletracking_b200/csrc`). "
            "particletracking_b200 has no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, dbl, sz, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t, ctypes.c_longlong
    rigp = ctypes.POINTER(PtkRig)
    L.ptk_version.restype = i32
    L.ptk_strerror.restype = ctypes.c_char_p
    L.ptk_strerror.argtypes = [i32]
    L.ptk_last_cuda_error.restype = ctypes.c_char_p
    L.ptk_device_count.restype = i32
    L.ptk_launch_count.restype = ll
    L.ptk_launch_count.argtypes = [i32]
    L.ptk_workspace_bytes.restype = sz
    L.ptk_workspace_bytes.argtypes = [i32, ll, i32]
    L.ptk_undistort_batch.argtypes = [rigp, i32, i32] + [vp] * 6 + [vp]
    L.ptk_cost_batch.argtypes = [rigp, i32, i32] + [vp] * 8 + [vp, sz, vp]
    L.ptk_lsap_batch.argtypes = [i32, i32, i32] + [vp] * 7 + [vp]
    L.ptk_match_batch.argtypes = [rigp, i32, i32] + [vp] * 12 + [dbl, vp, vp, sz, vp]
    L.ptk_rows_batch.argtypes = [rigp, i32, i32] + [vp] * 8 + [vp, sz, vp]
    L.ptk_rows_to_ends.argtypes = [i32, i32, i32, i32, vp, vp, vp]
    L.ptk_track_batch.argtypes = [i32, i32, i32] + [vp] * 5 + [vp, sz, vp]
    L.ptk_chain_tracks.argtypes = [i32, i32, i32, vp, vp, vp, vp]
    L.ptk_chain_rebase.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp]
    L.ptk_create_weights.argtypes = [i32, i32, i32] + [vp] * 5 + [vp, sz, vp]
    L.ptk_reorder_endpoints.argtypes = [i32, i32, i32] + [vp] * 6 + [vp]
    L.ptk_triangulate_points.argtypes = [rigp, ll, vp, vp, i32, vp, vp]
    L.ptk_project_points.argtypes = [rigp, ll, vp, i32, vp, vp, vp]
    L.ptk_npartite3_batch.argtypes = [i32] * 4 + [vp] * 10 + [i32, ll, vp, sz, vp]
    L.ptk_matchnd_sequence_workspace_bytes.restype = sz
    L.ptk_matchnd_sequence_workspace_bytes.argtypes = [i32, i32, i32]
    L.ptk_matchnd_sequence.argtypes = [rigp, i32, i32, i32] + [vp] * 9 + [vp, sz, vp]
    L.ptk_csv_write_rods.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ll, i32, vp, vp, ctypes.c_char_p, vp, i32]
    L.ptk_format_double.argtypes = [dbl, ctypes.c_char_p]
    L.ptk_csv_open.argtypes = [ctypes.c_char_p, i32, ctypes.POINTER(vp)]
    L.ptk_csv_close.argtypes = [vp]
    L.ptk_csv_close.restype = None
    L.ptk_csv_error.argtypes = [vp]
    L.ptk_csv_error.restype = ctypes.c_char_p
    L.ptk_csv_n_rows.argtypes = [vp]
    L.ptk_csv_n_rows.restype = ll
    L.ptk_csv_n_cols.argtypes = [vp]
    L.ptk_csv_col_name.argtypes = [vp, i32]
    L.ptk_csv_col_name.restype = ctypes.c_char_p
    L.ptk_csv_col_kind.argtypes = [vp, i32]
    L.ptk_csv_col_copy.argtypes = [vp, i32, vp]
    L.ptk_csv_col_dict_size.argtypes = [vp, i32]
    L.ptk_csv_col_dict.argtypes = [vp, i32, i32, ctypes.POINTER(i32)]
    L.ptk_csv_col_dict.restype = vp
    L.ptk_parse_double.argtypes = [ctypes.c_char_p, i32, ctypes.POINTER(dbl)]
    L.ptk_ctx_create.argtypes = [i32, ctypes.POINTER(vp)]
    L.ptk_ctx_destroy.argtypes = [vp]
    L.ptk_ctx_destroy.restype = None
    L.ptk_reconstruct_device_workspace_bytes.restype = sz
    L.ptk_reconstruct_device_workspace_bytes.argtypes = [i32, i32, i32, i32]
    L.ptk_reconstruct_device.argtypes = [rigp, i32, i32, i32] + [vp] * 10 + [dbl, vp, i32] + [vp] * 4 + [vp, sz, vp]
    L.ptk_reconstruct_shard_workspace_bytes.restype = sz
    L.ptk_reconstruct_shard_workspace_bytes.argtypes = [i32, i32, i32, i32]
    L.ptk_reconstruct_shard.argtypes = [rigp, i32, i32, i32, i32] + [vp] * 10 + [dbl, vp] + [vp] * 5 + [vp, sz, vp]
    L.ptk_chain_rebase_gathered.argtypes = [i32, i32, i32, vp, vp, sz, sz, sz, vp, vp]
    L.ptk_reconstruct_host.argtypes = [vp, rigp, i32, i32, i32] + [vp] * 10 + [dbl, vp, i32] + [vp] * 4
    L.ptk_reconstruct_host_compact.argtypes = [vp, rigp, i32, i32, i32] + [vp] * 11 + [dbl, vp, i32] + [vp] * 4
    L.ptk_expand_rows_host.argtypes = [ll, i32] + [vp] * 7 + [i32]
    L.ptk_bounds_report.argtypes = [ctypes.POINTER(i32), ctypes.POINTER(ll), ctypes.POINTER(i32), ctypes.POINTER(ll),
                                    ctypes.POINTER(ll), i32]
    L.ptk_timing_enable.argtypes = [i32]
    L.ptk_timing_enable.restype = None
    L.ptk_timing_read.argtypes = [vp, vp, i32]
    L.ptk_measure_fp64_peak.argtypes = [i32, ctypes.POINTER(dbl), vp]
    for name in SYMBOLS:
        f = getattr(L, name)
        if f.restype is ctypes.c_int and name not in ("ptk_version", "ptk_device_count"):
            pass
    _lib = L
    return L


def check(status: int, where: str) -> None:
    if status != PTK_OK:
        raise PtkError(status, where)
