"""Columnar reader for ``rods_df_{color}.csv``: ``libptk.so``'s memory-mapped,
multi-threaded parser in place of ``pd.read_csv(f_in, sep=",", index_col=0)`` (reference
match2D.py:601-602, 1053; matchND.py:317).  The returned DataFrame equals pandas' (values bit for
bit, dtypes, column names, index) -- ``tests/test_csv_reader.py`` checks that."""
from __future__ import annotations

import ctypes
from typing import Dict, List, Tuple

import numpy as np
import pandas as pd

from .. import _native as nat

KIND_INT64, KIND_FLOAT64, KIND_STRING, KIND_BOOL = 0, 1, 2, 3


class CsvParseError(ValueError):
    pass


def read_columns(path: str, n_threads: int = 0) -> Tuple[List[str], Dict[int, np.ndarray]]:
    """Parse ``path`` -> (column names incl. the index column, {column position: ndarray}).
    String columns come back as object arrays (missing fields are NaN)."""
    L = nat.lib()
    h = ctypes.c_void_p()
    rc = L.ptk_csv_open(str(path).encode(), int(n_threads), ctypes.byref(h))
    try:
        if rc != 0:
            msg = L.ptk_csv_error(h).decode() if h else "out of memory"
            if msg == "cannot open file":
                raise FileNotFoundError(f"[Errno 2] No such file or directory: '{path}'")
            raise CsvParseError(f"{path}: {msg}")
        n, nc = L.ptk_csv_n_rows(h), L.ptk_csv_n_cols(h)
        names, cols = [], {}
        for c in range(nc):
            names.append(L.ptk_csv_col_name(h, c).decode())
            kind = L.ptk_csv_col_kind(h, c)
            if kind in (KIND_INT64, KIND_BOOL):
                a = np.empty(n, dtype=np.int64)
            elif kind == KIND_FLOAT64:
                a = np.empty(n, dtype=np.float64)
            else:
                a = np.empty(n, dtype=np.int32)
            L.ptk_csv_col_copy(h, c, a.ctypes.data_as(ctypes.c_void_p))
            if kind == KIND_BOOL:
                a = a.astype(bool)
            elif kind == KIND_STRING:
                nd = L.ptk_csv_col_dict_size(h, c)
                ln = ctypes.c_int()
                words = []
                for k in range(nd):
                    p = L.ptk_csv_col_dict(h, c, k, ctypes.byref(ln))
                    words.append(ctypes.string_at(p, ln.value).decode())
                lut = np.array(words + [np.nan], dtype=object)
                a = lut[a]  # code -1 -> the trailing NaN
            cols[c] = a
        return names, cols
    finally:
        if h:
            L.ptk_csv_close(h)


def read_rods_csv(path: str, n_threads: int = 0) -> pd.DataFrame:
    """``pd.read_csv(path, sep=",", index_col=0)`` through the native reader."""
    names, cols = read_columns(path, n_threads)
    if not names:
        raise CsvParseError(f"{path}: no columns")
    data = {names[c]: cols[c] for c in range(1, len(names))}
    if len(data) != len(names) - 1:
        raise CsvParseError(f"{path}: duplicate column names are not supported")
    idx = pd.Index(cols[0], name=names[0] if names[0] and not names[0].startswith("Unnamed") else None)
    return pd.DataFrame(data, index=idx, copy=False)
