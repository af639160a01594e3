"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_build/libptk_oracle.so
(the plain-C restatement in oracle/ptk_oracle.c).  See that file's header for
the parity status and who may load it."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from particletracking_b200.rig import PtkRig

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libptk_oracle.so")
_lib = None

_d = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_ll = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ptk_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.ptko_num_threads.restype = ctypes.c_int
    return _lib


def _opt(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def undistort_rods(CM, dist, rods):
    rods = np.ascontiguousarray(rods, dtype=np.float64)
    k = np.zeros(14)
    k[: np.size(dist)] = np.ravel(dist)
    out = np.empty_like(rods)
    lib().ptko_undistort_rods(_opt(np.ascontiguousarray(CM, dtype=np.float64)), _opt(k),
                              ctypes.c_int(len(rods)), _opt(rods), _opt(out))
    return out


def triangulate(P1, P2, x1, y1, x2, y2):
    X = np.zeros(3)
    Xh = np.zeros(4)
    st = np.zeros(2, dtype=np.int64)
    lib().ptko_triangulate(_opt(np.ascontiguousarray(P1, dtype=np.float64)),
                           _opt(np.ascontiguousarray(P2, dtype=np.float64)),
                           ctypes.c_double(x1), ctypes.c_double(y1), ctypes.c_double(x2), ctypes.c_double(y2),
                           _opt(X), _opt(Xh), _opt(st))
    return X, Xh, st


def project(R, t, CM, dist, X):
    k = np.zeros(14)
    k[: np.size(dist)] = np.ravel(dist)
    X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 3)
    uv = np.zeros((len(X), 2))
    R = np.ascontiguousarray(R, dtype=np.float64)
    t = np.ascontiguousarray(np.ravel(t), dtype=np.float64)
    CM = np.ascontiguousarray(CM, dtype=np.float64)
    for n in range(len(X)):
        lib().ptko_project(_opt(R), _opt(t), _opt(CM), _opt(k), _opt(X[n]), _opt(uv[n]))
    return uv


def cost_problem(rig: PtkRig, cam1, cam2, want_combos=False):
    cam1 = np.ascontiguousarray(cam1, dtype=np.float64)
    cam2 = np.ascontiguousarray(cam2, dtype=np.float64)
    N1, N2 = len(cam1), len(cam2)
    cost = np.zeros((N1, N2))
    choice = np.zeros((N1, N2), dtype=np.uint8)
    pw = np.zeros((N1, N2, 4, 3)) if want_combos else None
    re = np.zeros((N1, N2, 4, 2)) if want_combos else None
    st = np.zeros(2, dtype=np.int64)
    lib().ptko_cost_problem(ctypes.byref(rig), N1, N2, _opt(cam1), _opt(cam2), N2,
                            _opt(cost), _opt(choice), _opt(pw), _opt(re), _opt(st))
    return dict(cost=cost, choice=choice.astype(bool), p_world=pw, rep_errs=re, stats=st)


def lsap(cost):
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    n = min(nr, nc)
    ri = np.full(max(n, 1), -1, dtype=np.int32)
    ci = np.full(max(n, 1), -1, dtype=np.int32)
    st = lib().ptko_lsap(nr, nc, _opt(cost), nc, _opt(ri), _opt(ci))
    return st, ri[:n], ci[:n]


def match_batch(rig: PtkRig, cam1, cam2, n1, n2, want_cost=False, nthreads=0):
    cam1 = np.ascontiguousarray(cam1, dtype=np.float64)
    cam2 = np.ascontiguousarray(cam2, dtype=np.float64)
    n1 = np.ascontiguousarray(n1, dtype=np.int32)
    n2 = np.ascontiguousarray(n2, dtype=np.int32)
    P, Nmax = cam1.shape[:2]
    cost = np.full((P, Nmax, Nmax), np.nan) if want_cost else None
    row_ind = np.zeros((P, Nmax), dtype=np.int32)
    col_ind = np.zeros((P, Nmax), dtype=np.int32)
    n_out = np.zeros(P, dtype=np.int32)
    status = np.zeros(P, dtype=np.int32)
    rows = np.zeros((P, Nmax, 18))
    mc = np.zeros((P, Nmax))
    st = np.zeros(2, dtype=np.int64)
    lib().ptko_match_batch(ctypes.byref(rig), P, Nmax, _opt(cam1), _opt(cam2), _opt(n1), _opt(n2),
                           _opt(cost), _opt(row_ind), _opt(col_ind), _opt(n_out), _opt(status),
                           _opt(rows), _opt(mc), int(nthreads), _opt(st))
    return dict(cost=cost, row_ind=row_ind, col_ind=col_ind, n_out=n_out, status=status,
                rows=rows, match_cost=mc, stats=st)


def track_cost(cur, nxt):
    cur = np.ascontiguousarray(cur, dtype=np.float64)
    nxt = np.ascontiguousarray(nxt, dtype=np.float64)
    N = cur.shape[0]
    cost = np.zeros((N, N))
    lib().ptko_track_cost(N, _opt(cur), _opt(nxt), _opt(cost))
    return cost


def track_batch(ends, want_cost=False, nthreads=0):
    ends = np.ascontiguousarray(ends, dtype=np.float64)
    C, F, N = ends.shape[:3]
    cost = np.zeros((C, F - 1, N, N)) if want_cost else None
    col = np.zeros((C, F - 1, N), dtype=np.int32)
    tot = np.zeros((C, F - 1))
    status = np.zeros((C, F - 1), dtype=np.int32)
    lib().ptko_track_batch(C, F, N, _opt(ends), _opt(cost), _opt(col), _opt(tot), _opt(status), int(nthreads))
    return dict(cost=cost, col_ind=col, total_cost=tot, status=status)


def chain_tracks(col_ind, ids_in=None):
    col_ind = np.ascontiguousarray(col_ind, dtype=np.int32)
    C, Fm1, N = col_ind.shape
    out = np.zeros((C, Fm1 + 1, N), dtype=np.int32)
    ids = None if ids_in is None else np.ascontiguousarray(ids_in, dtype=np.int32)
    lib().ptko_chain_tracks(C, Fm1 + 1, N, _opt(col_ind), _opt(ids), _opt(out))
    return out


def create_weights(p3, prev, re):
    p3 = np.ascontiguousarray(p3, dtype=np.float64)
    prev = np.ascontiguousarray(prev, dtype=np.float64)
    re = np.ascontiguousarray(re, dtype=np.float64)
    Np, (N1, N2) = prev.shape[0], p3.shape[:2]
    w = np.zeros((Np, N1, N2))
    ch = np.zeros((Np, N1, N2))
    lib().ptko_create_weights(Np, N1, N2, _opt(p3), _opt(prev), _opt(re), _opt(w), _opt(ch))
    return w, ch


def num_threads() -> int:
    return int(lib().ptko_num_threads())
