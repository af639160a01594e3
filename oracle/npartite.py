"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's ``npartite_matching``
(``reconstruct_3D/matchND.py:45-138``) and of ``matchND.match_frame``'s bookkeeping around it
(:553-632).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg
may import this module.

PARITY UNPINNED for the solver itself: the reference hands the 0/1 programme

    max  sum_t w[t] x[t]     s.t.  for every axis d and index v:  sum_{t : t[d] = v} x[t] <= 1

to pulp's bundled CBC (``pulp>=2.7.0``, ``ParticleDetection/pyproject.toml:48``), and neither pulp
nor a CBC binary exists in this image, so the reference's own solver cannot be run
here and its tests (``tests/test_reconstruct_3D/test_match_nd.py:36-47``) only bound the indices.
What is restated below is the programme, constraint by constraint, solved by HiGHS through
``scipy.optimize.milp``; parity for the GPU solver is therefore on the OPTIMAL OBJECTIVE VALUE
(unique) and on feasibility, not on which optimum is returned when several exist.
``brute_force`` is an independent exhaustive check for tiny cases.
"""
from __future__ import annotations

import itertools
from typing import Tuple

import numpy as np


def npartite_matching(weights: np.ndarray, maximize: bool = True) -> Tuple[np.ndarray, ...]:
    """matchND.py:45-138 with HiGHS in place of CBC.  Returns ``np.where(x)``: one index array per
    axis, entries sorted lexicographically (so by the first axis)."""
    from scipy.optimize import Bounds, LinearConstraint, milp
    from scipy.sparse import coo_matrix

    w = np.asarray(weights, dtype=np.float64)
    dims = w.shape
    n = w.size
    idx = np.indices(dims).reshape(len(dims), -1)
    rows, r0 = [], 0
    for d, dim in enumerate(dims):  # matchND.py:121-126: one "<= 1" row per (axis, index)
        rows.append(idx[d] + r0)
        r0 += dim
    A = coo_matrix((np.ones(n * len(dims)), (np.concatenate(rows), np.tile(np.arange(n), len(dims)))),
                   shape=(r0, n)).tocsr()
    c = -w.ravel() if maximize else w.ravel()
    res = milp(c, constraints=LinearConstraint(A, -np.inf, 1.0), integrality=np.ones(n), bounds=Bounds(0, 1),
               options={"mip_rel_gap": 0.0})
    if res.x is None:
        raise RuntimeError(f"HiGHS failed: {res.message}")
    x = np.round(res.x).reshape(dims)
    return np.where(x)


def objective(weights: np.ndarray, sol) -> float:
    return float(np.asarray(weights)[tuple(np.asarray(s, dtype=np.int64) for s in sol)].sum())


def is_feasible(dims, sol) -> bool:
    """Every index of every axis used at most once, all in range."""
    for d, s in enumerate(sol):
        s = np.asarray(s)
        if len(s) and (s.min() < 0 or s.max() >= dims[d] or len(np.unique(s)) != len(s)):
            return False
    return len({len(s) for s in sol}) == 1


def brute_force(weights: np.ndarray) -> float:
    """Optimal value of the 3-index programme by enumeration (tiny sizes only): the smallest axis is
    matched against all injective choices of the other two."""
    w = np.asarray(weights, dtype=np.float64)
    order = np.argsort(w.shape)
    w = np.transpose(w, order)
    n0, n1, n2 = w.shape
    best = 0.0
    i = np.arange(n0)
    for js in itertools.permutations(range(n1), n0):
        for ks in itertools.permutations(range(n2), n0):
            v = w[i, js, ks]
            best = max(best, float(v[v > 0].sum()))  # "<= 1": a triple may stay unused
    return best


def match_frame_rows(p_triang: np.ndarray, rods1: np.ndarray, rods2: np.ndarray, costs: np.ndarray,
                     point_choices: np.ndarray, sol, nprev: int):
    """matchND.py:553-632 after the solver: index bookkeeping (NaN fill-in :574-583), end-point
    choice (:587-600) and the 18-column rows indexed by the previous frame's rod id.
    ``p_triang[N1,N2,4,3]``, ``rods1[N1,2,2]``, ``rods2[N2,2,2]``, ``costs[N1,N2]``,
    ``point_choices[Np,N1,N2]``, ``sol = (rod, cam1_ind, cam2_ind)``.
    Returns ``(out[nprev,18], costs_out[nprev])``; raises the reference's ValueError."""
    rod, cam1_ind, cam2_ind = (np.asarray(s) for s in sol)
    rod_nums = list(range(nprev))
    idx_out = np.full((3, nprev), np.nan)
    idx_out[:, rod] = np.stack([rod, cam1_ind, cam2_ind], axis=0)
    if np.isnan(idx_out).any():
        idx_out[0, np.isnan(idx_out[0])] = [r for r in rod_nums if r not in rod]
        idx_out[1, np.isnan(idx_out[1])] = [r for r in rod_nums if r not in cam1_ind]
        idx_out[2, np.isnan(idx_out[2])] = [r for r in rod_nums if r not in cam2_ind]
    idx_out = idx_out.astype(int)
    pc = point_choices.astype(int)
    out = np.zeros((nprev, 18))
    for rod_id in range(nprev):
        idx_r, i1, i2 = idx_out[:, rod_id]
        try:
            i3 = pc[idx_r, i1, i2]
            e1, e2 = p_triang[i1, i2, i3], p_triang[i1, i2, 3 - i3]
        except IndexError:
            raise ValueError("Unbalanced dataset given. Please use dataset with the same number of rods on every frame.")
        out[idx_r, 0:6] = np.concatenate((e1, e2))
        out[idx_r, 6:9] = (e1 + e2) / 2.0
        out[idx_r, 9] = np.linalg.norm(e1 - e2, axis=0)
        out[idx_r, 10:14] = rods1[i1].ravel()
        out[idx_r, 14:] = rods2[i2].ravel()
    return out, costs[idx_out[1, :], idx_out[2, :]]


def match_frame_arrays(rods1: np.ndarray, rods2: np.ndarray, prev_ends: np.ndarray, calibration, P1, P2, rot, trans,
                       r1, r2, t1, t2, solver=npartite_matching):
    """matchND.match_frame (matchND.py:446-632) on arrays: ``rods1[N1,2,2]``, ``rods2[N2,2,2]`` (already
    filtered), ``prev_ends[Np,2,3]`` = the previous frame's ``x1..z2``.  Returns
    ``(rows[Np,18], costs[Np], lengths[Np], weights, sol)``."""
    from . import port

    m = port.match_arrays(rods1, rods2, calibration, P1, P2, rot, trans, r1, r2, t1, t2, agg="sum")
    weights, choices = port.create_weights(m["p_world"], prev_ends, m["rep_errs"])
    sol = solver(weights)
    out, costs = match_frame_rows(m["p_world"], rods1, rods2, m["cost"], choices, sol, weights.shape[0])
    return out, costs, out[:, 9].copy(), weights, sol
