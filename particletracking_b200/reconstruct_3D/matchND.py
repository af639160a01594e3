"""The reference's ``matchND`` (``reconstruct_3D/matchND.py``) on the B200: the shared
triangulation block with SUM aggregation (:446-551), ``create_weights`` (:141-217),
``npartite_matching`` (:45-138) for two and three groups, ``match_frame`` (:372-632) and the
``assign`` CSV driver (:220-369).

The reference hands the 3-index assignment to pulp/CBC as a 0/1 programme; here it is solved
exactly on the GPU (``libptk.so``: dual block-coordinate descent over 2-D assignments, then a
reduced-cost search if a gap is left -- ``csrc/ptk_ap3.cuh``).  Where several optima exist the
returned one may differ from CBC's; the optimal value cannot.
"""
from __future__ import annotations

import os
import pathlib
import warnings
from typing import Callable, Iterable, Optional, Tuple

import numpy as np
import pandas as pd
import torch

from .. import _native as nat
from .. import engine
from ..rig import make_rig
from ..utils import csv_io
from ..utils import data_loading as dl
from . import match2D


class NpartiteGapWarning(RuntimeWarning):
    """The solver hit its work limits: the matching returned is feasible and the best found, but
    not proven optimal (``npartite_matching.last`` holds value and bound)."""


def npartite_matching(weights: np.ndarray, maximize: bool = True, solver=None) -> Tuple[np.ndarray, ...]:
    """Solve an n-partite matching problem (reference matchND.py:45-138) for 2 or 3 groups.

    Returns one index array per axis of ``weights`` (what ``np.where`` of the reference's 0/1
    solution gives: sorted by the first axis).  ``solver`` is accepted for signature
    compatibility and ignored (there is no LP solver here).  ``maximize=False`` solves the same
    programme the reference poses for it -- with "at most once" constraints that is the empty
    matching for positive weights, which is the bug the reference warns about."""
    if not maximize:
        warnings.warn("Minimization is currently not fully supported.Results might be incorrect.")
    w = np.ascontiguousarray(weights, dtype=np.float64)
    if not maximize:
        w = -w
    engine._require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    if w.ndim == 3:
        r = engine.npartite3_batch(torch.from_numpy(w[None]).to(dev))
        status, n = int(r.status[0]), int(r.n_sol[0])
        npartite_matching.last = dict(status=status, value=float(r.value[0]), bound=float(r.bound[0]),
                                      stats=r.stats[0].cpu().numpy())
        if status == nat.AP3_INVALID:
            raise ValueError("weights contain invalid numeric entries")
        if status == nat.AP3_GAP:
            warnings.warn(f"npartite_matching: work limit reached, value {float(r.value[0])!r} <= optimum <= "
                          f"{float(r.bound[0])!r}", NpartiteGapWarning)
        sol = r.sol[0, :, :n].cpu().numpy().astype(np.int64)
        return sol[0], sol[1], sol[2]
    if w.ndim == 2:
        n0, n1 = w.shape
        if n0 == 0 or n1 == 0:
            return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
        if not np.isfinite(w).all():
            raise ValueError("weights contain invalid numeric entries")
        # "at most once" lets negative weights stay unmatched: clamp them, drop them afterwards
        row, col, n_out, _ = engine.lsap_batch(torch.from_numpy(-np.maximum(w, 0.0)[None]).to(dev))
        n = int(n_out[0])
        ri, ci = row[0, :n].cpu().numpy().astype(np.int64), col[0, :n].cpu().numpy().astype(np.int64)
        keep = w[ri, ci] >= 0
        return ri[keep], ci[keep]
    raise NotImplementedError("npartite_matching: only 2 and 3 groups are implemented (the reference's callers use 3)")


def create_weights(p_3D: np.ndarray, p_3D_prev: np.ndarray, repr_errs: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Weights = 3D displacement x reprojection error (reference matchND.py:141-217):
    ``p_3D[N1,N2,4,3]``, ``p_3D_prev[Np,2,3]``, ``repr_errs[N1,N2,4,2]`` ->
    ``(weights[Np,N1,N2], point_choices[Np,N1,N2])``."""
    engine._require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    w, ch = engine.create_weights(t(p_3D), t(p_3D_prev), t(repr_errs))
    return w.cpu().numpy(), ch.cpu().numpy()


def triangulate_frame(data: pd.DataFrame, cam1_name: str, cam2_name: str, frame: int, calibration: dict,
                      P1, P2, rot, trans, r1, r2, t1, t2):
    """matchND.py:446-551: all rod pairs x 4 combos of one frame with the SUM error.
    Returns ``(rods_cam1[N1,2,2], rods_cam2[N2,2,2], costs[N1,N2], p_triang[N1,N2,4,3],
    rep_errs[N1,N2,4,2])`` or None if a camera has no rods."""
    fidx = match2D._FrameIndex(data)
    v1 = data[match2D._cam_cols(cam1_name)].to_numpy(dtype=np.float64)
    v2 = data[match2D._cam_cols(cam2_name)].to_numpy(dtype=np.float64)
    a, b, _ = match2D._select(data, fidx, v1, v2, frame, True)
    if len(a) == 0 or len(b) == 0:
        return None
    rig = make_rig(calibration, P1, P2, rot, trans, r1, r2, t1, t2, agg_sum=True)
    c1, c2, n1, n2 = match2D._pack([(a, b)])
    cost, _, pw, re = engine.cost_batch(rig, c1, c2, n1, n2, want_choice=False, want_combos=True)
    N1, N2 = len(a), len(b)
    return (a.reshape(-1, 2, 2), b.reshape(-1, 2, 2), cost[0, :N1, :N2].cpu().numpy(),
            pw[0, :N1, :N2].cpu().numpy(), re[0, :N1, :N2].cpu().numpy())


def match_frame(data: pd.DataFrame, data_last_frame: pd.DataFrame, cam1_name: str, cam2_name: str, frame: int,
                color: str, calibration: dict, P1, P2, rot, trans, r1, r2, t1, t2, *,
                solver: Optional[Callable[[np.ndarray], Tuple[np.ndarray, np.ndarray, np.ndarray]]] = None):
    """Matches, triangulates and tracks the rods of one frame (reference matchND.py:372-632).

    ``solver(weights) -> (rod, cam1_ind, cam2_ind)`` defaults to this module's
    ``npartite_matching``; row bookkeeping follows matchND.py:564-632 (rows indexed by the previous
    frame's rod id, NaN fill-in, ``ValueError`` for unbalanced data)."""
    if solver is None:
        solver = npartite_matching
    tri = triangulate_frame(data, cam1_name, cam2_name, frame, calibration, P1, P2, rot, trans, r1, r2, t1, t2)
    if tri is None:
        return None
    rods1, rods2, costs, p_triang, rep_errs = tri
    last = data_last_frame.loc[data_last_frame.frame == frame - 1, ["x1", "y1", "z1", "x2", "y2", "z2"]]
    last_points = last.to_numpy().reshape((-1, 2, 3))
    weights, point_choices = create_weights(p_triang, last_points, rep_errs)
    rod, cam1_ind, cam2_ind = solver(weights)
    nprev = weights.shape[0]
    rod_nums = list(range(nprev))
    idx_out = np.full((3, nprev), np.nan)
    idx_out[:, rod] = np.stack([rod, cam1_ind, cam2_ind], axis=0)
    if np.isnan(idx_out).any():  # matchND.py:574-583
        idx_out[0, np.isnan(idx_out[0])] = [r for r in rod_nums if r not in rod]
        idx_out[1, np.isnan(idx_out[1])] = [r for r in rod_nums if r not in cam1_ind]
        idx_out[2, np.isnan(idx_out[2])] = [r for r in rod_nums if r not in cam2_ind]
    idx_out = idx_out.astype(int)
    point_choices = point_choices.astype(int)
    out = np.zeros((nprev, 18))
    for rod_id in range(nprev):
        idx_r, i1, i2 = idx_out[:, rod_id]
        try:
            i3 = point_choices[idx_r, i1, i2]
            e1, e2 = p_triang[i1, i2, i3], p_triang[i1, i2, 3 - i3]
        except IndexError:
            raise ValueError("Unbalanced dataset given. Please use dataset with the same number of rods on every frame.")
        out[idx_r, 0:6] = np.concatenate((e1, e2))
        out[idx_r, 6:9] = (e1 + e2) / 2.0
        out[idx_r, 9] = np.linalg.norm(e1 - e2, axis=0)
        out[idx_r, 10:14] = rods1[i1].ravel()
        out[idx_r, 14:] = rods2[i2].ravel()
    cols = [*match2D.COLS_3D, *match2D._cam_cols(cam1_name), *match2D._cam_cols(cam2_name)]
    df = pd.DataFrame(out, columns=cols)
    df["frame"] = frame
    df["color"] = color
    df["particle"] = rod_nums
    seen_cols = [c for c in data.columns if "seen" in c]
    if seen_cols:
        df[seen_cols] = 1
    return df, costs[idx_out[1, :], idx_out[2, :]], out[:, 9]


def assign(input_folder: str, output_folder: str, colors: Iterable[str], cam1_name: str = "gp1", cam2_name: str = "gp2",
           frame_numbers: Iterable[int] = None, calibration_file: str = None, transformation_file: str = None, solver=None):
    """Matches, triangulates and tracks rods over frames from ``rods_df_{color}.csv`` files
    (reference matchND.py:220-369): the first frame through ``match2D.match_frame``, every later
    frame through ``match_frame`` against the previous result.  Frames are sequential by
    construction ("replicas only" along frames).  Returns ``(reprojection errors, rod
    lengths)``."""
    if calibration_file is None or transformation_file is None:
        raise FileNotFoundError("calibration_file / transformation_file: the reference's defaults "
                                "(example_calibration/Matlab/*.json) are not part of this package; pass them explicitly")
    if not os.path.exists(output_folder):
        os.mkdir(output_folder)
    calibration = dl.load_camera_calibration(str(pathlib.Path(calibration_file)))
    transforms = dl.load_world_transformation(str(pathlib.Path(transformation_file)))
    from scipy.spatial.transform import Rotation

    # matchND.py:297-312
    r1 = np.eye(3)
    t1 = np.expand_dims(np.array([0.0, 0.0, 0.0]), 1)
    P1 = (np.vstack((r1.T, t1.T)) @ calibration["CM1"].T).T
    r2, t2 = calibration["R"], calibration["T"]
    P2 = (np.vstack((r2.T, t2.T)) @ calibration["CM2"].T).T
    rot = Rotation.from_matrix(transforms["rotation"])
    trans = transforms["translation"]
    colors = list(colors)
    if solver is None and colors and len(frame_numbers) > 0:
        return _assign_device(input_folder, output_folder, colors, cam1_name, cam2_name, list(frame_numbers), calibration,
                              P1, P2, rot, trans, r1, r2, t1, t2)
    all_repr_errs, all_rod_lengths = [], []
    for color in colors:
        data = csv_io.read_rods_csv(input_folder + f"/rods_df_{color}.csv")
        parts = []
        tmp_df = None
        for fn in range(len(frame_numbers)):
            frame = frame_numbers[fn]
            if fn == 0:
                res = match2D.match_frame(data, cam1_name, cam2_name, frame, color, calibration, P1, P2, rot, trans, r1, r2,
                                          t1, t2, renumber=True)
            else:
                res = match_frame(data, tmp_df, cam1_name, cam2_name, frame, color, calibration, P1, P2, rot, trans, r1, r2,
                                  t1, t2, solver=solver)
            tmp_df, tmp_costs, tmp_lengths = res  # TypeError on None, as in the reference (:322, :341)
            parts.append(tmp_df)
            all_repr_errs.append(tmp_costs)
            all_rod_lengths.append(tmp_lengths)
        df_out = pd.concat(parts) if parts else pd.DataFrame()
        df_out.reset_index(drop=True, inplace=True)
        df_out.to_csv(os.path.join(output_folder, f"rods_df_{color}.csv"), sep=",")
    return np.asarray(all_repr_errs), np.asarray(all_rod_lengths)


def _raise_for_nd_status(status: np.ndarray):
    """status[F, C] of ptk_matchnd_sequence -> the exception the reference's loop would have raised
    first (colours outer, frames inner, matchND.py:314-360)."""
    F, C = status.shape
    for c in range(C):
        for f in range(F):
            st = int(status[f, c])
            if st in (nat.ND_OK, nat.ND_GAP):
                continue
            if st == nat.ND_EMPTY:  # match_frame returned None (match2D.py:838-840, matchND.py:465-467)
                raise TypeError("cannot unpack non-iterable NoneType object")
            if st == nat.ND_UNBALANCED:
                raise ValueError("Unbalanced dataset given. Please use dataset with the same number of rods on every frame.")
            if st == nat.ND_INFEASIBLE:
                raise ValueError("cost matrix is infeasible")
            if st == nat.ND_INVALID:
                raise ValueError("matrix contains invalid numeric entries" if f == 0 else
                                 "weights contain invalid numeric entries")
            raise RuntimeError(f"matchND sequence: status {st} at frame index {f}, colour index {c}")


def _assign_device(input_folder, output_folder, colors, cam1_name, cam2_name, frames, calibration, P1, P2, rot, trans,
                   r1, r2, t1, t2):
    """``assign`` with every colour's whole frame sequence in one enqueue-only library call
    (``ptk_matchnd_sequence``): native CSV reader -> vectorised selection -> device loop -> native
    CSV writer."""
    engine._require_cuda()
    rig = make_rig(calibration, P1, P2, rot, trans, r1, r2, t1, t2)
    from concurrent.futures import ThreadPoolExecutor

    def load(color):
        data = match2D._Table.read(input_folder + f"/rods_df_{color}.csv")
        v1 = match2D._cols(data, match2D._cam_cols(cam1_name))
        v2 = match2D._cols(data, match2D._cam_cols(cam2_name))
        packed = match2D._select_pack(np.asarray(match2D._col(data, "frame")), v1, v2, frames, True)
        if packed is None:
            raise ValueError("frame numbers must be unique")
        return data, packed

    # reader and writer release the GIL: colours are read / written side by side
    pool = ThreadPoolExecutor(max_workers=min(len(colors), match2D._CSV_WORKERS))
    try:
        loaded = list(pool.map(load, colors))
        tables, packs = [l[0] for l in loaded], [l[1] for l in loaded]
        F, C = len(frames), len(colors)
        nmax = max(p[0].shape[1] for p in packs)
        c1 = np.full((F, C, nmax, 4), np.nan)
        c2 = np.full((F, C, nmax, 4), np.nan)
        n1 = np.zeros((F, C), dtype=np.int32)
        n2 = np.zeros((F, C), dtype=np.int32)
        for ci, (a, b, m1, m2, _) in enumerate(packs):
            c1[:, ci, : a.shape[1]] = a
            c2[:, ci, : b.shape[1]] = b
            n1[:, ci], n2[:, ci] = m1, m2
        dev = torch.device("cuda", torch.cuda.current_device())
        t = lambda x: torch.from_numpy(x).to(dev)
        rows, costs, n_rows, status, stats = engine.matchnd_sequence(rig, t(c1.reshape(F, C, nmax, 2, 2)), t(c2.reshape(F, C, nmax, 2, 2)),
                                                                     t(n1), t(n2))
        rows, costs, n_rows, status = rows.cpu().numpy(), costs.cpu().numpy(), n_rows.cpu().numpy(), status.cpu().numpy()
        assign.last = dict(status=status, solver_stats=stats.cpu().numpy())
        _raise_for_nd_status(status)
        if (status == nat.ND_GAP).any():
            warnings.warn(f"matchND.assign: the 3-index assignment hit its work limit in {int((status == nat.ND_GAP).sum())} "
                          "(frame, colour) problems; their matchings are the best found, not proven optimal", NpartiteGapWarning)
        cols = [*match2D.COLS_3D, *match2D._cam_cols(cam1_name), *match2D._cam_cols(cam2_name)]

        def store(ci):
            n = n_rows[:, ci]
            mask = np.arange(nmax)[None, :] < n[:, None]
            flat = rows[:, ci][mask]
            m = dict(rows=flat, frame=np.repeat(np.asarray(frames), n), particle=np.broadcast_to(np.arange(nmax), (F, nmax))[mask],
                     cols=cols, seen_cols=[c for c in tables[ci].columns if "seen" in c])
            match2D._write_csv(os.path.join(output_folder, f"rods_df_{colors[ci]}.csv"), colors[ci], m)
            cuts = np.cumsum(n)[:-1]
            return np.split(costs[:, ci][mask], cuts), np.split(flat[:, 9].copy(), cuts)

        stored = list(pool.map(store, range(C)))
    finally:
        pool.shutdown()
    all_repr_errs = [e for s_ in stored for e in s_[0]]
    all_rod_lengths = [l for s_ in stored for l in s_[1]]
    return np.asarray(all_repr_errs), np.asarray(all_rod_lengths)
