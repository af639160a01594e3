"""The parts of the reference's ``matchND`` (``reconstruct_3D/matchND.py``) on the hot path:
the shared triangulation block with SUM aggregation (:446-551) and ``create_weights``
(:141-217), both on the B200.

``npartite_matching`` (:45-138, a 3-index assignment ILP solved by pulp/CBC) is a "next" row
(SURVEY.md 8f.2): no oracle for it can run here (pulp and CBC are absent) and it is not the
LSAP the north star names.  ``match_frame`` therefore takes the solver as an argument.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import pandas as pd
import torch

from .. import engine
from ..rig import make_rig
from . import match2D


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

    Everything but the 3-index assignment runs on the GPU; ``solver(weights) -> (rod, cam1_ind,
    cam2_ind)`` must maximise the summed weights (the reference's ``npartite_matching``)."""
    if solver is None:
        raise NotImplementedError(
            "matchND.match_frame needs a 3-index assignment solver (the reference's pulp/CBC npartite_matching is "
            "out of scope, SURVEY.md 8f.2); pass solver=callable(weights) -> (rod, cam1_ind, cam2_ind)")
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
