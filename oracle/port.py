"""TEST INFRASTRUCTURE ONLY -- CPU oracle ("port") of the reference's stereo
match -> DLT -> reproject -> LSAP -> rows path and of its tracking assignment.

This is a restatement, in vectorised numpy, of what the reference's Python
does around the same third-party calls it makes (``cv2.undistortImagePoints``,
``cv2.triangulatePoints``, ``cv2.projectPoints``,
``scipy.optimize.linear_sum_assignment``).  Each function cites the reference
lines it follows (paths relative to
``ParticleDetection/src/ParticleDetection/reconstruct_3D/``).

Parity status: PINNED.  The reference's own tests hold no numeric goldens for
this path, so the port is pinned against *outputs of the
reference itself run in the build container* -- live in
``tests/test_oracle_vs_reference.py`` when ``/root/reference`` is present, and
through the committed fixtures under ``tests/golden/`` (made by
``tests/golden/make_golden.py``) everywhere else.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this module.  The product (``particletracking_b200``) never
does; it fails loudly without its CUDA library.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np


# --------------------------------------------------------------------------
# A.2  undistortion, including the reference's reshape quirk
# --------------------------------------------------------------------------
def undistort_quirk(rods: np.ndarray, CM: np.ndarray, dist: np.ndarray) -> np.ndarray:
    """match2D.py:847-863 (identical at matchND.py:474-490).

    ``rods[N,2,2]`` is handed to OpenCV as ``rods.reshape(2,-1)``, i.e. the
    flat coordinate list split in two halves, so pseudo-point j pairs
    ``flat[j]`` with ``flat[j+2N]``; the result is written back in the same
    flat positions.  For N == 1 OpenCV reads the 2x2 row-wise and the
    write-back transposes it.
    """
    import cv2

    tmp = cv2.undistortImagePoints(rods.reshape(2, -1), CM, dist).squeeze()
    flat = np.concatenate([tmp[:, 0], tmp[:, 1]])
    return flat.reshape(-1, 2, 2)


# --------------------------------------------------------------------------
# A.3-A.6  all pairs x 4 end-point combos -> triangulate -> reproject -> cost
# --------------------------------------------------------------------------
def pair_combos(a: np.ndarray, b: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """match2D.py:865-887: row-major (i, j, combo) with combo -> (e1, e2) =
    (0,0),(0,1),(1,0),(1,1).  Returns the cam1 and cam2 points ``[M,2]``."""
    N1, N2 = len(a), len(b)
    e1 = np.array([0, 0, 1, 1])
    e2 = np.array([0, 1, 0, 1])
    pa = np.broadcast_to(a[:, None, e1, :], (N1, N2, 4, 2)).reshape(-1, 2)
    pb = np.broadcast_to(b[None, :, e2, :], (N1, N2, 4, 2)).reshape(-1, 2)
    return np.ascontiguousarray(pa), np.ascontiguousarray(pb)


def triangulate_reproject(
    rods1: np.ndarray,
    rods2: np.ndarray,
    calibration: Dict[str, np.ndarray],
    P1: np.ndarray,
    P2: np.ndarray,
    r1: np.ndarray,
    r2: np.ndarray,
    t1: np.ndarray,
    t2: np.ndarray,
):
    """match2D.py:846-910.  Returns ``(p_cam[M,3], d[M,2])``: triangulated
    points in camera-1 coordinates and the per-camera reprojection distances."""
    import cv2

    u1 = undistort_quirk(rods1, calibration["CM1"], calibration["dist1"])
    u2 = undistort_quirk(rods2, calibration["CM2"], calibration["dist2"])
    pu1, pu2 = pair_combos(u1, u2)
    po1, po2 = pair_combos(rods1, rods2)
    # match2D.py:889-895
    ph = cv2.triangulatePoints(P1, P2, pu1.T.copy(), pu2.T.copy())
    p = (ph[0:3] / ph[3]).T.copy()
    # match2D.py:898-903
    q1 = cv2.projectPoints(p, r1, t1, calibration["CM1"], calibration["dist1"])[0].reshape(-1, 2)
    q2 = cv2.projectPoints(p, r2, t2, calibration["CM2"], calibration["dist2"])[0].reshape(-1, 2)
    # match2D.py:905-910 (norm over xy per camera)
    d = np.stack([np.linalg.norm(po1 - q1, axis=1), np.linalg.norm(po2 - q2, axis=1)], axis=1)
    return p, d


def world_transform(p: np.ndarray, rot, trans: np.ndarray) -> np.ndarray:
    """match2D.py:913 -- ``rot`` is a scipy Rotation or a 3x3 matrix."""
    if hasattr(rot, "apply"):
        return rot.apply(p) + trans
    return p @ np.asarray(rot).T + trans


def match_arrays(
    rods1: np.ndarray,
    rods2: np.ndarray,
    calibration: Dict[str, np.ndarray],
    P1: np.ndarray,
    P2: np.ndarray,
    rot,
    trans: np.ndarray,
    r1: np.ndarray,
    r2: np.ndarray,
    t1: np.ndarray,
    t2: np.ndarray,
    agg: str = "mean",
    max_repr_err: Optional[float] = None,
) -> Dict[str, np.ndarray]:
    """match2D.py:843-983 (``renumber=True``) on already-filtered
    ``rods_cam1[N1,2,2]``, ``rods_cam2[N2,2,2]``.

    ``agg='sum'`` gives matchND's error (matchND.py:525).  ``max_repr_err`` is
    the north-star threshold extension: a mask over the
    returned ``costs``; nothing else changes.
    """
    from scipy.optimize import linear_sum_assignment

    N1, N2 = len(rods1), len(rods2)
    p, d = triangulate_reproject(rods1, rods2, calibration, P1, P2, r1, r2, t1, t2)
    err = d.mean(axis=1) if agg == "mean" else d.sum(axis=1)
    pw = world_transform(p, rot, trans).reshape(N1, N2, 4, 3)
    e = err.reshape(-1, 4)
    # match2D.py:921-939
    s03 = np.sum(e[:, 0::3], axis=1)
    s12 = np.sum(e[:, 1:3], axis=1)
    cost = np.minimum(s03, s12).reshape(N1, N2)
    choice = (s03 <= s12).reshape(N1, N2)
    row_ind, col_ind = linear_sum_assignment(cost)
    costs = cost[row_ind, col_ind]
    # match2D.py:963-983 -- NB the loop index, not row_ind, selects the cam1 rod
    n = len(row_ind)
    rows = np.zeros((n, 18))
    for m in range(n):
        i2 = col_ind[m]
        if choice[m, i2]:
            ends = pw[m, i2, 0::3, :]
            c1, c2 = rods1[m], rods2[i2]
        else:
            ends = pw[m, i2, 1:3, :]
            c1, c2 = rods1[m, ::-1], rods2[i2, ::-1]
        rows[m, 0:6] = ends.ravel()
        rows[m, 6:9] = ends.sum(axis=0) / 2
        rows[m, 9] = np.linalg.norm(np.diff(ends, axis=0))
        rows[m, 10:14] = c1.ravel()
        rows[m, 14:18] = c2.ravel()
    out = {
        "cost": cost,
        "choice": choice,
        "row_ind": row_ind,
        "col_ind": col_ind,
        "costs": costs,
        "rows": rows,
        "p_world": pw,
        "rep_errs": d.reshape(N1, N2, 4, 2),
    }
    if max_repr_err is not None:
        out["keep"] = costs <= max_repr_err
    return out


# --------------------------------------------------------------------------
# A.1  row selection from the DataFrame; DataFrame-level entry points
# --------------------------------------------------------------------------
def select_rods(data, cam_name: str, frame: int) -> np.ndarray:
    """match2D.py:812-830,843-844 for one camera: rows of ``frame``, drop
    all-NaN rows, then all-zero rows; ``[N,2,2]``."""
    cols = [f"x1_{cam_name}", f"y1_{cam_name}", f"x2_{cam_name}", f"y2_{cam_name}"]
    v = data.loc[data.frame == frame, cols].to_numpy(dtype=np.float64)
    v = v[~np.isnan(v).all(axis=1)]
    v = v[(v != 0).any(axis=1)]
    return v.reshape(-1, 2, 2)


def match_arrays_keep_pairs(rods1, rods2, calibration, P1, P2, rot, trans, r1, r2, t1, t2):
    """match2D.py:876-884,941-983 (``renumber=False``): rod i of camera 1 stays paired with
    rod i of camera 2; only the end-point combination is chosen.  ``costs`` is the per-pair
    minimum."""
    n = len(rods1)
    import cv2

    u1 = undistort_quirk(rods1, calibration["CM1"], calibration["dist1"])
    u2 = undistort_quirk(rods2, calibration["CM2"], calibration["dist2"])
    e1 = np.array([0, 0, 1, 1])
    e2 = np.array([0, 1, 0, 1])
    pu1, pu2 = u1[:, e1, :].reshape(-1, 2), u2[:, e2, :].reshape(-1, 2)
    po1, po2 = rods1[:, e1, :].reshape(-1, 2), rods2[:, e2, :].reshape(-1, 2)
    ph = cv2.triangulatePoints(P1, P2, pu1.T.copy(), pu2.T.copy())
    p = (ph[0:3] / ph[3]).T.copy()
    q1 = cv2.projectPoints(p, r1, t1, calibration["CM1"], calibration["dist1"])[0].reshape(-1, 2)
    q2 = cv2.projectPoints(p, r2, t2, calibration["CM2"], calibration["dist2"])[0].reshape(-1, 2)
    err = np.stack([np.linalg.norm(po1 - q1, axis=1), np.linalg.norm(po2 - q2, axis=1)], axis=1).mean(axis=1)
    pw = world_transform(p, rot, trans).reshape(n, 4, 3)
    e = err.reshape(-1, 4)
    s03, s12 = e[:, 0] + e[:, 3], e[:, 1] + e[:, 2]
    costs = np.minimum(s03, s12)
    rows = np.zeros((n, 18))
    for m in range(n):
        if s03[m] <= s12[m]:
            ends, c1, c2 = pw[m, 0::3], rods1[m], rods2[m]
        else:
            ends, c1, c2 = pw[m, 1:3], rods1[m, ::-1], rods2[m, ::-1]
        rows[m, 0:6] = ends.ravel()
        rows[m, 6:9] = ends.sum(axis=0) / 2
        rows[m, 9] = np.linalg.norm(np.diff(ends, axis=0))
        rows[m, 10:14] = c1.ravel()
        rows[m, 14:18] = c2.ravel()
    return {"rows": rows, "costs": costs}


def match_frame(data, cam1_name, cam2_name, frame, color, calibration, P1, P2, rot, trans, r1, r2, t1, t2,
                renumber=True):
    """match2D.py:741-998."""
    import pandas as pd

    if not renumber:
        # match2D.py:831-837,990-993: rows valid in both cameras, particle ids kept
        cols1 = [f"{k}_{cam1_name}" for k in ("x1", "y1", "x2", "y2")]
        cols2 = [f"{k}_{cam2_name}" for k in ("x1", "y1", "x2", "y2")]
        sel = data.loc[data.frame == frame]
        v1, v2 = sel[cols1].to_numpy(dtype=np.float64), sel[cols2].to_numpy(dtype=np.float64)
        ok = lambda v: ~np.isnan(v).all(axis=1) & (v != 0).any(axis=1)
        both = ok(v1) & ok(v2)
        if not both.any():
            return None
        res = match_arrays_keep_pairs(v1[both].reshape(-1, 2, 2), v2[both].reshape(-1, 2, 2),
                                      calibration, P1, P2, rot, trans, r1, r2, t1, t2)
        cols = ["x1", "y1", "z1", "x2", "y2", "z2", "x", "y", "z", "l", *cols1, *cols2]
        df = pd.DataFrame(res["rows"], columns=cols)
        df["frame"] = frame
        df["color"] = color
        df["particle"] = sel["particle"].to_numpy()[both] if "particle" in data.columns else list(range(len(df)))
        seen = [c for c in data.columns if "seen" in c]
        df[seen] = 1
        return df, res["costs"], res["rows"][:, 9]
    rods1 = select_rods(data, cam1_name, frame)
    rods2 = select_rods(data, cam2_name, frame)
    if len(rods1) == 0 or len(rods2) == 0:
        return None
    res = match_arrays(rods1, rods2, calibration, P1, P2, rot, trans, r1, r2, t1, t2)
    cols = ["x1", "y1", "z1", "x2", "y2", "z2", "x", "y", "z", "l"]
    cols += [f"{k}_{cam1_name}" for k in ("x1", "y1", "x2", "y2")]
    cols += [f"{k}_{cam2_name}" for k in ("x1", "y1", "x2", "y2")]
    df = pd.DataFrame(res["rows"], columns=cols)
    df["frame"] = frame
    df["color"] = color
    df["particle"] = list(range(len(df)))
    seen = [c for c in data.columns if "seen" in c]
    df[seen] = 1
    return df, res["costs"], res["rows"][:, 9]


def match_complex(data, frame_numbers, color, calibration, transform, cam1_name="gp1", cam2_name="gp2"):
    """match2D.py:634-738 with ``renumber=True`` and the rotation/translation
    transform format."""
    import pandas as pd
    from scipy.spatial.transform import Rotation

    from particletracking_b200.synthetic import projection_matrices

    P1, P2, r1, r2, t1, t2 = projection_matrices(calibration)
    rot = Rotation.from_matrix(transform["rotation"])
    trans = transform["translation"]
    dfs, errs, lens = [], [], []
    for f in frame_numbers:
        df, c, l = match_frame(data, cam1_name, cam2_name, f, color, calibration, P1, P2, rot, trans, r1, r2, t1, t2)
        dfs.append(df)
        errs.append(c)
        lens.append(l)
    out = pd.concat(dfs) if dfs else pd.DataFrame()
    out.reset_index(drop=True, inplace=True)
    return out, errs, lens


# --------------------------------------------------------------------------
# D  tracking_global_assignment
# --------------------------------------------------------------------------
def tracking_cost(p1: np.ndarray, p2: np.ndarray) -> np.ndarray:
    """tracking.py:100-121.  ``p1, p2 : [F, N, 3]`` end points 1 and 2 in
    DataFrame row order -> ``cost[F-1, N, N]`` with
    ``cost[f,a,b] = min(|p1[f+1,a]-p1[f,a]| + |p2[f+1,b]-p2[f,b]|,
    |p1[f,a]-p2[f+1,b]| + |p2[f,b]-p1[f+1,a]|)``."""
    F, N, _ = p1.shape
    p1s = np.broadcast_to(p1[:, :, None, :], (F, N, N, 3)).reshape(F, N * N, 3)
    p2s = np.broadcast_to(p2[:, None, :, :], (F, N, N, 3)).reshape(F, N * N, 3)
    d0 = np.linalg.norm(np.diff(p1s, axis=0), axis=2) + np.linalg.norm(np.diff(p2s, axis=0), axis=2)
    d1 = np.linalg.norm(p1s[0:-1] - p2s[1:], axis=2) + np.linalg.norm(p2s[0:-1] - p1s[1:], axis=2)
    return np.minimum(d0, d1).reshape(F - 1, N, N)


def tracking_arrays(p1: np.ndarray, p2: np.ndarray):
    """tracking.py:97-123,136-137 -> ``(cost, col_ind[F-1,N], total_cost[F-1])``."""
    from scipy.optimize import linear_sum_assignment

    cost = tracking_cost(p1, p2)
    col = np.stack([linear_sum_assignment(c)[1] for c in cost])
    total = np.take_along_axis(cost, col[:, :, None], axis=2)[:, :, 0].sum(axis=1)
    return cost, col, total


def chain_tracks(col_ind: np.ndarray) -> np.ndarray:
    """Chain-pass extension (no reference counterpart):
    ``ids[0] = arange(N)``, ``ids[f+1][col_ind[f][a]] = ids[f][a]``; entries outside ``0..N-1``
    (a failed assignment) break the chain: unreached positions get -1 and -1 propagates."""
    Fm1, N = col_ind.shape
    ids = np.full((Fm1 + 1, N), -1, dtype=np.int64)
    ids[0] = np.arange(N)
    for f in range(Fm1):
        ok = (col_ind[f] >= 0) & (col_ind[f] < N)
        ids[f + 1, col_ind[f][ok]] = ids[f][ok]
    return ids


def tracking_global_assignment(data):
    """tracking.py:70-138, including the positional ``particle`` rewrite
    (SURVEY.md App. D.2)."""
    frames = data["frame"].unique()
    F = len(frames)
    p1 = data[["x1", "y1", "z1"]].to_numpy().reshape((F, -1, 3))
    p2 = data[["x2", "y2", "z2"]].to_numpy().reshape((F, -1, 3))
    _, col, total = tracking_arrays(p1, p2)
    out = data.copy()
    if "particle" not in out.columns:
        out["particle"] = 0
        out.loc[out.frame == frames[0], ["particle"]] = np.arange(int((out.frame == frames[0]).sum()))
    pc = out.columns.get_loc("particle")
    for f, c in zip(frames[1:], col):
        tmp = out.loc[out.frame == f].copy()
        tmp.iloc[c, pc] = c
        out.loc[out.frame == f] = tmp
    out = out.sort_values(by=["frame", "particle"]).reset_index(drop=True)
    return out, total


# --------------------------------------------------------------------------
# E  matchND.create_weights
# --------------------------------------------------------------------------
def create_weights(p_3D: np.ndarray, p_3D_prev: np.ndarray, repr_errs: np.ndarray):
    """matchND.py:141-217 vectorised; note ``c2_re`` sums combo 1 only
    (the ``1:2`` slice, matchND.py:190)."""
    q1 = p_3D_prev[:, None, None, 0, :]
    q2 = p_3D_prev[:, None, None, 1, :]
    c1p1, c2p1, c2p2, c1p2 = (p_3D[None, :, :, k, :] for k in range(4))
    c1_re = repr_errs[:, :, 0::3, :].sum(axis=(2, 3))[None]
    c2_re = repr_errs[:, :, 1:2, :].sum(axis=(2, 3))[None]

    def nrm(x):
        return np.linalg.norm(x, axis=-1)

    s = np.stack(
        [
            (nrm(c1p1 - q1) + nrm(c1p2 - q2)) * c1_re,
            (nrm(c2p1 - q1) + nrm(c2p2 - q2)) * c2_re,
            (nrm(c2p1 - q2) + nrm(c2p2 - q1)) * c2_re,
            (nrm(c1p1 - q2) + nrm(c1p2 - q1)) * c1_re,
        ]
    )
    w = s.min(axis=0)
    choice = s.argmin(axis=0).astype(np.float64)
    return w.max() - w, choice


# --------------------------------------------------------------------------
# "next" rows: reorder_endpoints_csv, project_points / reproject_points
# --------------------------------------------------------------------------
def reorder_endpoints_arrays(ends: np.ndarray, pts2d: Optional[np.ndarray] = None):
    """match2D.py:1055-1137 on ``ends[F,N,2,3]`` (rows sorted by particle): frame after frame,
    swap the end points of a particle when the swapped displacement to the previous, already
    re-ordered frame is strictly smaller.  Returns ``(out_ends, out_2d, swapped[F,N],
    mincost[F-1,N])``; ``pts2d[F,N,2(cam),2(end),2]`` follows the swaps."""
    F, N = ends.shape[:2]
    out = ends.copy()
    o2 = None if pts2d is None else pts2d.copy()
    swapped = np.zeros((F, N), dtype=bool)
    mincost = np.zeros((max(F - 1, 0), N))
    for f in range(1, F):
        for n in range(N):
            a0, b0 = out[f - 1, n]
            a1, b1 = ends[f, n]
            comb1 = np.linalg.norm(a0 - a1) + np.linalg.norm(b0 - b1)
            comb2 = np.linalg.norm(b0 - a1) + np.linalg.norm(a0 - b1)
            mincost[f - 1, n] = min(comb1, comb2)
            if comb2 < comb1:
                out[f, n] = ends[f, n, ::-1]
                swapped[f, n] = True
                if o2 is not None:
                    o2[f, n] = pts2d[f, n, :, ::-1]
    return out, o2, swapped, mincost


def project_points(p_cam1, p_cam2, calibration, transforms=None):
    """calibrate_cameras.py:137-199."""
    import cv2
    from scipy.spatial.transform import Rotation

    from particletracking_b200.synthetic import projection_matrices

    P1, P2, *_ = projection_matrices(calibration)
    p = cv2.triangulatePoints(P1, P2, np.asarray(p_cam1, dtype=np.float64), np.asarray(p_cam2, dtype=np.float64))
    p = p[0:3] / p[3]
    if transforms is not None:
        p = (Rotation.from_matrix(transforms["rotation"]).apply(p.T) + transforms["translation"]).T
    return p


def reproject_points(points, calibration, transforms=None):
    """calibrate_cameras.py:202-262."""
    import cv2
    from scipy.spatial.transform import Rotation

    from particletracking_b200.synthetic import projection_matrices

    _, _, r1, r2, t1, t2 = projection_matrices(calibration)
    if transforms is not None:
        points = Rotation.from_matrix(transforms["rotation"]).inv().apply(points - transforms["translation"])
    a = cv2.projectPoints(points, r1, t1, calibration["CM1"], calibration["dist1"])[0].squeeze()
    b = cv2.projectPoints(points, r2, t2, calibration["CM2"], calibration["dist2"])[0].squeeze()
    return a, b


def reproject_data(data, cam_ids, calibration, position_scaling=1.0, transformation=None):
    """RodTracker/src/RodTracker/backend/reconstruction.py:373-469 (``Plotter.reproject_data``), restated on
    the ``reproject_points`` port above.  Parity unpinned for this wrapper: the module imports PyQt5, which is
    not installed; its arithmetic is ``reproject_points`` (pinned by tests/golden/project.npz)."""
    if calibration is None:
        return None
    id1, id2 = cam_ids
    cols = ["x1", "y1", "z1", "x2", "y2", "z2", "x", "y", "z", "l",
            f"x1_{id1}", f"y1_{id1}", f"x2_{id1}", f"y2_{id1}", f"x1_{id2}", f"y1_{id2}", f"x2_{id2}", f"y2_{id2}"]
    d = data[cols]
    e1_3d = d.iloc[:, 0:3].to_numpy(dtype=float) * position_scaling
    e2_3d = d.iloc[:, 3:6].to_numpy(dtype=float) * position_scaling
    e1_2d_c1 = d.iloc[:, 10:12].to_numpy(dtype=float) * position_scaling
    e2_2d_c1 = d.iloc[:, 12:14].to_numpy(dtype=float) * position_scaling
    e1_2d_c2 = d.iloc[:, 14:16].to_numpy(dtype=float) * position_scaling
    e2_2d_c2 = d.iloc[:, 16:18].to_numpy(dtype=float) * position_scaling
    drop1 = np.isnan(e1_3d).all(axis=1)
    drop2 = np.isnan(e2_3d).all(axis=1)
    e1_3d, e1_2d_c1, e1_2d_c2 = e1_3d[~drop1], e1_2d_c1[~drop1], e1_2d_c2[~drop1]
    e2_3d, e2_2d_c1, e2_2d_c2 = e2_3d[~drop2], e2_2d_c1[~drop2], e2_2d_c2[~drop2]
    if not len(e1_3d) or not len(e2_3d):
        return None
    a1, a2 = reproject_points(e1_3d, calibration, transformation)
    b1, b2 = reproject_points(e2_3d, calibration, transformation)
    errs = [np.abs(a1 - e1_2d_c1), np.abs(a2 - e1_2d_c2), np.abs(b1 - e2_2d_c1), np.abs(b2 - e2_2d_c2)]
    return np.sum(np.asarray(errs), axis=-1)
