"""Generates tests/golden/*.npz by running the UNMODIFIED reference from
/root/reference (build container only; `python tests/golden/make_golden.py`).

The reference's own tests hold no numeric goldens for this path, so the
known answers are produced here: seeded synthetic inputs (particletracking_b200.synthetic,
SURVEY.md App. F) go through the reference's match2D.match_frame / match_complex,
tracking.tracking_global_assignment and matchND.create_weights, and inputs + outputs are
stored.  scipy's linear_sum_assignment is wrapped (not replaced) inside the reference
modules while they run, to also record the cost matrices and assignments the reference
saw -- those never leave match_frame otherwise.

Versions at generation time are stored in each file (`versions`).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402
from particletracking_b200 import synthetic as syn  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def versions():
    import cv2
    import pandas
    import scipy

    return np.array([f"cv2={cv2.__version__}", f"scipy={scipy.__version__}",
                     f"numpy={np.__version__}", f"pandas={pandas.__version__}"])


class Spy:
    """Wraps linear_sum_assignment inside a reference module and records calls."""

    def __init__(self, module):
        self.module = module
        self.orig = module.linear_sum_assignment
        self.calls = []

    def __enter__(self):
        def wrapped(cost, *a, **k):
            r = self.orig(cost, *a, **k)
            self.calls.append((np.array(cost, copy=True), np.array(r[0]), np.array(r[1])))
            return r

        self.module.linear_sum_assignment = wrapped
        return self

    def __exit__(self, *exc):
        self.module.linear_sum_assignment = self.orig


def rig_args():
    from scipy.spatial.transform import Rotation

    cal = syn.synthetic_calibration()
    trf = syn.synthetic_transformation()
    P1, P2, r1, r2, t1, t2 = syn.projection_matrices(cal)
    return cal, trf, (cal, P1, P2, Rotation.from_matrix(trf["rotation"]), trf["translation"], r1, r2, t1, t2)


def golden_match(name, F, C, N, seed, **kw):
    m2d = refload.load()[0]
    cal, trf, args = rig_args()
    data = syn.make_stereo(F, C, N, seed=seed, **kw)
    dfs = syn.to_dataframes(data)
    rows = np.full((F, C, N, 18), np.nan)
    costs = np.full((F, C, N), np.nan)
    lengths = np.full((F, C, N), np.nan)
    cost_mat = np.full((F, C, N, N), np.nan)
    row_ind = np.full((F, C, N), -1, dtype=np.int32)
    col_ind = np.full((F, C, N), -1, dtype=np.int32)
    n_out = np.zeros((F, C), dtype=np.int32)
    for ci, color in enumerate(data.colors):
        for fi, f in enumerate(data.frames):
            with Spy(m2d) as spy:
                df, c, l = m2d.match_frame(dfs[color], "gp1", "gp2", int(f), color, *args, renumber=True)
            n = len(df)
            n_out[fi, ci] = n
            rows[fi, ci, :n] = df.iloc[:, :18].to_numpy()
            costs[fi, ci, :n] = c
            lengths[fi, ci, :n] = l
            cm, ri, cj = spy.calls[0]
            cost_mat[fi, ci, : cm.shape[0], : cm.shape[1]] = cm
            row_ind[fi, ci, :n] = ri
            col_ind[fi, ci, :n] = cj
            assert list(df["particle"]) == list(range(n)) and (df["frame"] == f).all()
    np.savez_compressed(
        os.path.join(OUT, name), cam1=data.cam1, cam2=data.cam2, n1=data.n1, n2=data.n2,
        rows=rows, costs=costs, lengths=lengths, cost=cost_mat, row_ind=row_ind, col_ind=col_ind,
        n_out=n_out, versions=versions(),
        params=np.array([F, C, N, seed]), kw=np.array([f"{k}={v}" for k, v in kw.items()]))
    print(name, "ok", n_out.min(), n_out.max())
    return data, dfs


def golden_tracking(name, F, N, seed):
    m2d, _, trk, _ = refload.load()
    cal, trf, args = rig_args()
    data = syn.make_stereo(F, 1, N, seed=seed)
    df = syn.to_dataframes(data)["black"]
    matched = m2d.match_complex(df, list(data.frames), "black", cal, trf)[0]
    with Spy(trk) as spy:
        out, total = trk.tracking_global_assignment(matched)
    ends = matched[["x1", "y1", "z1", "x2", "y2", "z2"]].to_numpy().reshape(F, N, 2, 3)
    np.savez_compressed(
        os.path.join(OUT, name), ends=ends,
        cost=np.stack([c[0] for c in spy.calls]), col_ind=np.stack([c[2] for c in spy.calls]).astype(np.int32),
        total_cost=total, particle_in=matched["particle"].to_numpy(), frame_in=matched["frame"].to_numpy(),
        particle_out=out["particle"].to_numpy(), frame_out=out["frame"].to_numpy(),
        x1_out=out["x1"].to_numpy(), versions=versions(), params=np.array([F, N, seed]))
    print(name, "ok", total[:3])


def golden_weights(name, seed):
    mnd = refload.load()[1]
    rng = np.random.default_rng(seed)
    p3 = rng.normal(size=(6, 7, 4, 3)) * 25
    prev = rng.normal(size=(5, 2, 3)) * 25
    re = rng.random((6, 7, 4, 2)) * 3
    w, ch = mnd.create_weights(p3, prev, re)
    np.savez_compressed(os.path.join(OUT, name), p_3D=p3, p_prev=prev, rep_errs=re,
                        weights=w, choices=ch, versions=versions())
    print(name, "ok")


def golden_reorder(name, F, N, seed):
    """reference match2D.reorder_endpoints_csv on a 3D-matched experiment with random end-point flips."""
    import tempfile

    import pandas as pd

    m2d = refload.load()[0]
    cal, trf, args = rig_args()
    data = syn.make_stereo(F, 1, N, seed=seed)
    df = syn.to_dataframes(data)["black"]
    m = m2d.match_complex(df, list(data.frames), "black", cal, trf)[0]
    m["particle"] = np.concatenate([np.arange(N) for _ in range(F)])
    flip = np.random.default_rng(seed).random(len(m)) < 0.4
    a = ["x1", "y1", "z1", "x1_gp1", "y1_gp1", "x1_gp2", "y1_gp2"]
    b = ["x2", "y2", "z2", "x2_gp1", "y2_gp1", "x2_gp2", "y2_gp2"]
    va, vb = m.loc[flip, a].to_numpy(), m.loc[flip, b].to_numpy()
    m.loc[flip, a] = vb
    m.loc[flip, b] = va
    c2d = [f"{k}_{c}" for c in ("gp1", "gp2") for k in ("x1", "y1", "x2", "y2")]
    c3d = ["x1", "y1", "z1", "x2", "y2", "z2"]
    with tempfile.TemporaryDirectory() as tmp:
        os.mkdir(os.path.join(tmp, "in"))
        m.to_csv(os.path.join(tmp, "in", "rods_df_black.csv"), sep=",")
        src = pd.read_csv(os.path.join(tmp, "in", "rods_df_black.csv"), index_col=0)
        mc = m2d.reorder_endpoints_csv(os.path.join(tmp, "in"), os.path.join(tmp, "out"), ["black"], "gp1", "gp2", list(range(F)))
        out = pd.read_csv(os.path.join(tmp, "out", "rods_df_black.csv"), index_col=0, float_precision="round_trip")
    np.savez_compressed(
        os.path.join(OUT, name), ends=src[c3d].to_numpy().reshape(F, N, 2, 3), pts2d=src[c2d].to_numpy().reshape(F, N, 2, 2, 2),
        out_ends=out[c3d].to_numpy().reshape(F, N, 2, 3), out_2d=out[c2d].to_numpy().reshape(F, N, 2, 2, 2),
        mincosts_T=mc, versions=versions(), params=np.array([F, N, seed]))
    print(name, "ok", int((out[c3d].to_numpy() != src[c3d].to_numpy()).any(axis=1).sum()), "rows swapped")


def golden_project(name, seed):
    """reference calibrate_cameras.project_points / reproject_points (needs a matplotlib stub to import)."""
    import types

    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    refload.load()
    import ParticleDetection.reconstruct_3D.calibrate_cameras as cc

    cal, trf, _ = rig_args()
    X = np.random.default_rng(seed).uniform([-40, -30, 260], [40, 30, 320], size=(64, 3))
    a, b = cc.reproject_points(X, cal, None)
    aw, bw = cc.reproject_points(X, cal, trf)
    np.savez_compressed(os.path.join(OUT, name), X=X, uv1=a, uv2=b, uv1_w=aw, uv2_w=bw,
                        p3d=cc.project_points(a.T.copy(), b.T.copy(), cal, None),
                        p3d_w=cc.project_points(a.T.copy(), b.T.copy(), cal, trf), versions=versions())
    print(name, "ok")


def golden_matchnd(name, F, N, seed, **kw):
    """reference matchND.match_frame / assign chain with ``npartite_matching`` replaced by the HiGHS
    restatement of the same programme (oracle/npartite.py; pulp and CBC are absent here).  Everything
    around the solver -- triangulation block, create_weights, index bookkeeping, rows -- is the
    reference's own code.  Stored: input frames, every frame's weights / solution / output rows."""
    from oracle import npartite as onp

    m2d, mnd = refload.load()[:2]
    cal, trf, args = rig_args()
    data = syn.make_stereo(F, 1, N, seed=seed, **kw)
    color = data.colors[0]
    df = syn.to_dataframes(data)[color]
    seen = []
    orig_solver = mnd.npartite_matching

    def solver(weights, maximize=True, solver=None):
        sol = onp.npartite_matching(weights, maximize)
        seen.append((np.array(weights, copy=True), np.stack(sol)))
        return sol

    mnd.npartite_matching = solver
    frames = [int(f) for f in data.frames]
    outs, costs, lens = [], [], []
    tmp = None
    for fn, frame in enumerate(frames):  # matchND.py:318-360
        if fn == 0:
            tmp, c, l = m2d.match_frame(df, "gp1", "gp2", frame, color, *args, renumber=True)
        else:
            tmp, c, l = mnd.match_frame(df, tmp, "gp1", "gp2", frame, color, *args)
        outs.append(tmp[[*m2d_cols(m2d)]].to_numpy())
        costs.append(c)
        lens.append(l)
    mnd.npartite_matching = orig_solver
    np.savez_compressed(
        os.path.join(OUT, name), cam1=data.cam1[:, 0], cam2=data.cam2[:, 0], n1=data.n1[:, 0], n2=data.n2[:, 0],
        frames=np.array(frames), rows=np.stack(outs), costs=np.stack(costs), lens=np.stack(lens),
        weights=np.stack([w for w, _ in seen]), sols=np.stack([s_ for _, s_ in seen]),
        values=np.array([onp.objective(w, s_) for w, s_ in seen]), versions=versions(), params=np.array([F, N, seed]))
    print(name, "ok", [float(v) for v in np.array([onp.objective(w, s_) for w, s_ in seen])[:2]])


def m2d_cols(m2d):
    return ["x1", "y1", "z1", "x2", "y2", "z2", "x", "y", "z", "l", "x1_gp1", "y1_gp1", "x2_gp1", "y2_gp1",
            "x1_gp2", "y1_gp2", "x2_gp2", "y2_gp2"]


if __name__ == "__main__":
    assert refload.available(), "needs /root/reference"
    # C1 of BASELINE.json: one stereo frame pair, 3 colours x 12 rods (+ one more frame)
    golden_match("c1_3x12.npz", 2, 3, 12, seed=1)
    # ragged: dropped detections (N1 != N2, both orientations) and dummy -1 rods
    golden_match("ragged_2x16.npz", 5, 2, 16, seed=3, p_drop=0.2, p_dummy=0.1, sigma_px=0.5)
    # tiny problems: the N == 1 undistortion special case and N == 2
    golden_match("tiny_n1.npz", 3, 2, 1, seed=11)
    golden_match("tiny_n2.npz", 3, 2, 2, seed=12)
    golden_match("n32.npz", 1, 2, 32, seed=2)
    golden_tracking("track_12.npz", 6, 12, seed=21)
    golden_tracking("track_25.npz", 4, 25, seed=22)
    golden_weights("weights.npz", seed=5)
    golden_reorder("reorder.npz", 9, 10, seed=41)
    golden_project("project.npz", seed=6)
    golden_matchnd("matchnd_12.npz", 4, 12, seed=51)
    golden_matchnd("matchnd_25.npz", 3, 25, seed=52, sigma_px=0.5)
