"""Synthetic calibrated stereo rig + rod data.

Everything the benchmark and the parity tests run on comes from here: a
parametric two-camera rig that resembles the reference's example calibration
(``ParticleDetection/tests/example_data/gp34.json``) without copying it, and
ballistic rods in a box that are projected into both cameras with the
pinhole + radial/tangential model that ``cv2.projectPoints`` implements
(SURVEY.md App. B.3).

Two forms of the same data are produced:

* tensor form  ``cam1, cam2 : float64 [F, C, Nmax, 2(end), 2(xy)]`` with
  ``n1, n2 : int32 [F, C]`` -- what the CUDA path consumes;
* DataFrame/CSV form, one frame table per colour with the reference's column
  schema (``utils/datasets.py:53-76`` + ``particle``), 3D columns NaN.

Pure numpy; no cv2, no torch, so it is usable on either side of the oracle
boundary.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence

import numpy as np

COLORS = ["black", "blue", "green", "red", "yellow", "brown", "lilac", "orange"]


def synthetic_calibration() -> Dict[str, np.ndarray]:
    """Stereo calibration dict with the keys ``load_camera_calibration`` yields
    (reference ``utils/data_loading.py:323-346``): CM1, dist1, CM2, dist2, R, T."""
    a = np.deg2rad(89.25)
    c, s = np.cos(a), np.sin(a)
    return {
        "CM1": np.array([[2640.0, 0.0, 653.0], [0.0, 2626.0, 487.0], [0.0, 0.0, 1.0]]),
        "dist1": np.array([-0.11, 1.2, 0.0, 0.0]),
        "CM2": np.array([[3398.0, 0.0, 667.0], [0.0, 3365.0, 465.0], [0.0, 0.0, 1.0]]),
        "dist2": np.array([-0.41, 10.1, 0.0, 0.0]),
        "R": np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]]),
        "T": np.array([[-2.5], [285.0], [350.0]]),
    }


def synthetic_transformation() -> Dict[str, np.ndarray]:
    """World transform dict as ``load_world_transformation`` yields it
    (reference ``utils/data_loading.py:263-320``)."""
    return {
        "rotation": np.diag([1.0, -1.0, -1.0]),
        "translation": np.array([2.0, 2.5, 286.0]),
    }


from .rig import projection_matrices  # noqa: E402,F401  (lives with the rig; re-exported for the tests / tools)


def project(X: np.ndarray, Rm: np.ndarray, t: np.ndarray, CM: np.ndarray, dist: np.ndarray) -> np.ndarray:
    """Pinhole + distortion projection of ``X[...,3]``."""
    k = np.zeros(14)
    k[: len(np.ravel(dist))] = np.ravel(dist)
    Y = X @ Rm.T + np.ravel(t)
    x = Y[..., 0] / Y[..., 2]
    y = Y[..., 1] / Y[..., 2]
    r2 = x * x + y * y
    r4 = r2 * r2
    r6 = r4 * r2
    cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6
    icdist2 = 1.0 / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6)
    a1 = 2 * x * y
    a2 = r2 + 2 * x * x
    a3 = r2 + 2 * y * y
    xd = x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4
    yd = y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4
    return np.stack([xd * CM[0, 0] + CM[0, 2], yd * CM[1, 1] + CM[1, 2]], axis=-1)


@dataclasses.dataclass
class SyntheticStereo:
    """One synthetic experiment in tensor form (+ ground truth)."""

    cam1: np.ndarray  # [F, C, Nmax, 2, 2] float64; rows >= n1 are NaN
    cam2: np.ndarray  # [F, C, Nmax, 2, 2]
    n1: np.ndarray  # [F, C] int32
    n2: np.ndarray  # [F, C] int32
    frames: np.ndarray  # [F] int64 frame numbers
    colors: List[str]
    truth_perm: np.ndarray  # [F, C, N]: cam2 row r shows rod truth_perm[f,c,r]
    world_ends: np.ndarray  # [F, C, N, 2, 3] true 3D end points (world)
    calibration: Dict[str, np.ndarray]
    transformation: Dict[str, np.ndarray]

    def problems(self):
        """Flatten (frame, colour) -> P problems: cam1[P,Nmax,2,2], cam2, n1[P], n2[P]."""
        F, C, Nm = self.cam1.shape[:3]
        return (
            self.cam1.reshape(F * C, Nm, 2, 2),
            self.cam2.reshape(F * C, Nm, 2, 2),
            self.n1.reshape(F * C),
            self.n2.reshape(F * C),
        )


def make_stereo(
    n_frames: int,
    n_colors: int,
    n_rods: int,
    seed: int,
    sigma_px: float = 0.25,
    p_drop: float = 0.0,
    p_dummy: float = 0.0,
    first_frame: int = 0,
    rod_length: float = 10.0,
    calibration: Optional[Dict[str, np.ndarray]] = None,
    transformation: Optional[Dict[str, np.ndarray]] = None,
) -> SyntheticStereo:
    """Generate a synthetic stereo experiment.

    ``p_drop``: each detection is dropped per camera independently; the
    remaining rows are compacted (this is what the reference's ``dropna`` does
    to all-NaN rows, ``match2D.py:827-830``) so N1 != N2 arises.
    ``p_dummy``: a detection is replaced by the detector's dummy rod
    ``(-1,-1,-1,-1)`` which the reference *keeps* (SURVEY.md App. A.1).
    """
    rng = np.random.default_rng(seed)
    cal = calibration or synthetic_calibration()
    trf = transformation or synthetic_transformation()
    F, C, N = n_frames, n_colors, n_rods
    Rw = np.asarray(trf["rotation"], dtype=np.float64)
    tw = np.asarray(trf["translation"], dtype=np.float64)

    box = np.array([50.0, 32.0, 32.0])
    centre0 = rng.uniform(-box, box, size=(C, N, 3))
    d = rng.normal(size=(C, N, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    vel = rng.normal(0.0, 0.2, size=(C, N, 3))
    t = np.arange(F, dtype=np.float64)[:, None, None, None]
    # ballistic flight with elastic reflection at the box walls (triangle-wave folding), so
    # that arbitrarily long sequences stay inside both cameras' fields of view
    centre = _fold(centre0[None] + vel[None] * t, box)  # [F,C,N,3]
    half = 0.5 * rod_length * d[None]
    ends_w = np.stack([centre - half, centre + half], axis=3)  # [F,C,N,2,3]

    # world -> camera-1 coordinates: X_c = Rw^T (X_w - t_w)
    ends_c = (ends_w - tw) @ Rw
    P1, P2, r1, r2, t1, t2 = projection_matrices(cal)
    uv1 = project(ends_c, r1, t1, cal["CM1"], cal["dist1"])  # [F,C,N,2,2]
    uv2 = project(ends_c, r2, t2, cal["CM2"], cal["dist2"])
    uv1 = uv1 + rng.normal(0.0, sigma_px, size=uv1.shape)
    uv2 = uv2 + rng.normal(0.0, sigma_px, size=uv2.shape)

    # scramble camera 2: permute rods per (frame, colour), swap end points w.p. 0.5
    perm = np.argsort(rng.random((F, C, N)), axis=-1)
    uv2 = np.take_along_axis(uv2, perm[..., None, None], axis=2)
    swap = rng.random((F, C, N)) < 0.5
    uv2 = np.where(swap[..., None, None], uv2[..., ::-1, :], uv2)

    if p_dummy > 0.0:
        dm1 = rng.random((F, C, N)) < p_dummy
        dm2 = rng.random((F, C, N)) < p_dummy
        uv1 = np.where(dm1[..., None, None], -1.0, uv1)
        uv2 = np.where(dm2[..., None, None], -1.0, uv2)

    n1 = np.full((F, C), N, dtype=np.int32)
    n2 = np.full((F, C), N, dtype=np.int32)
    if p_drop > 0.0:
        keep1 = rng.random((F, C, N)) >= p_drop
        keep2 = rng.random((F, C, N)) >= p_drop
        # never drop everything: the reference returns None for an empty camera
        keep1[..., 0] = True
        keep2[..., 0] = True
        uv1, n1 = _compact(uv1, keep1)
        uv2, n2 = _compact(uv2, keep2)

    return SyntheticStereo(
        cam1=np.ascontiguousarray(uv1),
        cam2=np.ascontiguousarray(uv2),
        n1=n1,
        n2=n2,
        frames=np.arange(first_frame, first_frame + F, dtype=np.int64),
        colors=[COLORS[c % len(COLORS)] + ("" if c < len(COLORS) else str(c)) for c in range(C)],
        truth_perm=perm,
        world_ends=ends_w,
        calibration=cal,
        transformation=trf,
    )


def _fold(x: np.ndarray, half: np.ndarray) -> np.ndarray:
    """Reflect coordinates into [-half, half] (period 4*half triangle wave)."""
    y = np.mod(x + half, 4.0 * half)
    y = np.where(y > 2.0 * half, 4.0 * half - y, y)
    return y - half


def _compact(uv: np.ndarray, keep: np.ndarray):
    """Stable-compact kept rows to the front of axis 2, NaN-pad the rest."""
    F, C, N = keep.shape
    order = np.argsort(~keep, axis=-1, kind="stable")
    out = np.take_along_axis(uv, order[..., None, None], axis=2)
    cnt = keep.sum(axis=-1).astype(np.int32)
    pad = np.arange(N)[None, None, :] >= cnt[..., None]
    out = np.where(pad[..., None, None], np.nan, out)
    return out, cnt


def to_dataframes(data: SyntheticStereo, cam1_name: str = "gp1", cam2_name: str = "gp2"):
    """DataFrame form: ``{color: DataFrame}`` in the reference's CSV schema
    (SURVEY.md App. A.8); unseen detections are all-NaN camera columns."""
    import pandas as pd

    F, C, Nm = data.cam1.shape[:3]
    out = {}
    cols3d = ["x1", "y1", "z1", "x2", "y2", "z2", "x", "y", "z", "l"]
    c1 = [f"x1_{cam1_name}", f"y1_{cam1_name}", f"x2_{cam1_name}", f"y2_{cam1_name}"]
    c2 = [f"x1_{cam2_name}", f"y1_{cam2_name}", f"x2_{cam2_name}", f"y2_{cam2_name}"]
    for ci, color in enumerate(data.colors):
        df = pd.DataFrame(np.full((F * Nm, 10), np.nan), columns=cols3d)
        df[c1] = data.cam1[:, ci].reshape(F * Nm, 4)
        df[c2] = data.cam2[:, ci].reshape(F * Nm, 4)
        df["frame"] = np.repeat(data.frames, Nm)
        df[f"seen_{cam1_name}"] = 1
        df[f"seen_{cam2_name}"] = 1
        df["color"] = color
        df["particle"] = np.tile(np.arange(Nm), F)
        out[color] = df
    return out


def write_csvs(data: SyntheticStereo, folder: str, cam1_name: str = "gp1", cam2_name: str = "gp2") -> Sequence[str]:
    """Write ``rods_df_{color}.csv`` files the way the reference reads them
    (``match2D.py:601-602``: ``index_col=0``)."""
    import os

    os.makedirs(folder, exist_ok=True)
    paths = []
    for color, df in to_dataframes(data, cam1_name, cam2_name).items():
        p = os.path.join(folder, f"rods_df_{color}.csv")
        df.to_csv(p, sep=",")
        paths.append(p)
    return paths
