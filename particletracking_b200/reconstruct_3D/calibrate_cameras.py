"""``project_points`` / ``reproject_points`` of the reference's
``reconstruct_3D/calibrate_cameras.py`` (:137-199, :202-262) on the B200 -- the DLT and the
projection of the hot path exposed stand-alone.  Chessboard calibration
(``stereo_calibrate``) is a one-off offline step and out of scope."""
from __future__ import annotations

from typing import Tuple, Union

import numpy as np
import torch

from .. import engine
from ..rig import make_rig
from ..rig import projection_matrices


def _rig(calibration: dict, transforms):
    P1, P2, r1, r2, t1, t2 = projection_matrices(calibration)
    if transforms is not None:
        from scipy.spatial.transform import Rotation

        rot, trans = Rotation.from_matrix(transforms["rotation"]), transforms["translation"]
    else:
        rot, trans = np.eye(3), np.zeros(3)
    cal = dict(calibration)
    cal.setdefault("dist1", np.zeros(4))
    cal.setdefault("dist2", np.zeros(4))
    return make_rig(cal, P1, P2, rot, trans, r1, r2, t1, t2)


def project_points(p_cam1: np.ndarray, p_cam2: np.ndarray, calibration: dict, transforms: Union[dict, None] = None):
    """Project points from a stereocamera system to 3D coordinates: ``p_cam1, p_cam2`` of shape
    ``(2, n)`` -> ``(3, n)`` in world coordinates if ``transforms`` is given, else in camera-1
    coordinates."""
    engine._require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).T)).to(dev)
    out = engine.triangulate_points(_rig(calibration, transforms), t(p_cam1), t(p_cam2), to_world=transforms is not None)
    return out.cpu().numpy().T.copy()


def reproject_points(points: np.ndarray, calibration: dict, transforms: Union[dict, None] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Project 3D coordinates (``(n, 3)``, or ``(3, n)`` without ``transforms``) to the image
    planes of camera 1 and 2."""
    engine._require_cuda()
    pts = np.asarray(points, dtype=np.float64)
    if transforms is None and pts.ndim == 2 and pts.shape[0] == 3 and pts.shape[1] != 3:
        pts = pts.T
    if pts.ndim != 2 or pts.shape[1] != 3:
        raise ValueError("points must have shape (n, 3)")
    dev = torch.device("cuda", torch.cuda.current_device())
    uv1, uv2 = engine.project_points(_rig(calibration, transforms), torch.from_numpy(np.ascontiguousarray(pts)).to(dev),
                                     from_world=transforms is not None)
    return uv1.cpu().numpy().squeeze(), uv2.cpu().numpy().squeeze()
