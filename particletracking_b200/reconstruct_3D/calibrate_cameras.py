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


def reproject_data(data, cam_ids, calibration: dict, position_scaling: float = 1.0, transformation: Union[dict, None] = None):
    """Absolute reprojection errors of every row of a rod DataFrame -- RodTracker's
    ``Plotter.reproject_data`` (``RodTracker/backend/reconstruction.py:373-469``, the numeric body behind
    its reprojection-error histogram; the Qt plumbing around it is out of scope): both 3D end points are
    projected into both cameras (K9, one launch for all of them) and compared with the stored 2D end points.
    Returns ``None`` without a calibration or without reconstructable rows, else an array ``[4, n]``:
    end 1 in camera 1, end 1 in camera 2, end 2 in camera 1, end 2 in camera 2 (sum of |dx| + |dy|)."""
    if calibration is None:
        return None
    id1, id2 = cam_ids[0], cam_ids[1]
    cols = ["x1", "y1", "z1", "x2", "y2", "z2", "x", "y", "z", "l",
            f"x1_{id1:s}", f"y1_{id1:s}", f"x2_{id1:s}", f"y2_{id1:s}",
            f"x1_{id2:s}", f"y1_{id2:s}", f"x2_{id2:s}", f"y2_{id2:s}"]
    a = data[cols].to_numpy(dtype=float) * position_scaling
    e1_3d, e2_3d = a[:, 0:3], a[:, 3:6]
    keep1 = ~np.isnan(e1_3d).all(axis=1)
    keep2 = ~np.isnan(e2_3d).all(axis=1)
    e1_3d, e1_c1, e1_c2 = e1_3d[keep1], a[keep1, 10:12], a[keep1, 14:16]
    e2_3d, e2_c1, e2_c2 = e2_3d[keep2], a[keep2, 12:14], a[keep2, 16:18]
    if not len(e1_3d) or not len(e2_3d):
        return None
    n1 = len(e1_3d)
    uv1, uv2 = reproject_points(np.concatenate([e1_3d, e2_3d]), calibration, transformation)
    uv1, uv2 = np.atleast_2d(uv1), np.atleast_2d(uv2)
    errs = [np.abs(uv1[:n1] - e1_c1), np.abs(uv2[:n1] - e1_c2), np.abs(uv1[n1:] - e2_c1), np.abs(uv2[n1:] - e2_c2)]
    return np.sum(np.asarray(errs), axis=-1)
