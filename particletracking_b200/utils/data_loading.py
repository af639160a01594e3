"""Rig input formats (boundary only).

Same two loaders, same file formats and return structures as the reference's
``ParticleDetection/utils/data_loading.py`` (``load_camera_calibration`` :323-346,
``load_world_transformation`` :263-320)."""
from __future__ import annotations

import json

import numpy as np


def load_camera_calibration(file_name: str) -> dict:
    """``*.json`` stereo calibration -> dict of ndarrays (every list-valued entry;
    CM1, dist1, CM2, dist2, R, T, E, F in the reference's files)."""
    with open(file_name, "r") as f:
        raw = json.load(f)
    return {k: np.asarray(v) for k, v in raw.items() if isinstance(v, list)}


def load_world_transformation(file_name: str) -> dict:
    """``*.json`` camera-1 -> world transform as ``{"rotation": (3,3), "translation": (3,)}``.
    Accepts the new format (those two keys) and the legacy five-matrix format nested under
    ``"transformations"``, which is folded into one rotation + translation."""
    with open(file_name, "r") as f:
        raw = json.load(f)
    if "transformations" in raw:
        from scipy.spatial.transform import Rotation

        t = raw["transformations"]
        rx = Rotation.from_matrix(np.asarray(t["M_rotate_x"])[0:3, 0:3])
        ry = Rotation.from_matrix(np.asarray(t["M_rotate_y"])[0:3, 0:3])
        rz = Rotation.from_matrix(np.asarray(t["M_rotate_z"])[0:3, 0:3])
        rot = rz * ry * rx
        trans = rot.apply(np.asarray(t["M_trans"])[0:3, 3]) + np.asarray(t["M_trans2"])[0:3, 3]
        return {"rotation": rot.as_matrix(), "translation": trans}
    if set(raw.keys()) == {"rotation", "translation"}:
        return {"rotation": np.array(raw["rotation"]), "translation": np.array(raw["translation"])}
    raise ValueError(f"Incompatible structure of the given file: {file_name}")
