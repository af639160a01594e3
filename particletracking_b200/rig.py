"""The stereo rig as the C ABI sees it (``ptk_rig`` in include/ptk.h).

``make_rig`` takes exactly what the reference's ``match_frame`` receives
(``reconstruct_3D/match2D.py:741-757``): the calibration dict, P1, P2, the world
rotation / translation and the redundant ``r1, r2, t1, t2`` -- and honours them
as passed.
"""
from __future__ import annotations

import ctypes
from typing import Dict

import numpy as np


class PtkRig(ctypes.Structure):
    _fields_ = [
        ("CM1", ctypes.c_double * 9),
        ("CM2", ctypes.c_double * 9),
        ("dist1", ctypes.c_double * 14),
        ("dist2", ctypes.c_double * 14),
        ("R1", ctypes.c_double * 9),
        ("t1", ctypes.c_double * 3),
        ("R2", ctypes.c_double * 9),
        ("t2", ctypes.c_double * 3),
        ("P1", ctypes.c_double * 12),
        ("P2", ctypes.c_double * 12),
        ("Rw", ctypes.c_double * 9),
        ("tw", ctypes.c_double * 3),
        ("agg_sum", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


def _fill(dst, src, n, name):
    a = np.asarray(src, dtype=np.float64).ravel()
    if a.size > n:
        raise ValueError(f"{name}: expected at most {n} values, got {a.size}")
    for i in range(n):
        dst[i] = float(a[i]) if i < a.size else 0.0


def _rotation_matrix(r, name):
    """cv2.projectPoints accepts a 3x3 matrix or a Rodrigues 3-vector as rvec; the
    reference always passes matrices (match2D.py:584-590).  Vectors are converted."""
    a = np.asarray(r, dtype=np.float64)
    if a.size == 9:
        return a.reshape(3, 3)
    if a.size == 3:
        from scipy.spatial.transform import Rotation

        return Rotation.from_rotvec(a.ravel()).as_matrix()
    raise ValueError(f"{name}: expected a 3x3 matrix or a rotation vector")


def projection_matrices(calibration: Dict[str, np.ndarray]):
    """P1 = CM1 [I|0], P2 = CM2 [R|T] and the (r, t) pairs the reference's
    drivers derive (``match2D.py:584-592``)."""
    r1 = np.eye(3)
    t1 = np.zeros((3, 1))
    r2 = np.asarray(calibration["R"], dtype=np.float64)
    t2 = np.asarray(calibration["T"], dtype=np.float64).reshape(3, 1)
    P1 = np.asarray(calibration["CM1"], dtype=np.float64) @ np.hstack([r1, t1])
    P2 = np.asarray(calibration["CM2"], dtype=np.float64) @ np.hstack([r2, t2])
    return P1, P2, r1, r2, t1, t2


def make_rig(calibration: Dict[str, np.ndarray], P1, P2, rot, trans, r1, r2, t1, t2,
             agg_sum: bool = False) -> PtkRig:
    rig = PtkRig()
    _fill(rig.CM1, calibration["CM1"], 9, "CM1")
    _fill(rig.CM2, calibration["CM2"], 9, "CM2")
    if np.asarray(calibration["CM1"]).size != 9 or np.asarray(calibration["CM2"]).size != 9:
        raise ValueError("camera matrices must be 3x3")
    _fill(rig.dist1, calibration["dist1"], 14, "dist1")
    _fill(rig.dist2, calibration["dist2"], 14, "dist2")
    if rig.dist1[12] != 0.0 or rig.dist1[13] != 0.0 or rig.dist2[12] != 0.0 or rig.dist2[13] != 0.0:
        # cv2's tilted-sensor terms (tauX, tauY) are not implemented by the kernels; computing with a
        # different camera model than the reference's cv2 calls would be silent garbage
        raise ValueError("distortion coefficients 13 and 14 (tauX, tauY: tilted sensor model) are not supported")
    _fill(rig.R1, _rotation_matrix(r1, "r1"), 9, "r1")
    _fill(rig.R2, _rotation_matrix(r2, "r2"), 9, "r2")
    _fill(rig.t1, t1, 3, "t1")
    _fill(rig.t2, t2, 3, "t2")
    if np.asarray(P1).shape != (3, 4) or np.asarray(P2).shape != (3, 4):
        raise ValueError("P1, P2 must be 3x4 projection matrices")
    _fill(rig.P1, P1, 12, "P1")
    _fill(rig.P2, P2, 12, "P2")
    # scipy Rotation (what the reference passes) or a plain 3x3 matrix
    Rw = rot.as_matrix() if hasattr(rot, "as_matrix") else np.asarray(rot, dtype=np.float64)
    _fill(rig.Rw, Rw, 9, "rot")
    _fill(rig.tw, trans, 3, "trans")
    rig.agg_sum = 1 if agg_sum else 0
    rig.reserved = 0
    return rig


def rig_from_dicts(calibration: Dict[str, np.ndarray], transformation: Dict[str, np.ndarray],
                   agg_sum: bool = False) -> PtkRig:
    """Rig the way ``match_complex`` / ``match_csv_complex`` derive it
    (match2D.py:584-596, 689-710)."""
    from scipy.spatial.transform import Rotation

    P1, P2, r1, r2, t1, t2 = projection_matrices(calibration)
    if "translation" in transformation:
        rot = Rotation.from_matrix(transformation["rotation"])
        trans = np.asarray(transformation["translation"], dtype=np.float64)
    else:  # legacy 5-matrix format, match2D.py:703-710
        rotx = Rotation.from_matrix(np.asarray(transformation["M_rotate_x"])[0:3, 0:3])
        roty = Rotation.from_matrix(np.asarray(transformation["M_rotate_y"])[0:3, 0:3])
        rotz = Rotation.from_matrix(np.asarray(transformation["M_rotate_z"])[0:3, 0:3])
        tw1 = np.asarray(transformation["M_trans"])[0:3, 3]
        tw2 = np.asarray(transformation["M_trans2"])[0:3, 3]
        rot = rotz * roty * rotx
        trans = rot.apply(tw1) + tw2
    return make_rig(calibration, P1, P2, rot, trans, r1, r2, t1, t2, agg_sum)
