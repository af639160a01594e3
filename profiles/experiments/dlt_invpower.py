#!/usr/bin/env python
"""Accuracy / convergence of dlt_null_invpower (QR + inverse power by squaring, csrc/ptk_device.cuh)
against cv2.triangulatePoints, on the CPU, through the host build of the same function
(dlt_invpower.cu).  Every cam1 x cam2 rod pair x 4 end-point combinations, i.e. mostly MIS-paired
rods, which is what K1 evaluates.  Design tooling (reads /root/reference example data when present)."""
import ctypes
import json
import os
import subprocess
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from particletracking_b200 import synthetic as syn  # noqa: E402


def build():
    so = "/tmp/ipw/dlt_invpower.so"
    os.makedirs("/tmp/ipw", exist_ok=True)
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-fmad=false", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
                           "-o", so, os.path.join(HERE, "dlt_invpower.cu")])
    lib = ctypes.CDLL(so)
    return lib


def combos(a, b):
    i, j, e1, e2 = np.meshgrid(np.arange(len(a)), np.arange(len(b)), [0, 1], [0, 1], indexing="ij")
    return np.ascontiguousarray(np.concatenate([a[i.ravel(), e1.ravel()], b[j.ravel(), e2.ravel()]], axis=1))


def check(lib, label, P1, P2, pts):
    P1 = np.ascontiguousarray(P1, dtype=np.float64)
    P2 = np.ascontiguousarray(P2, dtype=np.float64)
    M = len(pts)
    X = np.empty((M, 3))
    conv = np.empty(M, np.int32)
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    lib.tri_invpower_host(p(P1), p(P2), ctypes.c_longlong(M), p(pts), p(X), p(conv))
    X4 = cv2.triangulatePoints(P1, P2, pts[:, :2].T.copy(), pts[:, 2:].T.copy())
    Xc = (X4[:3] / X4[3]).T
    A = np.stack([pts[:, 0, None] * P1[2] - P1[0], pts[:, 1, None] * P1[2] - P1[1],
                  pts[:, 2, None] * P2[2] - P2[0], pts[:, 3, None] * P2[2] - P2[1]], axis=1)
    sv = np.linalg.svd(A, compute_uv=False)
    ratio = sv[:, 3] / sv[:, 2]
    ok = conv == 1
    err = np.abs(X - Xc).max(1) / np.abs(Xc).max(1)
    out = dict(case=label, combos=int(M), not_converged=int((~ok).sum()), max_rel_err_converged=float(err[ok].max()) if ok.any() else None,
               median_rel_err=float(np.median(err[ok])) if ok.any() else None,
               sigma4_over_sigma3=dict(median=float(np.median(ratio)), p999=float(np.percentile(ratio, 99.9)), max=float(ratio.max())),
               max_ratio_converged=float(ratio[ok].max()) if ok.any() else None,
               min_ratio_not_converged=float(ratio[~ok].min()) if (~ok).any() else None,
               sigma1_over_sigma3_max=float((sv[:, 0] / sv[:, 2]).max()))
    print(json.dumps(out), flush=True)
    return out


def main():
    lib = build()
    for label, kw in [("synthetic 32 rods sigma 0.25 px", dict(n_rods=32, seed=2, sigma_px=0.25)),
                      ("synthetic noise-free", dict(n_rods=32, seed=3, sigma_px=0.0)),
                      ("synthetic sigma 3 px + 10% dummies", dict(n_rods=32, seed=4, sigma_px=3.0, p_dummy=0.1)),
                      ("synthetic 64 rods", dict(n_rods=64, seed=5, sigma_px=0.5)),
                      ("synthetic 256 rods", dict(n_rods=256, seed=6, sigma_px=0.5))]:
        s = syn.make_stereo(2, 1, kw.pop("n_rods"), kw.pop("seed"), **kw)
        P1, P2 = syn.projection_matrices(s.calibration)[:2]
        check(lib, label, P1, P2, combos(s.cam1[0, 0, : s.n1[0, 0]], s.cam2[0, 0, : s.n2[0, 0]]))
    ref = "/root/reference/ParticleDetection/tests/example_data"
    if os.path.isdir(ref):  # the reference's own rig and detections (build container only)
        import pandas as pd

        cal = json.load(open(os.path.join(ref, "gp34.json")))
        CM1, CM2, R, T = (np.array(cal[k], dtype=float) for k in ("CM1", "CM2", "R", "T"))
        P1 = CM1 @ np.hstack([np.eye(3), np.zeros((3, 1))])
        P2 = CM2 @ np.hstack([R, T.reshape(3, 1)])
        for color in ("black", "green"):
            df = pd.read_csv(os.path.join(ref, f"rods_df_{color}.csv"))
            pts = []
            for f in sorted(df.frame.unique()):
                d = df[df.frame == f]
                a = d[["x1_gp3", "y1_gp3", "x2_gp3", "y2_gp3"]].to_numpy(float).reshape(-1, 2, 2)
                b = d[["x1_gp4", "y1_gp4", "x2_gp4", "y2_gp4"]].to_numpy(float).reshape(-1, 2, 2)
                a = a[~np.isnan(a).any(axis=(1, 2))]
                b = b[~np.isnan(b).any(axis=(1, 2))]
                pts.append(combos(a, b))
            check(lib, f"reference example data, {color} (gp34 rig)", P1, P2, np.concatenate(pts))


if __name__ == "__main__":
    main()
