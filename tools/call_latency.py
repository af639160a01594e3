#!/usr/bin/env python
"""Latency of single reference-shaped calls (the RodTracker GUI's use: one match_frame per frame
and colour).  GPU only; prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np


def main():
    import torch
    from scipy.spatial.transform import Rotation

    from particletracking_b200 import synthetic as syn
    from particletracking_b200.reconstruct_3D import match2D, tracking

    out = {}
    for n in (12, 25, 32, 64):
        data = syn.make_stereo(12, 1, n, seed=n)
        df = syn.to_dataframes(data)["black"]
        cal, trf = data.calibration, data.transformation
        P1, P2, r1, r2, t1, t2 = syn.projection_matrices(cal)
        rot = Rotation.from_matrix(trf["rotation"])
        args = (cal, P1, P2, rot, trf["translation"], r1, r2, t1, t2)
        for _ in range(3):
            match2D.match_frame(df, "gp1", "gp2", 3, "black", *args)
        t0 = time.perf_counter()
        for f in range(10):
            match2D.match_frame(df, "gp1", "gp2", f, "black", *args)
        out[f"match_frame_ms_{n}rods"] = (time.perf_counter() - t0) / 10 * 1e3
        m = match2D.match_complex(df, list(range(12)), "black", cal, trf)[0]
        tracking.tracking_global_assignment(m)
        t0 = time.perf_counter()
        for _ in range(5):
            tracking.tracking_global_assignment(m)
        out[f"tracking_global_assignment_ms_per_pair_{n}rods"] = (time.perf_counter() - t0) / 5 / 11 * 1e3
    print(json.dumps(out))


if __name__ == "__main__":
    main()
