#!/usr/bin/env python
"""End-to-end time of matchND.assign (GPU): CSV in -> match + track frame by frame -> CSV out, on a
synthetic experiment (BASELINE configs[1] size by default), plus the CPU restatement of the same chain
(oracle/npartite.py: cv2 + create_weights + HiGHS for the reference's 0/1 programme) on a small sample
as the baseline.  Test tooling: imports oracle/ for the baseline leg only."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--colors", type=int, default=8)
    ap.add_argument("--rods", type=int, default=32)
    ap.add_argument("--cpu-frames", type=int, default=3, help="frames of one colour for the CPU baseline (0 = skip)")
    a = ap.parse_args()
    import torch

    from particletracking_b200 import engine
    from particletracking_b200 import synthetic as syn
    from particletracking_b200.reconstruct_3D import matchND

    data = syn.make_stereo(a.frames, a.colors, a.rods, seed=2)
    out = {"call": "matchND.assign", "frames": a.frames, "colours": a.colors, "rods": a.rods}
    with tempfile.TemporaryDirectory() as tmp:
        inp, outp = os.path.join(tmp, "in"), os.path.join(tmp, "out")
        syn.write_csvs(data, inp)
        cal, trf = os.path.join(tmp, "cal.json"), os.path.join(tmp, "trf.json")
        json.dump({k: np.asarray(v).tolist() for k, v in data.calibration.items()}, open(cal, "w"))
        json.dump({k: np.asarray(v).tolist() for k, v in data.transformation.items()}, open(trf, "w"))
        frames = [int(f) for f in data.frames]
        matchND.assign(inp, outp, data.colors[:1], "gp1", "gp2", frames[:5], cal, trf)  # warm-up
        engine.kernel_timing(True)
        engine.kernel_timing_read(reset=True)
        t0 = time.perf_counter()
        errs, lens = matchND.assign(inp, outp, data.colors, "gp1", "gp2", frames, cal, trf)
        dt = time.perf_counter() - t0
        torch.cuda.synchronize()
        timing = engine.kernel_timing_read(reset=True)
        engine.kernel_timing(False)
        st = matchND.assign.last
        # tracking quality against the generator's ground truth: every particle id follows one rod
        import pandas as pd

        res = pd.read_csv(os.path.join(outp, f"rods_df_{data.colors[0]}.csv"), index_col=0)
        ctr = res.sort_values(["frame", "particle"])[["x", "y", "z"]].to_numpy().reshape(a.frames, -1, 3)
        out.update(seconds=dt, frames_per_s=a.frames / dt, errs_shape=list(errs.shape),
                   status_ok=bool((st["status"] == 0).all()), solver_stats_sum=st["solver_stats"].sum(axis=0).tolist(),
                   max_centre_jump=float(np.abs(np.diff(ctr, axis=0)).max()),
                   kernel_ms={k: round(v[0], 3) for k, v in timing.items() if v[1]},
                   kernel_launches={k: int(v[1]) for k, v in timing.items() if v[1]})
    if a.cpu_frames > 1:
        from oracle import npartite as onp
        from oracle import port

        P1, P2, r1, r2, t1, t2 = syn.projection_matrices(data.calibration)
        args = (data.calibration, P1, P2, data.transformation["rotation"], data.transformation["translation"], r1, r2, t1, t2)
        t0 = time.perf_counter()
        prev = None
        for f in range(a.cpu_frames):
            c1, c2 = data.cam1[f, 0, : data.n1[f, 0]], data.cam2[f, 0, : data.n2[f, 0]]
            rows = port.match_arrays(c1, c2, *args)["rows"] if f == 0 else onp.match_frame_arrays(c1, c2, prev, *args)[0]
            prev = rows[:, :6].reshape(-1, 2, 3)
        cdt = time.perf_counter() - t0
        per_fc = cdt / a.cpu_frames
        out["cpu_baseline"] = {"kind": "port (cv2 + numpy create_weights + HiGHS via scipy.optimize.milp)", "cores": 1,
                               "sample": f"{a.cpu_frames} frames x 1 colour x {a.rods} rods",
                               "seconds_per_frame_colour": per_fc, "frames_per_s": 1.0 / (per_fc * a.colors)}
        out["speedup_vs_cpu_port"] = out["frames_per_s"] / out["cpu_baseline"]["frames_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
