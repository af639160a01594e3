#!/usr/bin/env python
"""End-to-end time of the reference-facing CSV call (GPU): match_csv_complex on a synthetic
experiment written as rods_df_{color}.csv, BASELINE configs[1] size by default."""
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
    a = ap.parse_args()
    from particletracking_b200 import synthetic as syn
    from particletracking_b200.reconstruct_3D import match2D

    data = syn.make_stereo(a.frames, a.colors, a.rods, seed=2)
    with tempfile.TemporaryDirectory() as tmp:
        inp, outp = os.path.join(tmp, "in"), os.path.join(tmp, "out")
        syn.write_csvs(data, inp)
        cal, trf = os.path.join(tmp, "cal.json"), os.path.join(tmp, "trf.json")
        json.dump({k: np.asarray(v).tolist() for k, v in data.calibration.items()}, open(cal, "w"))
        json.dump({k: np.asarray(v).tolist() for k, v in data.transformation.items()}, open(trf, "w"))
        frames = [int(f) for f in data.frames]
        match2D.match_csv_complex(inp, outp, data.colors[:1], "gp1", "gp2", frames[:10], cal, trf)  # warm-up
        t0 = time.perf_counter()
        errs, lens = match2D.match_csv_complex(inp, outp, data.colors, "gp1", "gp2", frames, cal, trf)
        dt = time.perf_counter() - t0
        size = sum(os.path.getsize(os.path.join(outp, f)) for f in os.listdir(outp))
    st = getattr(match2D.match_csv_complex, "last_stage_seconds", {})
    stage_ms = {k: round(1e3 * sum(v[k] for v in st.values()) / max(len(st), 1), 2) for k in ("read", "match", "write")}
    print(json.dumps({"call": "match2D.match_csv_complex", "mean_stage_ms_per_colour": stage_ms, "frames": a.frames, "colours": a.colors, "rods": a.rods,
                      "seconds": dt, "frames_per_s": a.frames / dt, "output_MB": size / 1e6,
                      "errs_shape": list(errs.shape)}))


if __name__ == "__main__":
    main()
