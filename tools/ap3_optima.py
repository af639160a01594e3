#!/usr/bin/env python
"""Writes tests/golden/ap3_rod_optima.json: optimal values (HiGHS restatement of the reference's 0/1
programme, oracle/npartite.py) of 3-index weight cubes built from synthetic rod frames with dropped
and dummy detections.  The cubes themselves are regenerated from the recorded arguments
(tools/ap3_check.rod_weights: oracle/port.py around cv2, CPU only); only the optima are stored, since
HiGHS needs up to a minute per 48^3 cube.  Test tooling: imports oracle/."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def specs():
    s = {}
    for seed in range(4):  # the detector's dummy rods (-1,-1,-1,-1): identical rows / columns of the cube
        s[f"d16_{seed}"] = dict(N=16, seed=400 + seed, sigma=0.5, p_dummy=0.15)
        s[f"d32_{seed}"] = dict(N=32, seed=500 + seed, sigma=0.5, p_dummy=0.1)
    for seed in range(3):
        s[f"drop16_{seed}"] = dict(N=16, seed=200 + seed, sigma=1.0, p_drop=0.15)
        s[f"dd25_{seed}"] = dict(N=25, seed=700 + seed, sigma=1.0, p_drop=0.1, p_dummy=0.1)
    for seed in range(6):
        s[f"e32_{seed}"] = dict(N=32, seed=800 + seed, sigma=0.5, p_dummy=0.15)
        s[f"e24_{seed}"] = dict(N=24, seed=900 + seed, sigma=2.0, p_dummy=0.2, p_drop=0.1)
    for seed in range(2):
        s[f"e48_{seed}"] = dict(N=48, seed=1000 + seed, sigma=0.5, p_dummy=0.1)
    return s


def main():
    from oracle import npartite as onp
    from tools.ap3_check import rod_weights

    out = {}
    for name, args in specs().items():
        w = rod_weights(**args)
        out[name] = dict(args=args, shape=list(w.shape), milp=onp.objective(w, onp.npartite_matching(w)))
        print(name, out[name], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "ap3_rod_optima.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
