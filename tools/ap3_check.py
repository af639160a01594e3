#!/usr/bin/env python
"""GPU check / timing of the 3-index assignment solver (K10) against the HiGHS restatement of the
reference's programme (oracle/npartite.py): optimal value, feasibility, solver statistics.
Test tooling: imports oracle/."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def rod_weights(N, seed, sigma=0.25, p_drop=0.0, p_dummy=0.0):
    from oracle import port
    from particletracking_b200 import synthetic as syn

    s = syn.make_stereo(2, 1, N, seed, sigma_px=sigma, p_drop=p_drop, p_dummy=p_dummy)
    cal = s.calibration
    P1, P2, r1, r2, t1, t2 = syn.projection_matrices(cal)
    trf = s.transformation
    out = []
    for f in range(2):
        a = s.cam1[f, 0, : s.n1[f, 0]]
        b = s.cam2[f, 0, : s.n2[f, 0]]
        out.append(port.match_arrays(a, b, cal, P1, P2, trf["rotation"], trf["translation"], r1, r2, t1, t2, agg="sum"))
    prev = out[0]["rows"][:, :6].reshape(-1, 2, 3)
    return port.create_weights(out[1]["p_world"], prev, out[1]["rep_errs"])[0]


def main():
    from oracle import npartite as onp
    from particletracking_b200 import engine

    dev = torch.device("cuda", 0)
    cases = []
    for N in (5, 12, 25, 32):
        for seed in range(3):
            cases.append((f"rods N={N} seed={seed}", rod_weights(N, 100 + seed), None))
    for seed in range(3):  # very noisy detections
        cases.append((f"rods N=25 sigma=3px seed={seed}", rod_weights(25, 600 + seed, sigma=3.0), None))
    # dropped / dummy detections (identical rows and columns, near-degenerate duals): optima stored by
    # tools/ap3_optima.py
    with open(os.path.join(ROOT, "tests", "golden", "ap3_rod_optima.json")) as f:
        for name, rec in json.load(f).items():
            cases.append((name, rod_weights(**rec["args"]), rec["milp"]))
    rng = np.random.default_rng(0)
    for dims in [(2, 3, 4), (4, 3, 7), (6, 6, 6), (8, 12, 5), (10, 10, 10), (12, 12, 12), (1, 1, 1), (1, 5, 3)]:
        for rep in range(3):
            cases.append((f"random {dims} #{rep}", rng.random(dims), None))
    cases.append(("integer ties (6,6,6)", rng.integers(0, 4, (6, 6, 6)).astype(float), None))
    cases.append(("mixed sign (6,7,5)", rng.normal(size=(6, 7, 5)), None))
    cases.append(("all negative (4,4,4)", -rng.random((4, 4, 4)), None))
    cases.append(("zeros (5,5,5)", np.zeros((5, 5, 5)), None))
    cases.append(("constant (5,4,6)", np.full((5, 4, 6), 2.5), None))
    bad = 0
    for name, w, known in cases:
        t0 = time.perf_counter()
        r = engine.npartite3_batch(torch.from_numpy(np.ascontiguousarray(w)[None]).to(dev))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n = int(r.n_sol[0])
        sol = r.sol[0, :, :n].cpu().numpy()
        val = float(r.value[0])
        ref = known if known is not None else onp.objective(w, onp.npartite_matching(w))
        feas = onp.is_feasible(w.shape, sol)
        ok = feas and abs(val - ref) <= 1e-9 * max(1.0, abs(ref)) and abs(onp.objective(w, sol) - val) <= 1e-9 * max(1.0, abs(val))
        bad += not ok
        print(json.dumps({"case": name, "ok": bool(ok), "value": val, "milp": ref, "bound": float(r.bound[0]), "n": n,
                          "status": int(r.status[0]), "stats": r.stats[0].cpu().tolist(), "ms": round(dt * 1e3, 3)}))
    # batch of 8 identical-size problems (the colours of one frame), timing with events
    W = np.stack([rod_weights(32, 300 + c) for c in range(8)])
    Wd = torch.from_numpy(W).to(dev)
    for _ in range(3):
        r = engine.npartite3_batch(Wd)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        r = engine.npartite3_batch(Wd)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"batch": "8 x rods N=32", "ms_per_call": e0.elapsed_time(e1) / 20, "status": r.status.cpu().tolist(),
                      "stats": r.stats.cpu().tolist()}))
    print("FAILED" if bad else "ALL OK", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
