#!/usr/bin/env python
"""A/B timing of libptk.so variants on the bench workload (development tool, GPU only).

  python tools/ab_bench.py particletracking_b200/csrc/ab/*.so

For every library: K0+K1 (ptk_cost_batch) and the full device step are timed with CUDA events on
configs[1]-sized data (1000 x 8 x 32); results are also checked against the first library
(assignments identical, rows within 1e-9)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from particletracking_b200 import _native as nat
from particletracking_b200 import engine, rig as rigmod, synthetic as syn


def use(path):
    nat._lib = None
    nat.LIB_PATH = os.path.abspath(path)
    return nat.lib()


def main():
    F, C, N = int(os.environ.get("AB_F", 1000)), 8, int(os.environ.get("AB_N", 32))
    data = syn.make_stereo(F, C, N, seed=1001)
    rig = rigmod.rig_from_dicts(data.calibration, data.transformation)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    c1, c2, n1, n2 = (t(x) for x in data.problems())
    ref = None
    for path in sys.argv[1:]:
        use(path)
        out = {}
        for name, fn in (("cost", lambda: engine.cost_batch(rig, c1, c2, n1, n2, want_choice=False)),
                         ("step", lambda: engine.reconstruct_device(rig, c1, c2, n1, n2, F, C, track=True))):
            for _ in range(3):
                r = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                r = fn()
            e1.record()
            torch.cuda.synchronize()
            out[name] = e0.elapsed_time(e1) / 10
        m = r[0]
        rows, col = m.rows.cpu().numpy(), m.col_ind.cpu().numpy()
        tcol, ttot = r[1].cpu().numpy(), r[2].cpu().numpy()
        if ref is None:
            ref = (rows, col, tcol, ttot)
            out["vs_first"] = "reference"
        else:
            same = bool((col == ref[1]).all())
            tsame = bool((tcol == ref[2]).all())
            terr = float(np.max(np.abs(ttot - ref[3]) / np.abs(ref[3])))
            err = float(np.nanmax(np.abs(rows[..., :10] - ref[0][..., :10])) / max(np.nanmax(np.abs(ref[0][..., :10])), 1))
            out["vs_first"] = (f"assignments identical={same}, rows max rel diff={err:.2e}, tracking col_ind identical={tsame}, "
                               f"total_cost max rel diff={terr:.1e}")
        print(json.dumps({"lib": os.path.basename(path), **out}), flush=True)


if __name__ == "__main__":
    main()
