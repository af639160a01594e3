#!/usr/bin/env python
"""BASELINE config 5: rods per colour 8 -> 256, 8 colours, device-resident step (match + track +
chain) timed with CUDA events; one JSON line per N.  GPU only.

  python tools/sweep.py [--frames 2000] > gpurun_out/sweep.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from particletracking_b200 import engine, rig as rigmod, synthetic as syn

FLOP_PER_PAIR = 2400.0  # K1 as built (bench.py); the reference-algorithm model is 8840


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--rods", type=int, nargs="*", default=[8, 16, 32, 64, 128, 256])
    a = ap.parse_args()
    C = 8
    peak = engine.measure_fp64_peak(20000)
    for N in a.rods:
        F = a.frames if N <= 128 else max(a.frames // 4, 50)
        data = syn.make_stereo(F, C, N, seed=5)
        rig = rigmod.rig_from_dicts(data.calibration, data.transformation)
        t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
        c1, c2, n1, n2 = (t(x) for x in data.problems())
        for _ in range(2):
            out = engine.reconstruct_device(rig, c1, c2, n1, n2, F, C, track=True)
        torch.cuda.synchronize()
        assert int((out[0].status != 0).sum()) == 0
        engine.kernel_timing(True)
        engine.kernel_timing_read(reset=True)
        K = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            engine.reconstruct_device(rig, c1, c2, n1, n2, F, C, track=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        tm = engine.kernel_timing_read(reset=True)
        engine.kernel_timing(False)
        k1 = tm["cost"][0] / K
        flops = FLOP_PER_PAIR * F * C * N * N
        print(json.dumps({
            "rods": N, "frames": F, "colours": C, "ms_per_step": ms, "frames_per_s": F / (ms * 1e-3),
            "k_cost_ms": k1, "k_cost_tflops": flops / (k1 * 1e-3) / 1e12, "fp64_peak_tflops": peak / 1e12,
            "k_cost_frac": flops / (k1 * 1e-3) / peak,
            "k_cost_frac_reference_algorithm": flops * (8840.0 / FLOP_PER_PAIR) / (k1 * 1e-3) / peak,
            "kernel_ms": {k: v[0] / K for k, v in tm.items() if v[1]},
        }), flush=True)
        del c1, c2, n1, n2, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
