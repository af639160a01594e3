#!/usr/bin/env python
"""BASELINE config 5: rods per colour 8 -> 256, 8 colours, device-resident step (match + track +
chain) timed with CUDA events; one JSON line per N.  GPU only.  At 1 GPU: one library call per step
(ptk_reconstruct_device).  Under torchrun (2 / 4 / 8 ranks): every rank runs `--frames` frames of its
own (weak scaling) through ShardedReconstructor (halo frame re-computed, one NCCL gather of the result
block per step, root-side re-base), max over ranks.

  python tools/sweep.py [--frames 2000] > gpurun_out/sweep.jsonl
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tools/sweep.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from particletracking_b200 import distributed as pdist
from particletracking_b200 import engine, rig as rigmod, synthetic as syn

FLOP_PER_PAIR = 1448.0  # K1 as built (bench.py: 141.3 DFMA + 61.2 DMUL + 18.1 DADD per combination); the reference-algorithm model is 8840


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--rods", type=int, nargs="*", default=[8, 16, 32, 64, 128, 256])
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    C = 8
    peak = engine.measure_fp64_peak(20000)
    for N in a.rods:
        F = a.frames if N <= 128 else max(a.frames // 4, 50)
        halo = 1 if (world > 1 and rank < world - 1) else 0
        data = syn.make_stereo(F + halo, C, N, seed=5 + rank)
        rig = rigmod.rig_from_dicts(data.calibration, data.transformation)
        t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
        c1, c2, n1, n2 = (t(x) for x in data.problems())
        if world > 1:
            rec = pdist.ShardedReconstructor(rig, F, C, N, dev, use_graph=False)
            step = lambda: rec.step(c1, c2, n1, n2)
            fin = rec.finish
        else:
            pipe = engine.DevicePipeline(rig, F, C, N, dev, track=True)
            step = lambda: pipe.run(c1, c2, n1, n2)
            fin = lambda: None
        for _ in range(2):
            out = step()
        fin()
        torch.cuda.synchronize()
        st = out["status"] if world > 1 else out.status
        assert int((st != 0).sum()) == 0
        if world > 1:
            dist.barrier()
        engine.kernel_timing(True)
        engine.kernel_timing_read(reset=True)
        K = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            step()
        fin()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        tm = engine.kernel_timing_read(reset=True)
        engine.kernel_timing(False)
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        k1 = tm["cost"][0] / K
        flops = FLOP_PER_PAIR * (F + halo) * C * N * N
        if rank == 0:
            print(json.dumps({
                "n_gpus": world, "rods": N, "frames_per_gpu": F, "colours": C, "ms_per_step": ms,
                "frames_per_s": world * F / (ms * 1e-3),
                "k_cost_ms": k1, "k_cost_tflops": flops / (k1 * 1e-3) / 1e12, "fp64_peak_tflops": peak / 1e12,
                "k_cost_frac": flops / (k1 * 1e-3) / peak,
                "k_cost_frac_reference_algorithm": flops * (8840.0 / FLOP_PER_PAIR) / (k1 * 1e-3) / peak,
                # DFMA + DMUL + DADD + DSETP = 224.6 FP64-pipe instructions per combination, two pipe cycles per
                # warp instruction, 16 lanes per scheduler (profiles/r02_k_cost_ncu.txt)
                "k_cost_fp64_pipe_busy_est": 224.6 * 4 * (F + halo) * C * N * N / 32 * 2 / (148 * 4) / (k1 * 1e-3 * 1.965e9),
                "kernel_ms": {k: v[0] / K for k, v in tm.items() if v[1]},
            }), flush=True)
        del c1, c2, n1, n2, out
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
