#!/usr/bin/env python
"""BASELINE config 4: matching + frame-to-frame tracking over 50,000 frames, 8 colours x 64 rods,
frames sharded across the GPUs of one box (run under torchrun, one rank per GPU).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/config4.py

Every rank synthesises its own contiguous frame range (seeded per rank), runs
ShardedReconstructor (match -> halo -> track -> chain -> gathers), and rank 0 checks
size-independent properties on the gathered result and prints one JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=50000)
    ap.add_argument("--colors", type=int, default=8)
    ap.add_argument("--rods", type=int, default=64)
    ap.add_argument("--repeat", type=int, default=3)
    a = ap.parse_args()
    from particletracking_b200 import distributed as pdist
    from particletracking_b200 import rig as rigmod, synthetic as syn

    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ["NCCL_DEBUG"] = "WARN"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    f0, f1 = pdist.shard_frames(a.frames, world, rank)
    F, C, N = f1 - f0, a.colors, a.rods
    data = syn.make_stereo(F, C, N, seed=4000 + rank, first_frame=f0)
    rig = rigmod.rig_from_dicts(data.calibration, data.transformation)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    c1, c2, n1, n2 = (t(x) for x in data.problems())
    rec = pdist.ShardedReconstructor(rig, F, C, N, dev, n_frames_total=a.frames)
    out = rec.step(c1, c2, n1, n2)
    rec.finish()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(a.repeat):
        out = rec.step(c1, c2, n1, n2)
    rec.finish()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.repeat], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    wall = (time.perf_counter() - t0) / a.repeat
    # per-rank checks
    ok_status = int((out["status"] != 0).sum().item()) == 0 and int((out["track_status"] != 0).sum().item()) == 0
    col = out["col_ind"].cpu().numpy()
    truth = data.truth_perm.reshape(-1, N)
    hit = float((np.take_along_axis(truth, col, axis=1) == np.arange(N)).mean())
    tid = out["track_id"].cpu().numpy()
    perm_ok = bool((np.sort(tid, axis=-1) == np.arange(N)).all())
    stats = torch.tensor([float(ok_status), hit, float(perm_ok)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MIN)
    if rank == 0:
        rows_all, tid_all = out["rows_all"], out["track_id_all"]
        n_rows = sum(int(r.shape[0]) for r in rows_all)
        # chain continuity across rank boundaries: ids follow the boundary pair's assignment
        cont = True
        tcol0 = out["track_col"].cpu().numpy()  # rank 0's pairs, the last one is the halo pair
        if world > 1:
            a_ids = tid_all[0][:, -1].cpu().numpy()
            b_ids = tid_all[1][:, 0].cpu().numpy()
            cont = bool((np.take_along_axis(b_ids, tcol0[:, -1].astype(np.int64), axis=1) == a_ids).all())
        print(json.dumps({
            "config": f"{a.frames} frames x {C} colours x {N} rods, {world} GPU(s)", "ms": float(ms.item()),
            "wall_s": wall, "frames_per_s": a.frames / (float(ms.item()) * 1e-3),
            "all_status_ok": bool(stats[0].item() == 1.0), "planted_correspondence_recovered_min_over_ranks": float(stats[1].item()),
            "track_ids_are_permutations": bool(stats[2].item() == 1.0), "chain_continuous_across_rank0_1": cont,
            "frames_gathered_on_rank0": n_rows,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
