"""Multi-GPU correctness of ShardedReconstructor on real GPUs (run under torchrun, one rank per GPU):
the gathered rows / global track ids / pair totals on rank 0 must equal a single-GPU run over the
whole frame range, bit for bit.  Unequal shards, several steps (both buffer slots), CUDA-graph replay.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py
"""
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

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = {}
for (F, C, N, drop) in ((203, 4, 16, 0.0), (64, 8, 32, 0.0), (61, 3, 12, 0.03)):
    data = syn.make_stereo(F, C, N, seed=31 + F, p_drop=drop)
    rig = rigmod.rig_from_dicts(data.calibration, data.transformation)
    f0, f1 = pdist.shard_frames(F, world, rank)
    i0, i1 = pdist.shard_input_range(F, world, rank)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    Fin = i1 - i0
    args = (t(data.cam1[i0:i1].reshape(Fin * C, N, 2, 2)), t(data.cam2[i0:i1].reshape(Fin * C, N, 2, 2)),
            t(data.n1[i0:i1].reshape(-1)), t(data.n2[i0:i1].reshape(-1)))
    for to_root in (True, False):
        rec = pdist.ShardedReconstructor(rig, f1 - f0, C, N, dev, n_frames_total=F, gather_to_root=to_root)
        for _ in range(5):
            out = rec.step(*args)
        rec.finish()
        torch.cuda.synchronize()
        # single-GPU truth over the whole range
        pipe = engine.DevicePipeline(rig, F, C, N, dev, track=True)
        c1, c2, n1, n2 = (t(x) for x in data.problems())
        pipe.run(c1, c2, n1, n2)
        torch.cuda.synchronize()
        key = f"{F}x{C}x{N}{'d' if drop else ''}_{'root' if to_root else 'all'}"
        if to_root:
            if rank == 0:
                rows = torch.cat(list(out["rows_all"]), dim=0).reshape(F * C, N, 18)
                tid = torch.cat(list(out["track_id_all"]), dim=1)
                tot = torch.cat(list(out["track_total_all"]), dim=1)
                ok = (torch.equal(torch.nan_to_num(rows, nan=-7.0), torch.nan_to_num(pipe.rows, nan=-7.0)) and torch.equal(tid, pipe.track_id)
                      and torch.equal(torch.nan_to_num(tot, nan=-7.0), torch.nan_to_num(pipe.track_total, nan=-7.0)))
                res[key] = bool(ok)
        else:
            ok = torch.equal(out["track_id_global"], pipe.track_id[:, f0:f1])
            flag = torch.tensor([1.0 if ok else 0.0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                res[key] = bool(flag.item() == 1.0)
if rank == 0:
    print(json.dumps({"world": world, "checks": res, "all_ok": all(res.values())}), flush=True)
dist.destroy_process_group()
