"""What saturates the host-buffer path at 4 and 8 ranks?  Copy-only probe: every
active rank moves the e2e path's bytes per step (H2D 16.4 MB, D2H 43.2 MB, pinned host buffers) with no
kernels, alone and together, for 1 / 2 / 4 / 8 concurrently active ranks.  Run under torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/copy_probe.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
H2D, D2H, IT = 16_448_000, 43_166_880, 40
hin = torch.empty(H2D, dtype=torch.uint8).pin_memory()
hout = torch.empty(D2H, dtype=torch.uint8).pin_memory()
din = torch.empty(H2D, dtype=torch.uint8, device=dev)
dout = torch.empty(D2H, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
n_active = 1
while n_active <= world:
    for mode in ("h2d", "d2h", "both"):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if rank < n_active:
            for _ in range(IT):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        din.copy_(hin, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        hout.copy_(dout, non_blocking=True)
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        nbytes = (H2D if mode != "d2h" else 0) + (D2H if mode != "h2d" else 0)
        gbs = torch.tensor([nbytes * IT / dt / 1e9 if rank < n_active else 0.0], device=dev, dtype=torch.float64)
        mn = gbs.clone() if rank < n_active else torch.tensor([1e30], device=dev, dtype=torch.float64)
        dist.all_reduce(gbs, op=dist.ReduceOp.SUM)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        res[f"{n_active}_{mode}"] = {"aggregate_GBs": round(float(gbs.item()), 1), "slowest_rank_GBs": round(float(mn.item()), 1),
                                     "steps_per_s_per_rank": round(float(mn.item()) * 1e9 / nbytes, 1)}
    n_active *= 2
if rank == 0:
    try:
        topo = os.popen("nvidia-smi topo -m 2>/dev/null | head -14").read()
    except Exception:
        topo = ""
    print(json.dumps({"world": world, "bytes": {"h2d": H2D, "d2h": D2H}, "results": res, "cpus": len(os.sched_getaffinity(0))}, indent=1))
    print(topo, file=sys.stderr)
dist.destroy_process_group()
