#!/usr/bin/env python
"""One launch pattern for profiling K10 (k_ap3): 8 rod-data problems of 32^3 weights, as one frame of
matchND.assign poses them.  Prints the CUDA-event time per launch with outputs pre-allocated (no
Python allocation between launches).  Test tooling: imports oracle/ to build the weights."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from tools.ap3_check import rod_weights


def main():
    from particletracking_b200 import _native as nat
    from particletracking_b200 import engine

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    W = torch.from_numpy(np.stack([rod_weights(n, 300 + c) for c in range(8)])).cuda()
    r = engine.npartite3_batch(W)
    torch.cuda.synchronize()
    B, D = W.shape[0], W.shape[1]
    L = nat.lib()
    ws = torch.empty(L.ptk_workspace_bytes(nat.WS_AP3, B, 0), dtype=torch.uint8, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    call = lambda: L.ptk_npartite3_batch(B, D, D, D, p(W), None, None, None, p(r.sol), p(r.n_sol), p(r.value), p(r.bound),
                                         p(r.status), p(r.stats), 0, 0, p(ws), ws.numel(), st)
    for _ in range(5):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        call()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"kernel": "k_ap3", "problems": B, "n": n, "us_per_launch": 1e3 * e0.elapsed_time(e1) / 50,
                      "status": r.status.cpu().tolist(), "steps": r.stats[:, 0].cpu().tolist()}))


if __name__ == "__main__":
    main()
