"""Per-kernel breakdown of the device step (K0..K5) for one shape: python tools/step_breakdown.py [frames] [rods]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from particletracking_b200 import engine, rig as rigmod, synthetic as syn

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
C = 8
data = syn.make_stereo(F, C, N, seed=2)
rig = rigmod.rig_from_dicts(data.calibration, data.transformation)
c1, c2, n1, n2 = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in data.problems())
for _ in range(3):
    engine.reconstruct_device(rig, c1, c2, n1, n2, F, C, track=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for _ in range(K):
    engine.reconstruct_device(rig, c1, c2, n1, n2, F, C, track=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
engine.kernel_timing(True); engine.kernel_timing_read(reset=True)
for _ in range(K):
    engine.reconstruct_device(rig, c1, c2, n1, n2, F, C, track=True)
torch.cuda.synchronize()
t = engine.kernel_timing_read()
engine.kernel_timing(False)
print(json.dumps({"frames": F, "rods": N, "ms_per_step": ms, "frames_per_s": F / ms * 1e3,
                  "kernel_ms": {k: round(v[0] / K, 4) for k, v in t.items() if v[1]}}))
