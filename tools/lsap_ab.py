"""A/B of libptk.so variants on the matching LSAP (K2) alone: python tools/lsap_ab.py lib1.so lib2.so ..."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from particletracking_b200 import _native as nat
from particletracking_b200 import engine, rig as rigmod, synthetic as syn

def use(path):
    nat._lib = None
    nat.LIB_PATH = os.path.abspath(path)
    return nat.lib()

N = int(os.environ.get("AB_N", 32)); F = int(os.environ.get("AB_F", 1000 if N <= 32 else 250))
data = syn.make_stereo(F, 8, N, seed=1001)
rig = rigmod.rig_from_dicts(data.calibration, data.transformation)
args = tuple(torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in data.problems())
ref = None
for path in sys.argv[1:]:
    use(path)
    for _ in range(3):
        r = engine.match_batch(rig, *args)
    torch.cuda.synchronize()
    engine.kernel_timing(True); engine.kernel_timing_read(reset=True)
    for _ in range(10):
        r = engine.match_batch(rig, *args)
    torch.cuda.synchronize()
    t = engine.kernel_timing_read(); engine.kernel_timing(False)
    col = r.col_ind.cpu().numpy()
    if ref is None: ref = col
    print(json.dumps({"lib": os.path.basename(path), "N": N, "lsap_match_ms": t["lsap_match"][0] / t["lsap_match"][1],
                      "cost_ms": t["cost"][0] / t["cost"][1], "same": bool((col == ref).all())}), flush=True)
