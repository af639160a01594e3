"""A/B of the tracking LSAP launch configurations (PTK_TRK_CFG) on configs[1]-shaped data."""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from particletracking_b200 import engine
    rng = np.random.default_rng(0)
    C, F, N = 8, int(os.environ.get("TRK_F", "1000")), int(os.environ.get("TRK_N", "32"))
    base = rng.uniform(-40, 40, size=(C, 1, N, 2, 3))
    ends = base + np.cumsum(rng.normal(size=(C, F, N, 2, 3)) * 0.2, axis=1)
    e = torch.from_numpy(ends).cuda()
    for _ in range(5):
        col, tot, st, _ = engine.track_batch(e)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(20):
        col, tot, st, _ = engine.track_batch(e)
    t1.record(); torch.cuda.synchronize()
    print(json.dumps({"cfg": os.environ.get("PTK_TRK_CFG", "0"), "N": N, "ms": t0.elapsed_time(t1) / 20, "ok": int((st != 0).sum()) == 0,
                      "chk": int(col.long().sum())}))
else:
    for n in (32,):
        for cfg in ("0", "1", "2", "3", "4", "5"):
            env = dict(os.environ, PTK_TRK_CFG=cfg, TRK_N=str(n))
            print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
