"""A/B of K1 launch-bound variants (PTK_KCOST_CFG = min blocks per SM) on configs[1]-shaped data."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from particletracking_b200 import engine, rig as rigmod, synthetic as syn
    N = int(os.environ.get("KC_N", "32"))
    data = syn.make_stereo(int(os.environ.get("KC_F", 1000 if N <= 32 else 250)), 8, N, seed=2)
    rig = rigmod.rig_from_dicts(data.calibration, data.transformation)
    args = tuple(torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in data.problems())
    for _ in range(3):
        engine.cost_batch(rig, *args, want_choice=False)
    torch.cuda.synchronize()
    engine.kernel_timing(True); engine.kernel_timing_read(reset=True)
    for _ in range(10):
        c, _, _, _ = engine.cost_batch(rig, *args, want_choice=False)
    torch.cuda.synchronize()
    t = engine.kernel_timing_read()
    print(json.dumps({"cfg": os.environ.get("PTK_KCOST_CFG", "6"), "N": N, "k_cost_ms": t["cost"][0] / t["cost"][1], "chk": float(c.nan_to_num().sum())}))
else:
    for n in ("32", "64"):
        for cfg in ("4", "5", "6", "7", "8"):
            env = dict(os.environ, PTK_KCOST_CFG=cfg, KC_N=n)
            print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
