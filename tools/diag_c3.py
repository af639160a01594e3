"""Diagnostic: config-3 problems whose GPU assignment differs from the C oracle's."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import c_oracle as co
from particletracking_b200 import engine, rig as rigmod, synthetic as syn

data = syn.make_stereo(10000, 8, 32, seed=3, p_drop=0.05, p_dummy=0.02, sigma_px=0.5)
rig = rigmod.rig_from_dicts(data.calibration, data.transformation)
c1, c2, n1, n2 = data.problems()
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
r = engine.match_batch(rig, t(c1), t(c2), t(n1), t(n2), want_cost=False)
col = r.col_ind.cpu().numpy(); row = r.row_ind.cpu().numpy()
o = co.match_batch(rig, c1, c2, n1, n2, nthreads=0)
bad = np.where((col != o["col_ind"]).any(axis=1) | (row != o["row_ind"]).any(axis=1))[0]
print("mismatching problems:", len(bad), "of", len(n1))
out = []
for p in bad[:12]:
    sub = (t(c1[p:p+1]), t(c2[p:p+1]), t(n1[p:p+1]), t(n2[p:p+1]))
    cg, _, _, _ = engine.cost_batch(rig, *sub)
    cg = cg.cpu().numpy()[0]
    oc = co.match_batch(rig, c1[p:p+1], c2[p:p+1], n1[p:p+1], n2[p:p+1], want_cost=True)["cost"][0]
    k = o["n_out"][p]
    a_r, a_c = o["row_ind"][p, :k], o["col_ind"][p, :k]
    b_r, b_c = row[p, :k], col[p, :k]
    d1 = int((c1[p, :n1[p]].reshape(n1[p], -1) == -1).all(axis=1).sum()); d2 = int((c2[p, :n2[p]].reshape(n2[p], -1) == -1).all(axis=1).sum())
    sub_o = oc[:n1[p], :n2[p]]; sub_g = cg[:n1[p], :n2[p]]
    fin = np.isfinite(sub_o)
    out.append(dict(p=int(p), n1=int(n1[p]), n2=int(n2[p]), dummies=(d1, d2),
                    total_oracle_assign_on_oracle_cost=float(sub_o[a_r, a_c].sum()), total_gpu_assign_on_oracle_cost=float(sub_o[b_r, b_c].sum()),
                    total_oracle_assign_on_gpu_cost=float(sub_g[a_r, a_c].sum()), total_gpu_assign_on_gpu_cost=float(sub_g[b_r, b_c].sum()),
                    max_cost_relerr=float(np.abs(sub_o[fin] - sub_g[fin]).max() / max(np.abs(sub_o[fin]).max(), 1)),
                    max_cost=float(sub_o[fin].max()), n_diff=int((a_c != b_c).sum()),
                    rows_diff=[int(x) for x in np.where(a_c != b_c)[0][:6]]))
print(json.dumps(out, indent=1))
