#!/usr/bin/env python
"""CPU prototype (numpy) of k_ap3's bounding loop -- dual block-coordinate descent, then the
deflected-subgradient phase -- used to choose AP3_SG_ROUND / AP3_SG_PATIENCE / AP3_SG_LAMBDA /
AP3_SG_DEFLECT in csrc/ptk_ap3.cuh without spending GPU time.  It is a design tool, not an oracle and
not on any product path: its 2-D assignment breaks ties differently from the kernel, so iteration
counts differ while the conclusions (which schedules converge on every stored cube) carry over.

    python tools/ap3_proto.py "patience=30,lam0=2.0,T=300,max_sg=1200,defl=0.7"

runs every cube of tests/golden/ap3_rod_optima.json and reports: optimum found / bound closed.
Findings recorded in profiles/README.md (K10 section).  Test tooling: imports oracle/ (through
tools/ap3_check.rod_weights)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def lsap_duals(cost):
    """Square min-cost assignment with duals (cost - u - v >= 0, = 0 on the assignment)."""
    n = cost.shape[0]
    u = np.zeros(n + 1)
    v = np.zeros(n + 1)
    p = np.zeros(n + 1, int)
    way = np.zeros(n + 1, int)
    a = np.zeros((n + 1, n + 1))
    a[1:, 1:] = cost
    for i in range(1, n + 1):
        p[0] = i
        j0 = 0
        minv = np.full(n + 1, np.inf)
        used = np.zeros(n + 1, bool)
        while True:
            used[j0] = True
            i0 = p[j0]
            cur = a[i0, 1:] - u[i0] - v[1:]
            upd = (~used[1:]) & (cur < minv[1:])
            minv[1:][upd] = cur[upd]
            way[1:][upd] = j0
            cand = np.where(~used[1:], minv[1:], np.inf)
            j1 = int(np.argmin(cand)) + 1
            delta = cand[j1 - 1]
            um = used.copy()
            u[p[um]] += delta
            v[um] -= delta
            minv[~um] -= delta
            j0 = j1
            if p[j0] == 0:
                break
        while True:
            j1 = way[j0]
            p[j0] = p[j1]
            j0 = j1
            if j0 == 0:
                break
    col4row = np.zeros(n, int)
    for j in range(1, n + 1):
        col4row[p[j] - 1] = j - 1
    return col4row, u[1:].copy(), v[1:].copy()


def reduce_axis(w, y, r):
    """c[p,q] = max(0, max_r (w - y_r)) and the arg-max (-1 where nothing is positive)."""
    sh = [1, 1, 1]
    sh[r] = -1
    z = w - y.reshape(sh)
    c = z.max(axis=r)
    arg = np.where(c > 0, z.argmax(axis=r), -1)
    return np.maximum(c, 0), arg


def complete(w, r, pairs):
    """Value of the pairs completed over axis r by a second assignment (the kernel's lower bound)."""
    from scipy.optimize import linear_sum_assignment as lsa

    wm = np.moveaxis(w, r, 0)
    M = np.array([np.maximum(wm[:, p, q], 0) for p, q in pairs])
    ri, ci = lsa(M, maximize=True)
    return M[ri, ci].sum()


def run(w, max_bcd=60, T=300, max_sg=1200, patience=30, lam0=2.0, defl=0.7, relstop=1e-10):
    n = w.shape
    y = [np.zeros(n[0]), np.zeros(n[1]), np.zeros(n[2])]
    st = dict(ub=np.inf, best=0.0, ybest=None, since=0)
    eps = 16 * 2.2e-16 * min(n) * max(np.abs(w).max(), 1e-300)

    def step(r):
        c, arg = reduce_axis(w, y[r], r)
        m = max(c.shape)
        cp = np.zeros((m, m))
        cp[: c.shape[0], : c.shape[1]] = c
        col, u, v = lsap_duals(-cp)
        a, b = -u, -v
        amin = a.min()
        pq = [x for x in range(3) if x != r]
        y[pq[0]] = np.maximum(0, a - amin)[: n[pq[0]]]
        y[pq[1]] = np.maximum(0, b + amin)[: n[pq[1]]]
        L = sum(x.sum() for x in y)
        if L < st["ub"] - 1e-13 * abs(L):
            st["ub"] = L
            st["ybest"] = [x.copy() for x in y]
            st["since"] = 0
        else:
            st["since"] += 1
        pairs = [(i, col[i]) for i in range(c.shape[0]) if col[i] < c.shape[1]]
        st["best"] = max(st["best"], complete(w, r, pairs))
        used = [arg[i, j] for i, j in pairs if c[i, j] > 0 and arg[i, j] >= 0]
        return L, 1 - np.bincount(np.array(used, int), minlength=n[r])

    hist = []
    for s_ in range(max_bcd):
        step(s_ % 3)
        hist.append(st["ub"])
        if st["ub"] - st["best"] <= eps:
            return "descent", s_, st
        if s_ >= 5 and hist[-4] - st["ub"] <= 1e-13 * abs(st["ub"]):
            break
    r, lam, d = 0, lam0, None
    st["since"] = 0
    for it in range(max_sg):
        if it and it % T == 0:
            r, lam, d = (r + 1) % 3, lam0, None
            st["since"] = 0
            for x in range(3):
                y[x] = st["ybest"][x].copy()
        L, g = step(r)
        if st["ub"] - st["best"] <= eps:
            return "subgradient", it, st
        if st["ub"] - st["best"] <= relstop * abs(st["ub"]):
            return "search", it, st  # small enough for the kernel's reduced-cost search
        if st["since"] >= patience:
            lam *= 0.5
            st["since"] = 0
            y[r] = st["ybest"][r].copy()
            d = None
            continue
        d = g + defl * d if (defl > 0 and d is not None) else g.astype(float)
        gg = (d * d).sum()
        if gg == 0:
            lam *= 0.5
            continue
        y[r] = np.maximum(0, y[r] - lam * (L - st["best"]) / gg * d)
    return "gap", it, st


def main():
    from tools.ap3_check import rod_weights

    kw = eval("dict(" + (sys.argv[1] if len(sys.argv) > 1 else "") + ")")
    with open(os.path.join(ROOT, "tests", "golden", "ap3_rod_optima.json")) as f:
        table = json.load(f)
    found = closed = 0
    for name, rec in table.items():
        w = rod_weights(**rec["args"])
        opt = rec["milp"]
        t = time.time()
        how, it, st = run(w, **kw)
        ok = abs(st["best"] - opt) <= 1e-9 * abs(opt)
        found += ok
        closed += how != "gap"
        print(f"{name:10s} {how:11s} it {it:4d} best-opt {st['best'] - opt:+.2e} ub-opt {st['ub'] - opt:+.2e} {time.time() - t:.1f}s", flush=True)
    print(json.dumps(dict(args=kw, cubes=len(table), optimum_found=int(found), bound_closed=int(closed))))


if __name__ == "__main__":
    main()
