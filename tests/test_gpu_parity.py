"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI
(particletracking_b200.engine -> libptk.so), against the oracle on the same seeded inputs
and against the committed goldens made by the unmodified reference.

Bars (BASELINE.json north_star): assignments / track ids bit-exact; coordinates and
reprojection errors within TOL = 1e-9 norm-relative (max|a-b| / max(max|b|, 1))."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.fixture(scope="module")
def eng():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from particletracking_b200 import engine

    return engine


@pytest.fixture(scope="module")
def crig(rig):
    from particletracking_b200 import rig as rigmod

    return rigmod.rig_from_dicts(rig["calibration"], rig["transformation"])


def dev(a, dtype=None):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def run_match(eng, crig, cam1, cam2, n1, n2, **kw):
    import torch

    r = eng.match_batch(crig, dev(cam1), dev(cam2), dev(n1, torch.int32), dev(n2, torch.int32), **kw)
    torch.cuda.synchronize()
    return r


# ------------------------------------------------------------------ K0
def test_undistort_bit_exact(eng, crig, rig):
    """K0 follows OpenCV's unfused operation order: identical bits to cv2 (via the C oracle,
    itself bit-exact vs cv2 in tests/test_oracle_cpu.py), including N == 1 and ragged n."""
    import torch
    from oracle import c_oracle as co

    rng = np.random.default_rng(0)
    P, Nmax = 40, 33
    cam1 = rng.uniform(50, 1250, size=(P, Nmax, 2, 2))
    cam2 = rng.uniform(50, 1250, size=(P, Nmax, 2, 2))
    n1 = rng.integers(0, Nmax + 1, size=P).astype(np.int32)
    n2 = rng.integers(0, Nmax + 1, size=P).astype(np.int32)
    n1[:3] = [1, 2, Nmax]
    n2[:3] = [1, 1, Nmax]
    cam1[5, 0] = -1.0  # dummy rod
    u1, u2 = eng.undistort_batch(crig, dev(cam1), dev(cam2), dev(n1), dev(n2))
    u1, u2 = u1.cpu().numpy(), u2.cpu().numpy()
    cal = rig["calibration"]
    for p in range(P):
        np.testing.assert_array_equal(u1[p, : n1[p]], co.undistort_rods(cal["CM1"], cal["dist1"], cam1[p, : n1[p]]))
        np.testing.assert_array_equal(u2[p, : n2[p]], co.undistort_rods(cal["CM2"], cal["dist2"], cam2[p, : n2[p]]))


# ------------------------------------------------------------------ K1
@pytest.mark.parametrize("agg_sum", [False, True])
def test_cost_matrix_and_combos_vs_oracle(eng, rig, agg_sum):
    import torch
    from oracle import c_oracle as co
    from particletracking_b200 import rig as rigmod
    from particletracking_b200 import synthetic as syn

    r = rigmod.rig_from_dicts(rig["calibration"], rig["transformation"], agg_sum=agg_sum)
    data = syn.make_stereo(2, 2, 20, seed=31, p_drop=0.15, p_dummy=0.05)
    c1, c2, n1, n2 = data.problems()
    cost, choice, pw, re = eng.cost_batch(r, dev(c1), dev(c2), dev(n1), dev(n2), want_combos=True)
    cost, choice, pw, re = (x.cpu().numpy() for x in (cost, choice, pw, re))
    for p in range(len(n1)):
        o = co.cost_problem(r, c1[p, : n1[p]], c2[p, : n2[p]], want_combos=True)
        a, b = n1[p], n2[p]
        assert rel_err(cost[p, :a, :b], o["cost"]) < TOL
        np.testing.assert_array_equal(choice[p, :a, :b].astype(bool), o["choice"])
        assert rel_err(re[p, :a, :b], o["rep_errs"]) < TOL
        # 3D: norm-relative per point
        d = np.abs(pw[p, :a, :b] - o["p_world"]).max(axis=-1)
        s = np.maximum(np.abs(o["p_world"]).max(axis=-1), 1.0)
        assert (d / s).max() < TOL


# ------------------------------------------------------------------ K2
def test_lsap_identical_to_scipy(eng):
    import torch
    from scipy.optimize import linear_sum_assignment
    from test_oracle_cpu import _lsap_cases

    cases = _lsap_cases()
    R = max(c.shape[0] for c in cases)
    Cc = max(c.shape[1] for c in cases)
    P = len(cases)
    cost = np.full((P, R, Cc), np.nan)
    nr = np.zeros(P, dtype=np.int32)
    nc = np.zeros(P, dtype=np.int32)
    for p, c in enumerate(cases):
        cost[p, : c.shape[0], : c.shape[1]] = c
        nr[p], nc[p] = c.shape
    row, col, n_out, status = eng.lsap_batch(dev(cost), dev(nr), dev(nc))
    row, col, n_out, status = (x.cpu().numpy() for x in (row, col, n_out, status))
    assert (status == 0).all()
    for p, c in enumerate(cases):
        er, ec = linear_sum_assignment(c)
        k = len(er)
        assert n_out[p] == k
        np.testing.assert_array_equal(row[p, :k], er)
        np.testing.assert_array_equal(col[p, :k], ec)
        assert (row[p, k:] == -1).all() and (col[p, k:] == -1).all()


def test_lsap_negative_and_mixed_sign_costs(eng):
    """The solver's keys are the raw bits of the shortest-path cost (valid for non-negative values); a negative
    candidate is caught by one compare and redone with the order-preserving transform.  Reprojection errors
    never take that path, a caller's own matrices may: mixed-sign, all-negative, heavily tied integers around
    zero, signed zeros -- assignments must equal SciPy's."""
    from scipy.optimize import linear_sum_assignment

    rng = np.random.default_rng(11)
    cases = []
    for _ in range(60):
        nr, nc = rng.integers(1, 20, size=2)
        cases.append(rng.normal(size=(nr, nc)))
        cases.append(-rng.random((nr, nc)))
        cases.append(rng.integers(-3, 4, size=(nr, nc)).astype(float))
        m = rng.integers(-1, 2, size=(nr, nc)).astype(float)
        m[m == 0] = np.where(rng.random((m == 0).sum()) < 0.5, 0.0, -0.0)
        cases.append(m)
    for n in (32, 48, 64, 100):
        cases.append(rng.normal(size=(n, n)) * 1e3)
        cases.append(rng.integers(-5, 6, size=(n, n + 3)).astype(float))
        cases.append(rng.normal(size=(n + 4, n)))
        d1, d2 = rng.normal(size=n), rng.normal(size=n)
        cases.append(d1[:, None] + d2[None, :])
    R = max(c.shape[0] for c in cases)
    Cc = max(c.shape[1] for c in cases)
    cost = np.full((len(cases), R, Cc), np.nan)
    nr = np.zeros(len(cases), dtype=np.int32)
    nc = np.zeros(len(cases), dtype=np.int32)
    for p, c in enumerate(cases):
        cost[p, : c.shape[0], : c.shape[1]] = c
        nr[p], nc[p] = c.shape
    row, col, n_out, status = (x.cpu().numpy() for x in eng.lsap_batch(dev(cost), dev(nr), dev(nc)))
    assert (status == 0).all()
    for p, c in enumerate(cases):
        er, ec = linear_sum_assignment(c)
        k = len(er)
        assert n_out[p] == k
        np.testing.assert_array_equal(row[p, :k], er, err_msg=f"case {p} {c.shape}")
        np.testing.assert_array_equal(col[p, :k], ec, err_msg=f"case {p} {c.shape}")


def test_lsap_large_and_status_codes(eng):
    import torch
    from scipy.optimize import linear_sum_assignment

    rng = np.random.default_rng(5)
    # N = 256 (BASELINE config 5 upper end), random + tied + separable
    mats = [rng.random((256, 256)), rng.integers(0, 8, size=(256, 256)).astype(float),
            rng.random(256)[:, None] + rng.random(256)[None, :], rng.random((200, 256)), rng.random((256, 190))]
    cost = np.full((len(mats) + 3, 256, 256), 0.5)
    nr = np.full(len(cost), 4, dtype=np.int32)
    nc = np.full(len(cost), 4, dtype=np.int32)
    for p, m in enumerate(mats):
        cost[p, : m.shape[0], : m.shape[1]] = m
        nr[p], nc[p] = m.shape
    k = len(mats)
    cost[k, 1, 2] = np.nan            # -> INVALID (scipy: ValueError)
    cost[k + 1, :4, :4] = np.inf      # -> INFEASIBLE
    nr[k + 2] = 0                     # -> EMPTY
    row, col, n_out, status = (x.cpu().numpy() for x in eng.lsap_batch(dev(cost), dev(nr), dev(nc)))
    np.testing.assert_array_equal(status[k:], [1, 2, 3])
    assert (n_out[k:] == 0).all() and (col[k:] == -1).all()
    for p, m in enumerate(mats):
        er, ec = linear_sum_assignment(m)
        np.testing.assert_array_equal(row[p, : len(er)], er)
        np.testing.assert_array_equal(col[p, : len(ec)], ec)


# ------------------------------------------------------------------ K0..K3 vs goldens (reference outputs)
@pytest.mark.parametrize("name", ["c1_3x12.npz", "ragged_2x16.npz", "tiny_n1.npz", "tiny_n2.npz", "n32.npz"])
def test_match_batch_vs_reference_goldens(eng, crig, name):
    g = np.load(os.path.join(GOLDEN, name))
    F, C, Nm = g["cam1"].shape[:3]
    r = run_match(eng, crig, g["cam1"].reshape(F * C, Nm, 2, 2), g["cam2"].reshape(F * C, Nm, 2, 2),
                  g["n1"].ravel(), g["n2"].ravel(), want_cost=True)
    rows, mc, row, col, n_out, status, cost = (x.cpu().numpy() for x in
                                               (r.rows, r.match_cost, r.row_ind, r.col_ind, r.n_out, r.status, r.cost))
    assert (status == 0).all()
    np.testing.assert_array_equal(n_out, g["n_out"].ravel())
    np.testing.assert_array_equal(row, g["row_ind"].reshape(F * C, Nm))   # bit-exact assignments
    np.testing.assert_array_equal(col, g["col_ind"].reshape(F * C, Nm))
    for p in range(F * C):
        f, c = divmod(p, C)
        a, b, n = g["n1"][f, c], g["n2"][f, c], g["n_out"][f, c]
        np.testing.assert_array_equal(rows[p, :n, 10:], g["rows"][f, c, :n, 10:])  # 2D columns exact
        assert np.isnan(rows[p, n:]).all() and np.isnan(mc[p, n:]).all()
        for k in range(n):  # 3D end points / centre: norm-relative per point
            for s in (slice(0, 3), slice(3, 6), slice(6, 9)):
                assert rel_err(rows[p, k, s], g["rows"][f, c, k, s]) < TOL
        assert rel_err(rows[p, :n, 9], g["lengths"][f, c, :n]) < TOL
        assert rel_err(cost[p, :a, :b], g["cost"][f, c, :a, :b]) < TOL
        assert rel_err(mc[p, :n], g["costs"][f, c, :n]) < TOL


@pytest.mark.parametrize("n,drop,dummy,seed", [(12, 0.0, 0.0, 1), (32, 0.0, 0.0, 2), (24, 0.2, 0.1, 3),
                                               (64, 0.03, 0.02, 4), (7, 0.5, 0.0, 5)])
def test_match_batch_vs_c_oracle(eng, crig, n, drop, dummy, seed):
    from oracle import c_oracle as co
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(6, 3, n, seed=seed, p_drop=drop, p_dummy=dummy, sigma_px=0.5)
    c1, c2, n1, n2 = data.problems()
    o = co.match_batch(crig, c1, c2, n1, n2, want_cost=True)
    r = run_match(eng, crig, c1, c2, n1, n2, want_cost=True, max_repr_err=5.0)
    np.testing.assert_array_equal(r.status.cpu().numpy(), o["status"])
    np.testing.assert_array_equal(r.n_out.cpu().numpy(), o["n_out"])
    np.testing.assert_array_equal(r.row_ind.cpu().numpy(), o["row_ind"])
    np.testing.assert_array_equal(r.col_ind.cpu().numpy(), o["col_ind"])
    rows, mc, keep = r.rows.cpu().numpy(), r.match_cost.cpu().numpy(), r.keep.cpu().numpy()
    for p in range(len(n1)):
        k = o["n_out"][p]
        np.testing.assert_array_equal(rows[p, :k, 10:], o["rows"][p, :k, 10:])
        assert rel_err(rows[p, :k, :10], o["rows"][p, :k, :10]) < TOL
        assert rel_err(mc[p, :k], o["match_cost"][p, :k]) < TOL
        # threshold extension: mask over the returned costs, nothing else changes
        np.testing.assert_array_equal(keep[p, :k].astype(bool), mc[p, :k] <= 5.0)
        assert (keep[p, k:] == 0).all()


def test_match_loop_index_quirk_n1_gt_n2(eng, crig, rig):
    """N1 > N2: output row m pairs cam1 rod m (not row_ind[m]) with cam2 rod col_ind[m], while
    costs[m] = cost[row_ind[m], col_ind[m]] -- checked against the numpy port,
    which was checked against the reference."""
    from oracle import port
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(1, 1, 14, seed=77)
    c1, c2 = data.cam1[0, 0], data.cam2[0, 0, :9]
    n1, n2 = np.array([14], dtype=np.int32), np.array([9], dtype=np.int32)
    cam2 = np.full((1, 14, 2, 2), np.nan)
    cam2[0, :9] = c2
    r = run_match(eng, crig, c1[None], cam2, n1, n2)
    ref = port.match_arrays(c1, c2, rig["calibration"], rig["P1"], rig["P2"], rig["rot"], rig["trans"],
                            rig["r1"], rig["r2"], rig["t1"], rig["t2"])
    assert not np.array_equal(ref["row_ind"], np.arange(9))  # the quirk is actually exercised
    np.testing.assert_array_equal(r.row_ind.cpu().numpy()[0, :9], ref["row_ind"])
    np.testing.assert_array_equal(r.col_ind.cpu().numpy()[0, :9], ref["col_ind"])
    rows = r.rows.cpu().numpy()[0, :9]
    np.testing.assert_array_equal(rows[:, 10:], ref["rows"][:, 10:])
    assert rel_err(rows[:, :10], ref["rows"][:, :10]) < TOL
    assert rel_err(r.match_cost.cpu().numpy()[0, :9], ref["costs"]) < TOL


def test_match_nan_rod_reports_invalid(eng, crig):
    """A partially-NaN rod survives the reference's filter and makes scipy raise ValueError
    (SURVEY App. A.1); the batch reports it per problem and leaves the others intact."""
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(1, 3, 8, seed=9)
    c1, c2, n1, n2 = data.problems()
    c1 = c1.copy()
    c1[1, 2, 0, 0] = np.nan
    n1 = n1.copy()
    n1[2] = 0
    r = run_match(eng, crig, c1, c2, n1, n2)
    np.testing.assert_array_equal(r.status.cpu().numpy(), [0, 1, 3])
    np.testing.assert_array_equal(r.n_out.cpu().numpy(), [8, 0, 0])
    assert np.isnan(r.rows.cpu().numpy()[1:]).all()


# ------------------------------------------------------------------ K4 / K5
@pytest.mark.parametrize("name", ["track_12.npz", "track_25.npz"])
def test_tracking_vs_reference_goldens_bit_exact(eng, name):
    g = np.load(os.path.join(GOLDEN, name))
    col, tot, status, cost = eng.track_batch(dev(g["ends"][None]), want_cost=True)
    assert (status.cpu().numpy() == 0).all()
    np.testing.assert_array_equal(cost.cpu().numpy()[0], g["cost"])        # numpy-order arithmetic: identical bits
    np.testing.assert_array_equal(col.cpu().numpy()[0], g["col_ind"])      # bit-exact assignments
    assert rel_err(tot.cpu().numpy()[0], g["total_cost"]) < 1e-13


@pytest.mark.parametrize("C,F,N", [(2, 9, 16), (8, 40, 32), (1, 70, 64), (3, 2, 5), (2, 33, 7), (2, 65, 3)])
def test_tracking_and_chain_vs_c_oracle(eng, C, F, N):
    import torch
    from oracle import c_oracle as co

    rng = np.random.default_rng(C * 1000 + F)
    base = rng.normal(size=(C, 1, N, 2, 3)) * 30
    ends = base + np.cumsum(rng.normal(size=(C, F, N, 2, 3)) * 0.2, axis=1)
    o = co.track_batch(ends, want_cost=True)
    col, tot, status, cost = eng.track_batch(dev(ends), want_cost=True)
    np.testing.assert_array_equal(cost.cpu().numpy(), o["cost"])
    np.testing.assert_array_equal(col.cpu().numpy(), o["col_ind"])
    assert rel_err(tot.cpu().numpy(), o["total_cost"]) < 1e-13
    # same without the optional cost output (workspace path)
    col2, tot2, _, _ = eng.track_batch(dev(ends))
    np.testing.assert_array_equal(col2.cpu().numpy(), o["col_ind"])
    # chain pass
    tid = eng.chain_tracks(col).cpu().numpy()
    np.testing.assert_array_equal(tid, co.chain_tracks(o["col_ind"]))
    ids0 = np.stack([rng.permutation(N) for _ in range(C)]).astype(np.int32)
    tid2 = eng.chain_tracks(col, dev(ids0)).cpu().numpy()
    np.testing.assert_array_equal(tid2, co.chain_tracks(o["col_ind"], ids0))


def test_chain_long_sequence_properties(eng):
    """Size-independent property at a length the oracle loop would be slow for: ids are a
    permutation in every frame and following col_ind forward preserves the id."""
    import torch

    rng = np.random.default_rng(3)
    C, F, N = 4, 5000, 64
    col = np.stack([np.stack([rng.permutation(N) for _ in range(F - 1)]) for _ in range(C)]).astype(np.int32)
    tid = eng.chain_tracks(dev(col)).cpu().numpy()
    assert (np.sort(tid, axis=-1) == np.arange(N)).all()
    nxt = np.take_along_axis(tid[:, 1:], col, axis=2)  # ids[f+1][col[f][a]]
    np.testing.assert_array_equal(nxt, tid[:, :-1])


# ------------------------------------------------------------------ K6
def test_create_weights_vs_golden_and_oracle(eng):
    from oracle import c_oracle as co

    g = np.load(os.path.join(GOLDEN, "weights.npz"))
    w, ch = eng.create_weights(dev(g["p_3D"]), dev(g["p_prev"]), dev(g["rep_errs"]))
    np.testing.assert_array_equal(ch.cpu().numpy(), g["choices"])
    assert rel_err(w.cpu().numpy(), g["weights"]) < 1e-12
    rng = np.random.default_rng(8)
    p3, prev, re = rng.normal(size=(33, 31, 4, 3)) * 20, rng.normal(size=(29, 2, 3)) * 20, rng.random((33, 31, 4, 2))
    w, ch = eng.create_weights(dev(p3), dev(prev), dev(re))
    ow, och = co.create_weights(p3, prev, re)
    np.testing.assert_array_equal(ch.cpu().numpy(), och)
    assert rel_err(w.cpu().numpy(), ow) < 1e-12


# ------------------------------------------------------------------ host pipeline (e2e entry point)
def test_host_pipeline_matches_device_api(eng, crig):
    import torch
    from oracle import c_oracle as co
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(300, 4, 16, seed=12)   # > 1 chunk (2048 problems per chunk) would need more; force via frames
    pipe = eng.HostPipeline(0)
    out = pipe.reconstruct(crig, data.cam1, data.cam2, data.n1, data.n2, track=True, max_repr_err=5.0)
    c1, c2, n1, n2 = data.problems()
    r = run_match(eng, crig, c1, c2, n1, n2)
    F, C, N = 300, 4, 16
    np.testing.assert_array_equal(out["col_ind"].numpy().reshape(F * C, N), r.col_ind.cpu().numpy())
    np.testing.assert_array_equal(out["rows"].numpy().reshape(F * C, N, 18), r.rows.cpu().numpy())  # same kernels, same bits
    assert (out["status"].numpy() == 0).all()
    # tracking on the pipeline's own rows == device API on the same ends
    ends = out["rows"].numpy()[..., :6].reshape(F, C, N, 2, 3).transpose(1, 0, 2, 3, 4).copy()
    o = co.track_batch(ends)
    np.testing.assert_array_equal(out["track_col"].numpy(), o["col_ind"])
    np.testing.assert_array_equal(out["track_id"].numpy(), co.chain_tracks(o["col_ind"]))
    assert rel_err(out["track_total"].numpy(), o["total_cost"]) < 1e-13
    pipe.close()


# ------------------------------------------------------------------ BASELINE config 2 at full size, every problem
def test_full_size_c2_every_problem_vs_oracle(eng, crig):
    """1000 frames x 8 colours x 32 rods (the benchmarked workload): ALL 8000 problems against the C
    oracle (a couple of seconds of OpenMP), the tracking / chain pass on the GPU's own rows against the
    oracle, the size-independent properties, and the benchmarked host entry point (ptk_reconstruct_host,
    its multi-chunk path at this size) against the device API bit for bit."""
    import torch
    from oracle import c_oracle as co
    from particletracking_b200 import synthetic as syn

    F, C, N = 1000, 8, 32
    data = syn.make_stereo(F, C, N, seed=2)
    c1, c2, n1, n2 = data.problems()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    m, tcol, ttot, tid, tst = eng.reconstruct_device(crig, t(c1), t(c2), t(n1), t(n2), F, C, track=True)
    torch.cuda.synchronize()
    col = m.col_ind.cpu().numpy()
    rows = m.rows.cpu().numpy()
    assert (m.status.cpu().numpy() == 0).all() and (m.n_out.cpu().numpy() == N).all() and (tst.cpu().numpy() == 0).all()
    assert (np.sort(col, axis=1) == np.arange(N)).all()
    np.testing.assert_allclose(rows[..., 6:9], (rows[..., 0:3] + rows[..., 3:6]) / 2, rtol=0, atol=1e-12)
    np.testing.assert_allclose(rows[..., 9], np.linalg.norm(rows[..., 3:6] - rows[..., 0:3], axis=-1), rtol=1e-14)
    truth = data.truth_perm.reshape(-1, N)
    assert (np.take_along_axis(truth, col, axis=1) == np.arange(N)).mean() > 0.99
    # every problem against the oracle
    o = co.match_batch(crig, c1, c2, n1, n2, nthreads=0)
    np.testing.assert_array_equal(col, o["col_ind"])
    np.testing.assert_array_equal(m.row_ind.cpu().numpy(), o["row_ind"])
    assert rel_err(rows[..., :10], o["rows"][..., :10]) < TOL
    np.testing.assert_array_equal(rows[..., 10:], o["rows"][..., 10:])
    assert rel_err(m.match_cost.cpu().numpy(), o["match_cost"]) < TOL
    # tracking + chain on the rows the GPU tracked
    ends = rows[..., :6].reshape(F, C, N, 2, 3).transpose(1, 0, 2, 3, 4).copy()
    ot = co.track_batch(ends)
    np.testing.assert_array_equal(tcol.cpu().numpy(), ot["col_ind"])
    np.testing.assert_array_equal(tid.cpu().numpy(), co.chain_tracks(ot["col_ind"]))
    assert rel_err(ttot.cpu().numpy(), ot["total_cost"]) < 1e-13
    # the host entry point at the benchmarked size (4 chunks of 250 frames): same bits as the device API
    pipe = eng.HostPipeline(0)
    h = pipe.reconstruct(crig, data.cam1, data.cam2, data.n1, data.n2, track=True)
    np.testing.assert_array_equal(h["rows"].numpy().reshape(F * C, N, 18), rows)
    np.testing.assert_array_equal(h["col_ind"].numpy().reshape(F * C, N), col)
    np.testing.assert_array_equal(h["match_cost"].numpy().reshape(F * C, N), m.match_cost.cpu().numpy())
    np.testing.assert_array_equal(h["track_col"].numpy(), tcol.cpu().numpy())
    np.testing.assert_array_equal(h["track_id"].numpy(), tid.cpu().numpy())
    np.testing.assert_array_equal(h["track_total"].numpy(), ttot.cpu().numpy())
    pipe.close()


# ------------------------------------------------------------------ large N (BASELINE config 5: 8 -> 256 rods)
@pytest.mark.parametrize("n,P", [(8, 40), (128, 6), (256, 2)])
def test_match_large_n_vs_c_oracle(eng, crig, n, P):
    from oracle import c_oracle as co
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(P, 1, n, seed=500 + n, p_drop=0.02 if n > 8 else 0.0)
    c1, c2, n1, n2 = data.problems()
    o = co.match_batch(crig, c1, c2, n1, n2)
    r = run_match(eng, crig, c1, c2, n1, n2)
    np.testing.assert_array_equal(r.status.cpu().numpy(), o["status"])
    np.testing.assert_array_equal(r.col_ind.cpu().numpy(), o["col_ind"])
    np.testing.assert_array_equal(r.row_ind.cpu().numpy(), o["row_ind"])
    rows = r.rows.cpu().numpy()
    for p in range(P):
        k = o["n_out"][p]
        assert rel_err(rows[p, :k, :10], o["rows"][p, :k, :10]) < TOL
        np.testing.assert_array_equal(rows[p, :k, 10:], o["rows"][p, :k, 10:])
    assert rel_err(r.match_cost.cpu().numpy()[:, :8], o["match_cost"][:, :8]) < TOL


# ------------------------------------------------------------------ BASELINE config 3 (ragged, dummies, threshold), all 10 000 frames
def test_config3_ragged_all_frames_vs_oracle(eng, crig):
    """10 000 frames x 8 colours x 32 nominal rods with noisy / missing detections (N1 != N2), 2 % dummy
    rods, sigma 0.5 px, 5 px threshold: ALL 80 000 problems against the C oracle."""
    from oracle import c_oracle as co
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(10000, 8, 32, seed=3, p_drop=0.05, p_dummy=0.02, sigma_px=0.5)
    c1, c2, n1, n2 = data.problems()
    r = run_match(eng, crig, c1, c2, n1, n2, max_repr_err=5.0)
    n_out = r.n_out.cpu().numpy()
    col, row = r.col_ind.cpu().numpy(), r.row_ind.cpu().numpy()
    mc, keep = r.match_cost.cpu().numpy(), r.keep.cpu().numpy()
    rows = r.rows.cpu().numpy()
    assert (r.status.cpu().numpy() == 0).all()
    np.testing.assert_array_equal(n_out, np.minimum(n1, n2))
    assert (n1 != n2).mean() > 0.5  # the config really is ragged
    np.testing.assert_array_equal(keep.astype(bool), np.nan_to_num(mc, nan=np.inf) <= 5.0)
    o = co.match_batch(crig, c1, c2, n1, n2, nthreads=0)
    np.testing.assert_array_equal(n_out, o["n_out"])
    # Assignments: identical, except where the optimum itself is degenerate.  Dummy detections (-1,-1,-1,-1)
    # are bit-identical rods: swapping the partners of two of them (or leaving the other one unmatched when
    # N1 > N2) costs exactly the same, and which of the equal optima SciPy's tie rule lands on is steered by
    # the last bit of the duals, i.e. of a DLT that differs from the oracle's (and from cv2's) in the 15th
    # digit.  Measured: 1 of 80 000 problems.
    differs = np.where((col != o["col_ind"]).any(axis=1) | (row != o["row_ind"]).any(axis=1))[0]
    assert len(differs) <= 8, f"{len(differs)} problems differ from the oracle"
    for p in differs:
        k = n_out[p]
        oc = co.match_batch(crig, c1[p:p + 1], c2[p:p + 1], n1[p:p + 1], n2[p:p + 1], want_cost=True)["cost"][0]
        t_o = oc[o["row_ind"][p, :k], o["col_ind"][p, :k]].sum()
        t_g = oc[row[p, :k], col[p, :k]].sum()
        assert abs(t_o - t_g) <= 1e-12 * abs(t_o), "a different assignment must be an equally optimal one"
        assert len(set(col[p, :k])) == k and len(set(row[p, :k])) == k
        pairs_o = set(zip(o["row_ind"][p, :k].tolist(), o["col_ind"][p, :k].tolist()))
        r1 = c1[p, :n1[p]].reshape(n1[p], -1)
        r2 = c2[p, :n2[p]].reshape(n2[p], -1)
        for a, b in zip(row[p, :k].tolist(), col[p, :k].tolist()):
            if (a, b) in pairs_o:
                continue
            # the rods involved are exact duplicates of another rod (in camera 1 or in camera 2)
            assert (r1 == r1[a]).all(axis=1).sum() > 1 or (r2 == r2[b]).all(axis=1).sum() > 1
    same = np.ones(len(n1), dtype=bool)
    same[differs] = False
    valid = (o["row_ind"] >= 0) & same[:, None]
    assert rel_err(np.where(valid, mc, 0.0), np.where(valid, o["match_cost"], 0.0)) < TOL
    v3 = valid[..., None]
    assert rel_err(np.where(v3, rows[..., :10], 0.0), np.where(v3, o["rows"][..., :10], 0.0)) < TOL
    np.testing.assert_array_equal(np.where(v3, rows[..., 10:], 0.0), np.where(v3, o["rows"][..., 10:], 0.0))
    assert np.isnan(rows[o["row_ind"] < 0]).all()  # the padding is defined (NaN), as documented


# ------------------------------------------------------------------ BASELINE config 4 shape (8 colours x 64 rods, match + track), 2 000 frames
def test_config4_shape_match_and_track(eng, crig):
    import torch
    from oracle import c_oracle as co
    from particletracking_b200 import synthetic as syn

    F, C, N = 2000, 8, 64
    data = syn.make_stereo(F, C, N, seed=4)
    c1, c2, n1, n2 = data.problems()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    m, col, tot, tid, st = eng.reconstruct_device(crig, t(c1), t(c2), t(n1), t(n2), F, C, track=True)
    torch.cuda.synchronize()
    assert (m.status.cpu().numpy() == 0).all() and (st.cpu().numpy() == 0).all()
    o = co.match_batch(crig, c1, c2, n1, n2, nthreads=0)
    np.testing.assert_array_equal(m.col_ind.cpu().numpy(), o["col_ind"])
    rows = m.rows.cpu().numpy()
    assert rel_err(rows[..., :10], o["rows"][..., :10]) < TOL
    np.testing.assert_array_equal(rows[..., 10:], o["rows"][..., 10:])
    # tracking on the SAME 3D input the GPU tracked (its own rows): bit-exact vs the oracle
    ends = rows[..., :6].reshape(F, C, N, 2, 3).transpose(1, 0, 2, 3, 4).copy()
    ot = co.track_batch(ends, nthreads=0)
    np.testing.assert_array_equal(col.cpu().numpy(), ot["col_ind"])
    tid_o = co.chain_tracks(ot["col_ind"])
    np.testing.assert_array_equal(tid.cpu().numpy(), tid_o)
    assert (np.sort(tid_o, axis=-1) == np.arange(N)).all()
    assert rel_err(tot.cpu().numpy(), ot["total_cost"]) < 1e-13


def test_device_pipeline_single_call_and_cuda_graph(eng, crig):
    """ptk_reconstruct_device == the composition of the separate entry points (same kernels,
    same bits); the call only enqueues kernels, so it can be captured in a CUDA graph."""
    import torch
    from particletracking_b200 import synthetic as syn

    F, C, N = 50, 4, 16
    data = syn.make_stereo(F, C, N, seed=21)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    c1, c2, n1, n2 = (t(x) for x in data.problems())
    m, col, tot, tid, st = eng.reconstruct_device(crig, c1, c2, n1, n2, F, C, track=True)
    pipe = eng.DevicePipeline(crig, F, C, N, track=True, max_repr_err=3.0)
    pipe.run(c1, c2, n1, n2)
    torch.cuda.synchronize()
    for a, b in ((pipe.rows, m.rows), (pipe.col_ind, m.col_ind), (pipe.match_cost, m.match_cost),
                 (pipe.track_col, col), (pipe.track_id, tid), (pipe.track_total, tot)):
        np.testing.assert_array_equal(a.cpu().numpy(), b.cpu().numpy())
    np.testing.assert_array_equal(pipe.keep.cpu().numpy().astype(bool), pipe.match_cost.cpu().numpy() <= 3.0)
    # graph capture + replay on new inputs
    data2 = syn.make_stereo(F, C, N, seed=22)
    d1, d2, _, _ = (t(x) for x in data2.problems())
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        pipe.run(c1, c2, n1, n2)  # warm-up on the side stream
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            pipe.run(c1, c2, n1, n2)
    c1.copy_(d1)
    c2.copy_(d2)
    g.replay()
    torch.cuda.synchronize()
    m2 = eng.match_batch(crig, d1, d2, n1, n2)
    np.testing.assert_array_equal(pipe.rows.cpu().numpy(), m2.rows.cpu().numpy())
    np.testing.assert_array_equal(pipe.col_ind.cpu().numpy(), m2.col_ind.cpu().numpy())


def test_chain_rebase_emulated_ranks(eng):
    """Multi-GPU chain pass on one GPU: W emulated ranks chain their own frame ranges (+ the halo
    pair) from the identity, the boundary maps are 'gathered' and ptk_chain_rebase re-bases the
    local ids -- must equal the single-pass chain over all frames."""
    import torch
    from particletracking_b200.distributed import shard_frames

    rng = np.random.default_rng(9)
    C, F, N, W = 3, 37, 11, 4
    col = np.stack([np.stack([rng.permutation(N) for _ in range(F - 1)]) for _ in range(C)]).astype(np.int32)
    full = eng.chain_tracks(dev(col)).cpu().numpy()
    locals_, maps = [], []
    for r in range(W):
        f0, f1 = shard_frames(F, W, r)
        hi = f1 if r < W - 1 else f1 - 1  # pairs f0..hi-1 (the last one is the halo pair)
        loc = eng.chain_tracks(dev(col[:, f0:hi]))
        locals_.append(loc)
        maps.append(loc[:, -1] if r < W - 1 else torch.arange(N, dtype=torch.int32, device="cuda").expand(C, N))
    maps = torch.stack([m.contiguous() for m in maps]).contiguous()
    out = []
    for r in range(W):
        f0, f1 = shard_frames(F, W, r)
        loc = locals_[r].clone()
        eng.chain_rebase(maps, r, loc)
        out.append(loc[:, : f1 - f0].cpu().numpy())
    np.testing.assert_array_equal(np.concatenate(out, axis=1), full)


def test_full_distortion_model_vs_c_oracle(eng, rig):
    """Rig with tangential, k3 and rational (k4..k6) and thin-prism (s1..s4) coefficients: exercises
    every term of the undistortion / projection formulas."""
    from oracle import c_oracle as co
    from particletracking_b200 import rig as rigmod
    from particletracking_b200 import synthetic as syn

    cal = dict(syn.synthetic_calibration())
    cal["dist1"] = np.array([-0.11, 1.2, 1e-3, -2e-3, 0.3, 0.02, -0.01, 0.005, 1e-3, -1e-3, 2e-3, 5e-4])
    cal["dist2"] = np.array([-0.41, 10.1, -1e-3, 1e-3, 0.0, 0.01, 0.02, 0.0])
    trf = syn.synthetic_transformation()
    r = rigmod.rig_from_dicts(cal, trf)
    data = syn.make_stereo(3, 2, 16, seed=44, calibration=cal)
    c1, c2, n1, n2 = data.problems()
    o = co.match_batch(r, c1, c2, n1, n2, want_cost=True)
    g = run_match(eng, r, c1, c2, n1, n2, want_cost=True)
    np.testing.assert_array_equal(g.col_ind.cpu().numpy(), o["col_ind"])
    assert rel_err(g.cost.cpu().numpy(), o["cost"]) < TOL
    assert rel_err(g.rows.cpu().numpy()[..., :10], o["rows"][..., :10]) < TOL
    u1, u2 = eng.undistort_batch(r, dev(c1), dev(c2), dev(n1), dev(n2))
    for p in range(len(n1)):
        np.testing.assert_array_equal(u1.cpu().numpy()[p], co.undistort_rods(cal["CM1"], cal["dist1"], c1[p]))
        np.testing.assert_array_equal(u2.cpu().numpy()[p], co.undistort_rods(cal["CM2"], cal["dist2"], c2[p]))


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_rigs_vs_c_oracle(eng, seed):
    """Other stereo geometries than the bench rig: random focal lengths, principal points,
    relative pose and distortion.  The host-side static column order of the Jacobi is only a
    guess for these; points that disagree take the in-kernel sorting network -- results must not
    depend on which path ran."""
    from scipy.spatial.transform import Rotation
    from oracle import c_oracle as co
    from particletracking_b200 import rig as rigmod
    from particletracking_b200 import synthetic as syn

    rng = np.random.default_rng(seed)
    trf = syn.synthetic_transformation()
    f1, f2 = rng.uniform(800, 4000, size=2)
    cal = {
        "CM1": np.array([[f1, 0, rng.uniform(300, 900)], [0, f1 * rng.uniform(0.97, 1.03), rng.uniform(250, 700)], [0, 0, 1.0]]),
        "CM2": np.array([[f2, 0, rng.uniform(300, 900)], [0, f2 * rng.uniform(0.97, 1.03), rng.uniform(250, 700)], [0, 0, 1.0]]),
        "dist1": np.array([rng.uniform(-0.2, 0.1), rng.uniform(-0.5, 1.0), rng.uniform(-1e-3, 1e-3), rng.uniform(-1e-3, 1e-3)]),
        "dist2": np.array([rng.uniform(-0.2, 0.1), rng.uniform(-0.5, 1.0), 0.0, 0.0, rng.uniform(-0.2, 0.2)]),
    }
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    R = Rotation.from_rotvec(axis * np.deg2rad(rng.uniform(15, 95))).as_matrix()
    centre_c1 = -trf["rotation"].T @ trf["translation"]            # box centre in camera-1 coordinates
    cal["R"] = R
    cal["T"] = (np.array([0.0, 0.0, rng.uniform(250, 450)]) - R @ centre_c1).reshape(3, 1)
    r = rigmod.rig_from_dicts(cal, trf)
    data = syn.make_stereo(3, 2, 18, seed=100 + seed, p_dummy=0.1, calibration=cal)
    c1, c2, n1, n2 = data.problems()
    o = co.match_batch(r, c1, c2, n1, n2, want_cost=True)
    g = run_match(eng, r, c1, c2, n1, n2, want_cost=True)
    np.testing.assert_array_equal(g.status.cpu().numpy(), o["status"])
    np.testing.assert_array_equal(g.col_ind.cpu().numpy(), o["col_ind"])
    cost = g.cost.cpu().numpy()
    for p in range(len(n1)):
        assert rel_err(cost[p, : n1[p], : n2[p]], o["cost"][p, : n1[p], : n2[p]]) < TOL
    rows = g.rows.cpu().numpy()
    for p in range(len(n1)):
        for k in range(o["n_out"][p]):
            for s in (slice(0, 3), slice(3, 6), slice(6, 9)):
                assert rel_err(rows[p, k, s], o["rows"][p, k, s]) < TOL


@pytest.mark.parametrize("scale", [2e-3, 1e-3])
def test_near_degenerate_rig_takes_jacobi_fallback(eng, scale):
    """A stereo rig whose baseline is shrunk to (almost) nothing: sigma_4/sigma_3 of the DLT matrices
    is above 0.39 for 15 % (scale 2e-3) or 44 % (1e-3) of the end-point combinations: K1's inverse
    power iteration takes its extra squaring there, and above 0.62 (9 % at 1e-3) reports "not
    converged", where the one-sided Jacobi (OpenCV's algorithm) takes over per thread.  Mixed
    warps must give what the C oracle (OpenCV's Jacobi for every point) gives;
    profiles/experiments/dlt_invpower.py counts the fallbacks for this geometry on the CPU."""
    from oracle import c_oracle as co
    from particletracking_b200 import rig as rigmod
    from particletracking_b200 import synthetic as syn

    cal = dict(syn.synthetic_calibration())
    cal["R"] = np.eye(3)
    cal["T"] = np.asarray(cal["T"], dtype=np.float64) * scale
    r = rigmod.rig_from_dicts(cal, syn.synthetic_transformation())
    data = syn.make_stereo(2, 2, 24, seed=77)  # detections of the ordinary rig: geometrically inconsistent here
    c1, c2, n1, n2 = data.problems()
    o = co.match_batch(r, c1, c2, n1, n2, want_cost=True)
    g = run_match(eng, r, c1, c2, n1, n2, want_cost=True)
    np.testing.assert_array_equal(g.status.cpu().numpy(), o["status"])
    cost = g.cost.cpu().numpy()
    # ill-conditioned null vectors: both sides lose digits with 1/(1 - sigma_4/sigma_3)
    assert rel_err(cost, o["cost"]) < 1e-7
    np.testing.assert_array_equal(g.col_ind.cpu().numpy(), o["col_ind"])


@pytest.mark.parametrize("tangential", [False, True])
def test_general_rig_first_camera_not_canonical(eng, rig, tangential):
    """A rig whose FIRST camera has its own extrinsics (P1 = CM1 [R1|t1], R1 != I): none of the canonical-rig
    shortcuts apply, K1 / K3 run the general instantiation (tile kernel, QR of all four DLT columns per
    combination; with tangential terms: the full distortion model as well).  Same detections, the world frame
    moved by a rigid transform: costs, assignments and rows must equal the C oracle's for that rig."""
    from scipy.spatial.transform import Rotation

    from oracle import c_oracle as co
    from particletracking_b200 import rig as rigmod
    from particletracking_b200 import synthetic as syn

    cal = {k: np.array(v, dtype=np.float64) for k, v in rig["calibration"].items()}
    if tangential:
        for key in ("dist1", "dist2"):
            d = np.zeros(14)
            d[: cal[key].size] = cal[key].ravel()
            d[2], d[3] = 3e-4, -2e-4
            cal[key] = d
    trf = rig["transformation"]
    Rg = Rotation.from_rotvec([0.3, -0.2, 0.5]).as_matrix()
    tg = np.array([[40.0], [-25.0], [300.0]])
    R, T = cal["R"], cal["T"].reshape(3, 1)
    r1, t1 = Rg, tg
    r2, t2 = R @ Rg, R @ tg + T
    P1 = cal["CM1"] @ np.hstack([r1, t1])
    P2 = cal["CM2"] @ np.hstack([r2, t2])
    grig = rigmod.make_rig(cal, P1, P2, Rotation.from_matrix(trf["rotation"]), np.asarray(trf["translation"], dtype=np.float64),
                           r1, r2, t1, t2)
    data = syn.make_stereo(3, 2, 22, seed=4242, sigma_px=0.5, p_drop=0.1)
    c1, c2, n1, n2 = data.problems()
    o = co.match_batch(grig, c1, c2, n1, n2, want_cost=True)
    g = run_match(eng, grig, c1, c2, n1, n2, want_cost=True)
    np.testing.assert_array_equal(g.status.cpu().numpy(), o["status"])
    cost = g.cost.cpu().numpy()
    for p in range(len(n1)):
        assert rel_err(cost[p, : n1[p], : n2[p]], o["cost"][p, : n1[p], : n2[p]]) < TOL
    np.testing.assert_array_equal(g.col_ind.cpu().numpy(), o["col_ind"])
    rows = g.rows.cpu().numpy()
    for p in range(len(n1)):
        for k in range(o["n_out"][p]):
            for sl in (slice(0, 3), slice(3, 6), slice(6, 9)):
                assert rel_err(rows[p, k, sl], o["rows"][p, k, sl]) < TOL


def test_off_axis_and_inconsistent_detections(eng, crig):
    """Detections far off the optical axis (|x - cx| > fx), as a wild dummy detection would be: the depth is no
    longer the largest component of the DLT null vector (K1 iterates from the start vector (z, 1), no column is
    chosen), and the x-displaced ones are geometrically inconsistent (sigma_4 / sigma_3 around 0.8): K1's
    convergence test sends them through the out-of-line path (extra squaring, then OpenCV's Jacobi).  Mixed warps
    must give what the C oracle gives: costs to 1e-9, identical assignments, identical rows."""
    from oracle import c_oracle as co
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(4, 2, 20, seed=314, sigma_px=0.5)
    c1, c2, n1, n2 = (np.array(x) for x in data.problems())
    cal = data.calibration
    fx, cx = float(np.asarray(cal["CM1"])[0, 0]), float(np.asarray(cal["CM1"])[0, 2])
    fy, cy = float(np.asarray(cal["CM1"])[1, 1]), float(np.asarray(cal["CM1"])[1, 2])
    c1[:, ::4, 0, 0] = cx + 2.2 * fx     # every fourth camera-1 rod: one end point 65 degrees off axis
    c1[:, 1::4, 1, 1] = cy - 1.6 * fy
    c2[:, 2::5, 0, 0] = float(np.asarray(cal["CM2"])[0, 2]) - 1.9 * float(np.asarray(cal["CM2"])[0, 0])
    o = co.match_batch(crig, c1, c2, n1, n2, want_cost=True)
    g = run_match(eng, crig, c1, c2, n1, n2, want_cost=True)
    np.testing.assert_array_equal(g.status.cpu().numpy(), o["status"])
    cost = g.cost.cpu().numpy()
    for p in range(len(n1)):
        assert rel_err(cost[p, : n1[p], : n2[p]], o["cost"][p, : n1[p], : n2[p]]) < TOL
    np.testing.assert_array_equal(g.col_ind.cpu().numpy(), o["col_ind"])
    rows = g.rows.cpu().numpy()
    for p in range(len(n1)):
        for k in range(o["n_out"][p]):
            for sl in (slice(0, 3), slice(3, 6), slice(6, 9)):
                assert rel_err(rows[p, k, sl], o["rows"][p, k, sl]) < TOL


@pytest.mark.parametrize("C,F,N", [(1, 4, 100), (2, 3, 128), (1, 3, 256), (2, 5, 65)])
def test_tracking_on_the_fly_large_n(eng, C, F, N):
    """N > 64: the tracking LSAP evaluates costs on the fly from the end points (no cost matrix).
    Must equal both the matrix path (cost output requested) and the oracle, incl. flipped rods
    (cross term active) and a NaN frame."""
    from oracle import c_oracle as co

    rng = np.random.default_rng(N + F)
    base = rng.normal(size=(C, 1, N, 2, 3)) * 30
    ends = base + np.cumsum(rng.normal(size=(C, F, N, 2, 3)) * 0.2, axis=1)
    flip = rng.random((C, F, N)) < 0.2
    ends[flip] = ends[flip][:, ::-1]  # end points swapped in some frames -> the cross term wins there
    o = co.track_batch(ends, want_cost=True)
    col_m, tot_m, st_m, cost = eng.track_batch(dev(ends), want_cost=True)  # matrix path
    col_f, tot_f, st_f, _ = eng.track_batch(dev(ends))                    # on-the-fly path
    np.testing.assert_array_equal(cost.cpu().numpy(), o["cost"])
    np.testing.assert_array_equal(col_m.cpu().numpy(), o["col_ind"])
    np.testing.assert_array_equal(col_f.cpu().numpy(), o["col_ind"])
    assert rel_err(tot_f.cpu().numpy(), o["total_cost"]) < 1e-13
    assert (st_f.cpu().numpy() == 0).all()
    bad = ends.copy()
    bad[0, 1, 3, 0, 1] = np.nan
    _, _, st_b, _ = eng.track_batch(dev(bad))
    st_b = st_b.cpu().numpy()
    assert st_b[0, 0] == 1 and (F < 3 or st_b[0, 1] == 1) and (st_b[0, 2:] == 0).all()


@pytest.mark.parametrize("C,F,N,crowd", [(2, 6, 5, 0), (3, 12, 12, 0), (4, 30, 32, 0), (2, 9, 33, 0), (2, 9, 64, 0),
                                          (2, 8, 12, 1), (3, 20, 32, 1), (2, 6, 48, 1), (1, 4, 150, 1), (1, 3, 256, 1)])
def test_tracking_without_matrix_exceptions_and_masked(eng, C, F, N, crowd):
    """k_trk_prepare + k_lsap_trk (no cost matrix): the separable cost is formed in registers, the
    entries where the cross term wins are exceptions.  `crowd = 0`: flipped rods (one exception per
    column, register path); `crowd = 1`: all rods inside a ball the size of one frame's motion, so that
    the cross term wins all over the matrix (several exceptions per column: the masked instantiation).
    Assignments must equal the oracle's SciPy-exact LSAP on the full matrix bit for bit."""
    from oracle import c_oracle as co

    rng = np.random.default_rng(17 * N + F + crowd)
    if crowd:
        base = rng.normal(size=(C, 1, N, 2, 3)) * 0.3
        ends = base + np.cumsum(rng.normal(size=(C, F, N, 2, 3)) * 0.3, axis=1)
    else:
        base = rng.normal(size=(C, 1, N, 2, 3)) * 30
        ends = base + np.cumsum(rng.normal(size=(C, F, N, 2, 3)) * 0.2, axis=1)
        flip = rng.random((C, F, N)) < 0.3
        ends[flip] = ends[flip][:, ::-1]
    o = co.track_batch(ends, want_cost=True)
    d0 = (np.linalg.norm(ends[:, 1:, :, 0] - ends[:, :-1, :, 0], axis=-1)[..., :, None]
          + np.linalg.norm(ends[:, 1:, :, 1] - ends[:, :-1, :, 1], axis=-1)[..., None, :])
    exc = (o["cost"] != d0).sum(axis=2)  # exceptions per column
    assert exc.max() >= (2 if crowd else 1), "the test data does not exercise the intended path"
    col, tot, st, _ = eng.track_batch(dev(ends))
    assert (st.cpu().numpy() == 0).all()
    np.testing.assert_array_equal(col.cpu().numpy(), o["col_ind"])
    assert rel_err(tot.cpu().numpy(), o["total_cost"]) < 1e-13
    # identical rods (exact ties all over: SciPy's tie rule decides) and an infinite coordinate
    tie = ends.copy()
    tie[:, :, 1:] = tie[:, :, :1]
    ot = co.track_batch(tie)
    ct, _, stt, _ = eng.track_batch(dev(tie))
    np.testing.assert_array_equal(ct.cpu().numpy(), ot["col_ind"])
    inf = ends.copy()
    inf[0, 1, 0, 0, 0] = np.inf
    oi = co.track_batch(inf)
    ci, _, sti, _ = eng.track_batch(dev(inf))
    np.testing.assert_array_equal(sti.cpu().numpy(), oi["status"])
    assert (oi["status"] == 2).any()  # a row of +inf costs: SciPy's "cost matrix is infeasible"
    ok = oi["status"] == 0
    np.testing.assert_array_equal(ci.cpu().numpy()[ok], oi["col_ind"][ok])
    assert (ci.cpu().numpy()[~ok] == -1).all()
