"""CPU tests of the oracle itself (run with -m "not gpu").

The oracle has two layers (see their headers): oracle/port.py (numpy around the same
cv2 / scipy calls the reference makes) and oracle/ptk_oracle.c (plain-C restatement of
the arithmetic inside those binaries).  Both are pinned here against
 (1) the committed goldens made by the unmodified reference (tests/golden/),
 (2) the installed cv2 / scipy / numpy binaries, on seeded random inputs.
"""
import glob
import os

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment
from scipy.spatial.transform import Rotation

from conftest import GOLDEN, rel_err
from oracle import c_oracle as co
from oracle import port
from particletracking_b200 import rig as rigmod
from particletracking_b200 import synthetic as syn

TOL = 1e-9  # north-star tolerance for FP64 coordinates / reprojection errors (norm-relative)


@pytest.fixture(scope="module")
def crig(rig):
    return rigmod.rig_from_dicts(rig["calibration"], rig["transformation"])


def _port_args(rig):
    return (rig["calibration"], rig["P1"], rig["P2"], rig["rot"], rig["trans"],
            rig["r1"], rig["r2"], rig["t1"], rig["t2"])


MATCH_GOLDENS = ["c1_3x12.npz", "ragged_2x16.npz", "tiny_n1.npz", "tiny_n2.npz", "n32.npz"]


# ---------------------------------------------------------------- goldens (reference outputs)
@pytest.mark.parametrize("name", MATCH_GOLDENS)
def test_port_reproduces_reference_goldens(rig, name):
    g = np.load(os.path.join(GOLDEN, name))
    F, C = g["n1"].shape
    for f in range(F):
        for c in range(C):
            n1, n2, n = g["n1"][f, c], g["n2"][f, c], g["n_out"][f, c]
            o = port.match_arrays(g["cam1"][f, c, :n1], g["cam2"][f, c, :n2], *_port_args(rig))
            np.testing.assert_array_equal(o["row_ind"], g["row_ind"][f, c, :n])
            np.testing.assert_array_equal(o["col_ind"], g["col_ind"][f, c, :n])
            np.testing.assert_array_equal(o["rows"][:, 10:], g["rows"][f, c, :n, 10:])
            assert rel_err(o["rows"][:, :10], g["rows"][f, c, :n, :10]) < 1e-12
            assert rel_err(o["cost"], g["cost"][f, c, :n1, :n2]) < 1e-12
            assert rel_err(o["costs"], g["costs"][f, c, :n]) < 1e-12


@pytest.mark.parametrize("name", MATCH_GOLDENS)
def test_c_oracle_reproduces_reference_goldens(crig, name):
    g = np.load(os.path.join(GOLDEN, name))
    F, C, Nm = g["cam1"].shape[:3]
    o = co.match_batch(crig, g["cam1"].reshape(F * C, Nm, 2, 2), g["cam2"].reshape(F * C, Nm, 2, 2),
                       g["n1"].ravel(), g["n2"].ravel(), want_cost=True)
    assert (o["status"] == 0).all()
    np.testing.assert_array_equal(o["n_out"], g["n_out"].ravel())
    np.testing.assert_array_equal(o["row_ind"], g["row_ind"].reshape(F * C, Nm))
    np.testing.assert_array_equal(o["col_ind"], g["col_ind"].reshape(F * C, Nm))
    for p in range(F * C):
        f, c = divmod(p, C)
        n1, n2, n = g["n1"][f, c], g["n2"][f, c], g["n_out"][f, c]
        np.testing.assert_array_equal(o["rows"][p, :n, 10:], g["rows"][f, c, :n, 10:])
        assert rel_err(o["rows"][p, :n, :10], g["rows"][f, c, :n, :10]) < TOL
        assert rel_err(o["cost"][p, :n1, :n2], g["cost"][f, c, :n1, :n2]) < TOL
        assert rel_err(o["match_cost"][p, :n], g["costs"][f, c, :n]) < TOL
        assert rel_err(o["rows"][p, :n, 9], g["lengths"][f, c, :n]) < TOL


@pytest.mark.parametrize("name", ["track_12.npz", "track_25.npz"])
def test_tracking_goldens_bit_exact(name):
    g = np.load(os.path.join(GOLDEN, name))
    ends = g["ends"]  # [F,N,2,3]
    # numpy port
    cost, col, total = port.tracking_arrays(ends[:, :, 0, :], ends[:, :, 1, :])
    np.testing.assert_array_equal(cost, g["cost"])
    np.testing.assert_array_equal(col, g["col_ind"])
    np.testing.assert_array_equal(total, g["total_cost"])
    # C restatement: cost bit-exact, assignment identical, total within tolerance
    o = co.track_batch(ends[None], want_cost=True)
    np.testing.assert_array_equal(o["cost"][0], g["cost"])
    np.testing.assert_array_equal(o["col_ind"][0], g["col_ind"])
    assert rel_err(o["total_cost"][0], g["total_cost"]) < 1e-14
    # positional particle rewrite of the reference
    F, N = ends.shape[:2]
    np.testing.assert_array_equal(g["particle_out"].reshape(F, N)[1:], np.tile(np.arange(N), (F - 1, 1)))


def test_create_weights_golden():
    g = np.load(os.path.join(GOLDEN, "weights.npz"))
    for fn in (port.create_weights, co.create_weights):
        w, ch = fn(g["p_3D"], g["p_prev"], g["rep_errs"])
        np.testing.assert_array_equal(ch, g["choices"])
        assert rel_err(w, g["weights"]) < 1e-13


# ---------------------------------------------------------------- C restatement vs the binaries
def test_c_undistort_bit_exact_vs_cv2(rig):
    import cv2

    rng = np.random.default_rng(0)
    cal = rig["calibration"]
    for cm, d in ((cal["CM1"], cal["dist1"]), (cal["CM2"], cal["dist2"]),
                  (cal["CM1"], np.array([-0.2, 0.5, 1e-3, -2e-3, 0.1, 0.01, 0.02, 0.03]))):
        pts = rng.uniform(-50, 1400, size=(4000, 1, 2))
        ref = cv2.undistortImagePoints(pts, cm, d).reshape(-1, 2)
        rods = pts.reshape(-1, 2, 2)
        # feed proper points through the quirk-free core by using N == 1 problems
        got = np.stack([co.undistort_rods(cm, d, r[None]) for r in rods[:300]])
        exp = ref.reshape(-1, 2, 2)[:300]
        # N == 1 write-back transposes: out = [[x1', x2'], [y1', y2']]
        np.testing.assert_array_equal(got[:, 0], np.swapaxes(exp, 1, 2))
    for N in (1, 2, 3, 7, 32, 64):
        rods = rng.uniform(50, 1200, size=(N, 2, 2))
        for cm, d in ((cal["CM1"], cal["dist1"]), (cal["CM2"], cal["dist2"])):
            np.testing.assert_array_equal(co.undistort_rods(cm, d, rods), port.undistort_quirk(rods, cm, d))


def test_c_triangulate_vs_cv2(rig):
    import cv2

    data = syn.make_stereo(1, 1, 24, seed=9)
    u1 = port.undistort_quirk(data.cam1[0, 0], rig["calibration"]["CM1"], rig["calibration"]["dist1"])
    u2 = port.undistort_quirk(data.cam2[0, 0], rig["calibration"]["CM2"], rig["calibration"]["dist2"])
    a, b = port.pair_combos(u1, u2)
    ph = cv2.triangulatePoints(rig["P1"], rig["P2"], a.T.copy(), b.T.copy()).T
    worst = 0.0
    rot = swp = 0
    for k in range(len(a)):
        X, Xh, st = co.triangulate(rig["P1"], rig["P2"], a[k, 0], a[k, 1], b[k, 0], b[k, 1])
        s = np.sign(Xh @ ph[k])
        worst = max(worst, np.abs(s * Xh - ph[k]).max() / np.abs(ph[k]).max())
        rot += st[0]
        swp += st[1]
    assert worst < 1e-12
    # about 20 rotations / 5 sweeps per point
    assert 15 < rot / len(a) < 26 and 4 < swp / len(a) < 6.5


def test_c_project_vs_cv2(rig):
    import cv2

    rng = np.random.default_rng(4)
    X = rng.uniform([-60, -40, 250], [60, 40, 330], size=(500, 3))
    cal = rig["calibration"]
    for R_, t_, cm, d in ((rig["r1"], rig["t1"], cal["CM1"], cal["dist1"]),
                          (rig["r2"], rig["t2"], cal["CM2"], cal["dist2"])):
        ref = cv2.projectPoints(X, R_, t_, cm, d)[0].reshape(-1, 2)
        got = co.project(R_, t_, cm, d, X)
        assert rel_err(got, ref) < 1e-13
        assert rel_err(syn.project(X, R_, t_, cm, d), ref) < 1e-12


def _lsap_cases():
    rng = np.random.default_rng(7)
    cases = []
    for _ in range(150):
        nr, nc = rng.integers(1, 14, size=2)
        cases.append(rng.random((nr, nc)))
        cases.append(rng.integers(0, 4, size=(nr, nc)).astype(float))  # heavily tied
        m = rng.random((nr, nc))
        m[rng.integers(0, nr)] = m[0]  # identical rows
        m[:, rng.integers(0, nc)] = m[:, 0]  # duplicate columns
        cases.append(m)
    cases.append(np.zeros((9, 9)))  # constant -> identity (scipy gh-11602)
    cases.append(np.ones((5, 8)))
    cases.append(np.ones((8, 5)))
    for n in (32, 64, 100):
        cases.append(rng.random((n, n)))
        cases.append(rng.integers(0, 6, size=(n, n + 7)).astype(float))
        cases.append(rng.integers(0, 6, size=(n + 5, n)).astype(float))
        d1, d2 = rng.random(n), rng.random(n)
        cases.append(d1[:, None] + d2[None, :])  # separable: the tracking cost's tie structure
    return cases


def test_c_lsap_identical_to_scipy():
    for c in _lsap_cases():
        st, ri, ci = co.lsap(c)
        er, ec = linear_sum_assignment(c)
        assert st == 0
        np.testing.assert_array_equal(ri, er)
        np.testing.assert_array_equal(ci, ec)


def test_c_lsap_error_codes():
    c = np.random.default_rng(0).random((4, 5))
    c[1, 2] = np.nan
    assert co.lsap(c)[0] == 1  # scipy: ValueError("matrix contains invalid numeric entries")
    with pytest.raises(ValueError):
        linear_sum_assignment(c)
    c[1, 2] = -np.inf
    assert co.lsap(c)[0] == 1
    c = np.full((3, 3), np.inf)
    assert co.lsap(c)[0] == 2  # scipy: ValueError("cost matrix is infeasible")
    with pytest.raises(ValueError):
        linear_sum_assignment(c)
    c = np.random.default_rng(0).random((3, 3))
    c[0, 1] = np.inf  # finite optimum still exists
    st, ri, ci = co.lsap(c)
    np.testing.assert_array_equal(ci, linear_sum_assignment(c)[1])


@pytest.mark.parametrize("n,drop,dummy,seed", [(12, 0.0, 0.0, 1), (20, 0.25, 0.1, 2), (48, 0.05, 0.02, 3)])
def test_c_match_batch_vs_port(rig, crig, n, drop, dummy, seed):
    data = syn.make_stereo(3, 2, n, seed=seed, p_drop=drop, p_dummy=dummy)
    c1, c2, n1, n2 = data.problems()
    o = co.match_batch(crig, c1, c2, n1, n2, want_cost=True, nthreads=2)
    for p in range(len(n1)):
        r = port.match_arrays(c1[p, : n1[p]], c2[p, : n2[p]], *_port_args(rig))
        k = len(r["row_ind"])
        assert o["n_out"][p] == k
        np.testing.assert_array_equal(o["row_ind"][p, :k], r["row_ind"])
        np.testing.assert_array_equal(o["col_ind"][p, :k], r["col_ind"])
        assert (o["row_ind"][p, k:] == -1).all()
        np.testing.assert_array_equal(o["rows"][p, :k, 10:], r["rows"][:, 10:])
        assert rel_err(o["rows"][p, :k, :10], r["rows"][:, :10]) < TOL
        assert rel_err(o["cost"][p, : n1[p], : n2[p]], r["cost"]) < TOL
        assert rel_err(o["match_cost"][p, :k], r["costs"]) < TOL


def test_c_match_sum_aggregation(rig):
    """matchND's error is the SUM over cameras (matchND.py:525)."""
    r = rigmod.rig_from_dicts(rig["calibration"], rig["transformation"], agg_sum=True)
    data = syn.make_stereo(1, 1, 6, seed=5)
    o = co.cost_problem(r, data.cam1[0, 0], data.cam2[0, 0], want_combos=True)
    ref = port.match_arrays(data.cam1[0, 0], data.cam2[0, 0], *_port_args(rig), agg="sum")
    assert rel_err(o["cost"], ref["cost"]) < TOL
    assert rel_err(o["p_world"], ref["p_world"]) < TOL
    assert rel_err(o["rep_errs"], ref["rep_errs"]) < TOL


def test_c_tracking_and_chain_vs_port():
    rng = np.random.default_rng(11)
    C, F, N = 2, 7, 9
    ends = rng.normal(size=(C, F, N, 2, 3)) * 30
    o = co.track_batch(ends, want_cost=True)
    for c in range(C):
        cost, col, total = port.tracking_arrays(ends[c, :, :, 0], ends[c, :, :, 1])
        np.testing.assert_array_equal(o["cost"][c], cost)
        np.testing.assert_array_equal(o["col_ind"][c], col)
        assert rel_err(o["total_cost"][c], total) < 1e-14
        np.testing.assert_array_equal(co.chain_tracks(o["col_ind"])[c], port.chain_tracks(col))
    ids0 = rng.permutation(N).astype(np.int32)[None].repeat(C, 0)
    t = co.chain_tracks(o["col_ind"], ids0)
    np.testing.assert_array_equal(t[:, 0], ids0)
    # chained ids are a permutation in every frame
    assert (np.sort(t, axis=-1) == np.arange(N)).all()


def test_synthetic_data_is_matchable(rig, crig):
    """App. F sanity: the LSAP recovers the planted correspondence and costs are ~1 px."""
    data = syn.make_stereo(2, 2, 32, seed=2)
    c1, c2, n1, n2 = data.problems()
    o = co.match_batch(crig, c1, c2, n1, n2)
    truth = data.truth_perm.reshape(-1, 32)  # cam2 row r shows rod truth[r]
    ok = 0
    for p in range(len(n1)):
        ok += (truth[p][o["col_ind"][p]] == np.arange(32)).sum()
    assert ok >= 0.98 * truth.size
    assert 0.3 < np.median(o["match_cost"]) < 3.0
    assert abs(np.median(o["rows"][..., 9]) - 10.0) < 0.2


def test_reorder_and_project_goldens():
    g = np.load(os.path.join(GOLDEN, "reorder.npz"))
    oe, o2, sw, mc = port.reorder_endpoints_arrays(g["ends"], g["pts2d"])
    np.testing.assert_array_equal(oe, g["out_ends"])
    np.testing.assert_array_equal(o2, g["out_2d"])
    np.testing.assert_array_equal(mc.T, g["mincosts_T"])
    assert sw[1:].any()
    g = np.load(os.path.join(GOLDEN, "project.npz"))
    cal, trf = syn.synthetic_calibration(), syn.synthetic_transformation()
    a, b = port.reproject_points(g["X"], cal, None)
    np.testing.assert_array_equal(a, g["uv1"])
    np.testing.assert_array_equal(port.project_points(a.T.copy(), b.T.copy(), cal, trf), g["p3d_w"])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_c_oracle_vs_cv2_on_random_rigs(seed):
    """The C restatement against the cv2/scipy port on other stereo geometries than the bench rig."""
    from scipy.spatial.transform import Rotation

    rng = np.random.default_rng(seed)
    trf = syn.synthetic_transformation()
    f1, f2 = rng.uniform(800, 4000, size=2)
    cal = {
        "CM1": np.array([[f1, 0, rng.uniform(300, 900)], [0, f1 * 1.01, rng.uniform(250, 700)], [0, 0, 1.0]]),
        "CM2": np.array([[f2, 0, rng.uniform(300, 900)], [0, f2 * 0.99, rng.uniform(250, 700)], [0, 0, 1.0]]),
        "dist1": np.array([rng.uniform(-0.2, 0.1), rng.uniform(-0.5, 1.0), 1e-3, -1e-3]),
        "dist2": np.array([rng.uniform(-0.2, 0.1), rng.uniform(-0.5, 1.0), 0.0, 0.0, 0.1]),
    }
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    R = Rotation.from_rotvec(axis * np.deg2rad(rng.uniform(15, 95))).as_matrix()
    cal["R"] = R
    cal["T"] = (np.array([0.0, 0.0, 350.0]) - R @ (-trf["rotation"].T @ trf["translation"])).reshape(3, 1)
    r = rigmod.rig_from_dicts(cal, trf)
    P1, P2, r1, r2, t1, t2 = syn.projection_matrices(cal)
    data = syn.make_stereo(1, 1, 14, seed=seed, p_dummy=0.1, calibration=cal)
    ref = port.match_arrays(data.cam1[0, 0], data.cam2[0, 0], cal, P1, P2, Rotation.from_matrix(trf["rotation"]),
                            trf["translation"], r1, r2, t1, t2)
    o = co.cost_problem(r, data.cam1[0, 0], data.cam2[0, 0], want_combos=True)
    assert rel_err(o["cost"], ref["cost"]) < TOL
    np.testing.assert_array_equal(o["choice"], ref["choice"])
    d = np.abs(o["p_world"] - ref["p_world"]).max(axis=-1) / np.maximum(np.abs(ref["p_world"]).max(axis=-1), 1.0)
    assert d.max() < TOL


def test_chain_broken_maps_port_vs_c():
    """Failed assignments (col = -1 / out of range) break the chain: -1 from there on, in both
    restatements of the chain-pass extension."""
    from oracle import c_oracle as co
    from oracle import port

    rng = np.random.default_rng(4)
    F, N = 40, 7
    col = np.stack([rng.permutation(N) for _ in range(F - 1)]).astype(np.int32)
    col[9] = -1
    col[20, 2] = N + 3
    a = port.chain_tracks(col)
    b = co.chain_tracks(col[None])[0]
    np.testing.assert_array_equal(a, b)
    assert (a[:10] >= 0).all() and (a[10:] == -1).all()
    col = np.stack([rng.permutation(N) for _ in range(F - 1)]).astype(np.int32)
    col[20, 2] = N + 3
    a = port.chain_tracks(col)
    np.testing.assert_array_equal(a, co.chain_tracks(col[None])[0])
    assert (a[21:] == -1).sum(axis=1).tolist() == [1] * (F - 21)


def test_expand_rows_host_rebuilds_the_oracle_rows_bit_for_bit():
    """ptk_expand_rows_host (host code of libptk.so, no GPU needed): from the six computed coordinates,
    the straight / crossed verdict and col_ind it must rebuild the reference-shaped 18 columns exactly --
    checked against the C oracle's rows on ragged data (N1 != N2, dummies)."""
    import ctypes

    from oracle import c_oracle as co
    from particletracking_b200 import _native as nat
    from particletracking_b200 import rig as rigmod
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(40, 3, 17, seed=8, p_drop=0.08, p_dummy=0.03)
    crig = rigmod.rig_from_dicts(data.calibration, data.transformation)
    c1, c2, n1, n2 = data.problems()
    o = co.match_batch(crig, c1, c2, n1, n2)
    P, N = c1.shape[:2]
    rows = o["rows"]
    ends = np.ascontiguousarray(rows[..., :6])
    k = np.arange(N)[None, :] < o["n_out"][:, None]
    # straight <=> both cameras' end points are in input order (crossed rows carry both reversed); a dummy
    # rod (-1,-1,-1,-1) reads the same both ways, hence both cameras are consulted
    m = np.broadcast_to(np.arange(N), (P, N))
    pi = np.arange(P)[:, None]
    r1 = c1.reshape(P, N, 4)[pi, m]
    r2 = c2.reshape(P, N, 4)[pi, np.where(k, o["col_ind"], 0)]
    straight = (rows[..., 10:14] == r1).all(axis=-1) & (rows[..., 14:18] == r2).all(axis=-1)
    rev = (rows[..., 10:14] == r1[..., [2, 3, 0, 1]]).all(axis=-1) & (rows[..., 14:18] == r2[..., [2, 3, 0, 1]]).all(axis=-1)
    assert (straight | rev)[k].all()
    straight = np.ascontiguousarray((straight & k).astype(np.uint8))
    col = np.ascontiguousarray(np.where(k, o["col_ind"], -1).astype(np.int32))
    out = np.full((P, N, 18), 7.0)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    c1c, c2c = np.ascontiguousarray(c1), np.ascontiguousarray(c2)
    n_out = np.ascontiguousarray(o["n_out"].astype(np.int32))
    for nt in (1, 3):
        rc = nat.lib().ptk_expand_rows_host(P, N, p(c1c), p(c2c), p(ends), p(straight), p(col), p(n_out), p(out), nt)
        assert rc == 0
        assert np.array_equal(out[k].view(np.int64), rows[k].view(np.int64))  # bit for bit, incl. centre and length
        assert np.isnan(out[~k]).all()
    bad = col.copy()
    bad[0, 0] = N + 1
    assert nat.lib().ptk_expand_rows_host(P, N, p(c1c), p(c2c), p(ends), p(straight), p(bad), p(n_out), p(out), 1) == -1
