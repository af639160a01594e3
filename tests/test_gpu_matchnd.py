"""GPU tests of the matchND path: the 3-index assignment kernel (K10) against the
HiGHS restatement of the reference's programme (optimal value + feasibility; the optimum itself where
it is unique), and matchND.match_frame / assign against goldens produced by the reference's own code."""
import json
import os
import warnings

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from particletracking_b200 import engine

    return torch, engine


def _solve(gpu, w, **kw):
    torch, engine = gpu
    w = np.ascontiguousarray(w, dtype=np.float64)
    r = engine.npartite3_batch(torch.from_numpy(w[None] if w.ndim == 3 else w).cuda(), **kw)
    return r


def _args(rig):
    return (rig["calibration"], rig["P1"], rig["P2"], rig["rot"], rig["trans"], rig["r1"], rig["r2"], rig["t1"], rig["t2"])


@pytest.mark.parametrize("dims", [(2, 3, 4), (4, 3, 7), (3, 4, 5), (6, 6, 6), (8, 12, 5), (9, 9, 9), (1, 1, 1), (1, 5, 3), (10, 4, 10)])
def test_value_equals_milp_on_random_weights(gpu, dims):
    from oracle import npartite as onp

    rng = np.random.default_rng(sum(dims) * 7)
    for kind in ("uniform", "signed", "ties", "sparse"):
        w = {"uniform": rng.random(dims), "signed": rng.normal(size=dims), "ties": rng.integers(0, 4, dims).astype(float),
             "sparse": rng.random(dims) * (rng.random(dims) < 0.2)}[kind]
        r = _solve(gpu, w)
        n = int(r.n_sol[0])
        sol = r.sol[0, :, :n].cpu().numpy()
        assert int(r.status[0]) == 0
        assert onp.is_feasible(dims, sol) and np.all(np.diff(sol[0]) > 0)
        assert (r.sol[0, :, n:].cpu().numpy() == -1).all()
        ref = onp.objective(w, onp.npartite_matching(w))
        assert abs(float(r.value[0]) - ref) <= 1e-9 * max(1.0, abs(ref)), kind
        assert abs(onp.objective(w, sol) - float(r.value[0])) <= 1e-9 * max(1.0, abs(ref))
        assert float(r.bound[0]) >= float(r.value[0]) - 1e-9 * max(1.0, abs(ref))
        if kind == "uniform":  # all weights positive: min(dims) triples, as the reference's test expects
            assert n == min(dims)


@pytest.mark.parametrize("name", ["matchnd_12.npz", "matchnd_25.npz"])
def test_golden_weights_give_reference_solution(gpu, name):
    # rod data: the optimum is unique, so the triples themselves must be the reference chain's
    g = np.load(os.path.join(GOLDEN, name))
    r = _solve(gpu, g["weights"])  # batch of F-1 problems
    assert (r.status.cpu().numpy() == 0).all()
    for q in range(len(g["weights"])):
        n = int(r.n_sol[q])
        np.testing.assert_array_equal(r.sol[q, :, :n].cpu().numpy(), g["sols"][q])
        assert abs(float(r.value[q]) - g["values"][q]) <= 1e-12 * abs(g["values"][q])
    st = r.stats.cpu().numpy()
    assert (st[:, 3] == 0).all()  # proven by the first descent step, no search


def test_batch_with_ragged_sizes_and_flags(gpu):
    from oracle import npartite as onp

    torch, engine = gpu
    rng = np.random.default_rng(11)
    B, D = 6, (7, 6, 8)
    W = rng.random((B, *D))
    n0 = np.array([7, 3, 0, 5, 7, 1], dtype=np.int32)
    n1 = np.array([6, 6, 4, 2, 6, 1], dtype=np.int32)
    n2 = np.array([8, 5, 8, 8, 1, 1], dtype=np.int32)
    W[4, 2, 3, 0] = np.nan
    r = engine.npartite3_batch(torch.from_numpy(W).cuda(), torch.from_numpy(n0).cuda(), torch.from_numpy(n1).cuda(),
                               torch.from_numpy(n2).cuda())
    st = r.status.cpu().numpy()
    assert st[2] == 3 and st[4] == 2  # empty, invalid
    for b in (0, 1, 3, 5):
        w = W[b, : n0[b], : n1[b], : n2[b]]
        assert st[b] == 0
        ref = onp.objective(w, onp.npartite_matching(w))
        assert abs(float(r.value[b]) - ref) <= 1e-9
        n = int(r.n_sol[b])
        assert n == min(n0[b], n1[b], n2[b]) and onp.is_feasible(w.shape, r.sol[b, :, :n].cpu().numpy())


def _rod_optima():
    with open(os.path.join(GOLDEN, "ap3_rod_optima.json")) as f:
        return json.load(f)


def test_dropped_and_dummy_detections_batch(gpu):
    """Cubes from frames with dropped and dummy (-1,-1,-1,-1) detections: identical rows / columns make
    the dual descent stall, the subgradient phase has to close the bound.  All sizes differ, so they go
    through one ragged batch; optima are HiGHS's (tools/ap3_optima.py)."""
    from oracle import npartite as onp
    from tools.ap3_check import rod_weights

    torch, engine = gpu
    table = _rod_optima()
    names = ["d16_0", "d16_1", "d32_1", "dd25_1", "drop16_2", "e24_0", "e24_4", "e32_1", "e32_2", "e32_4"]
    cubes = [rod_weights(**table[k]["args"]) for k in names]
    D = max(max(c.shape) for c in cubes)
    W = np.zeros((len(cubes), D, D, D))
    n = np.zeros((3, len(cubes)), np.int32)
    for b, c in enumerate(cubes):
        W[b, : c.shape[0], : c.shape[1], : c.shape[2]] = c
        n[:, b] = c.shape
    r = engine.npartite3_batch(torch.from_numpy(W).cuda(), *(torch.from_numpy(x.copy()).cuda() for x in n))
    for b, (k, c) in enumerate(zip(names, cubes)):
        ref = table[k]["milp"]
        ns = int(r.n_sol[b])
        sol = r.sol[b, :, :ns].cpu().numpy()
        assert int(r.status[b]) == 0, (k, r.stats[b].tolist())
        assert onp.is_feasible(c.shape, sol) and ns == min(c.shape)
        assert abs(float(r.value[b]) - ref) <= 1e-9 * abs(ref), (k, float(r.value[b]) - ref)
        assert abs(onp.objective(c, sol) - float(r.value[b])) <= 1e-9 * abs(ref)
        assert float(r.bound[b]) >= ref * (1 - 1e-12)


def test_integrality_gap_is_reported(gpu):
    """e32_5: the LP relaxation of this cube is 306 above the integer optimum, so no dual can prove
    it and the reduced-cost search runs into its node limit: the (here optimal) incumbent comes back
    with status GAP and value <= optimum <= bound."""
    from tools.ap3_check import rod_weights

    rec = _rod_optima()["e32_5"]
    r = _solve(gpu, rod_weights(**rec["args"]), node_limit=100000)
    assert int(r.status[0]) in (0, 1)
    assert float(r.value[0]) <= rec["milp"] * (1 + 1e-12) <= float(r.bound[0]) * (1 + 1e-12)
    assert float(r.value[0]) >= rec["milp"] - 400.0  # within the LP gap


def test_work_limit_is_reported_not_hidden(gpu):
    from oracle import npartite as onp

    rng = np.random.default_rng(5)
    w = rng.random((10, 10, 10))
    ref = onp.objective(w, onp.npartite_matching(w))
    lim = _solve(gpu, w, node_limit=50)
    full = _solve(gpu, w)
    assert int(full.status[0]) == 0 and abs(float(full.value[0]) - ref) <= 1e-9
    if int(lim.status[0]) == 1:  # gap left: value <= optimum <= bound, solution still feasible
        n = int(lim.n_sol[0])
        assert onp.is_feasible(w.shape, lim.sol[0, :, :n].cpu().numpy())
        assert float(lim.value[0]) <= ref + 1e-9 <= float(lim.bound[0]) + 2e-9
    else:
        assert abs(float(lim.value[0]) - ref) <= 1e-9


def test_largest_size(gpu):
    from oracle import npartite as onp

    # 64 per axis (BASELINE config 4's rod count): rod-like weights = strong diagonal under a permutation
    rng = np.random.default_rng(64)
    n = 64
    w = rng.random((n, n, n)) * 0.5
    pj, pk = rng.permutation(n), rng.permutation(n)
    w[np.arange(n), pj, pk] += 5.0
    r = _solve(gpu, w)
    assert int(r.status[0]) == 0 and int(r.n_sol[0]) == n
    sol = r.sol[0].cpu().numpy()
    np.testing.assert_array_equal(sol[1], pj)
    np.testing.assert_array_equal(sol[2], pk)
    assert abs(float(r.value[0]) - w[np.arange(n), pj, pk].sum()) < 1e-9
    torch, engine = gpu
    with pytest.raises(ValueError):
        engine.npartite3_batch(torch.zeros((1, 65, 4, 4), dtype=torch.float64, device="cuda"))


def test_npartite_matching_api(gpu):
    from oracle import npartite as onp
    from particletracking_b200.reconstruct_3D import matchND as mnd

    rng = np.random.default_rng(2)
    for shape in [(2, 3, 4), (2, 3), (4, 3, 7)]:  # reference test_npartite_matching, the group counts built here
        w = rng.random(shape)
        res = mnd.npartite_matching(w)
        assert len(res) == len(shape)
        for i, dim in enumerate(shape):
            assert res[i].shape == (min(shape),) and res[i].max() <= dim - 1
        assert abs(onp.objective(w, res) - onp.objective(w, onp.npartite_matching(w))) < 1e-9
    with pytest.raises(NotImplementedError):
        mnd.npartite_matching(rng.random((3, 4, 5, 6)))
    with pytest.raises(ValueError):
        mnd.npartite_matching(np.full((3, 3, 3), np.nan))
    with pytest.warns(UserWarning, match="Minimization"):  # matchND.py:89-93; the programme's optimum is the empty matching
        res = mnd.npartite_matching(rng.random((3, 3, 3)), maximize=False)
    assert all(len(x) == 0 for x in res)
    w = rng.random((10, 10, 10))
    import particletracking_b200.engine as engine
    orig = engine.npartite3_batch
    engine.npartite3_batch = lambda t, **k: orig(t, node_limit=20)
    try:
        with warnings.catch_warnings(record=True) as rec:
            warnings.simplefilter("always")
            mnd.npartite_matching(w)
        if mnd.npartite_matching.last["status"] == 1:
            assert any(issubclass(x.category, mnd.NpartiteGapWarning) for x in rec)
    finally:
        engine.npartite3_batch = orig


def _golden_df(g):
    from particletracking_b200 import synthetic as syn

    F, N = g["cam1"].shape[:2]
    data = syn.SyntheticStereo(cam1=g["cam1"][:, None], cam2=g["cam2"][:, None], n1=g["n1"][:, None], n2=g["n2"][:, None],
                               frames=g["frames"], colors=["black"], truth_perm=None, world_ends=None, calibration=None,
                               transformation=None)
    return syn.to_dataframes(data)["black"]


@pytest.mark.parametrize("name", ["matchnd_12.npz", "matchnd_25.npz"])
def test_match_frame_chain_vs_reference_golden(gpu, rig, name):
    from particletracking_b200.reconstruct_3D import match2D as m2d
    from particletracking_b200.reconstruct_3D import matchND as mnd

    g = np.load(os.path.join(GOLDEN, name))
    df = _golden_df(g)
    tmp = None
    for fn, frame in enumerate(int(f) for f in g["frames"]):
        if fn == 0:
            tmp, c, l = m2d.match_frame(df, "gp1", "gp2", frame, "black", *_args(rig), renumber=True)
        else:
            tmp, c, l = mnd.match_frame(df, tmp, "gp1", "gp2", frame, "black", *_args(rig))
            assert mnd.npartite_matching.last["status"] == 0
        rows = tmp.iloc[:, :18].to_numpy()
        np.testing.assert_array_equal(rows[:, 10:], g["rows"][fn][:, 10:])  # which detections: exact
        assert rel_err(rows[:, :10], g["rows"][fn][:, :10]) < TOL
        assert rel_err(c, g["costs"][fn]) < TOL and rel_err(l, g["lens"][fn]) < TOL
        assert list(tmp["particle"]) == list(range(len(rows))) and (tmp["frame"] == frame).all()
        assert (tmp[[c_ for c_ in tmp.columns if "seen" in c_]] == 1).all().all()


def test_assign_csv_driver(gpu, rig, tmp_path):
    from oracle import npartite as onp
    from particletracking_b200 import synthetic as syn
    from particletracking_b200.reconstruct_3D import matchND as mnd

    data = syn.make_stereo(4, 2, 10, seed=77)
    inp, outp = str(tmp_path / "in"), str(tmp_path / "out")
    syn.write_csvs(data, inp)
    cal_f, trf_f = str(tmp_path / "cal.json"), str(tmp_path / "trf.json")
    json.dump({k: np.asarray(v).tolist() for k, v in rig["calibration"].items()}, open(cal_f, "w"))
    json.dump({k: np.asarray(v).tolist() for k, v in rig["transformation"].items()}, open(trf_f, "w"))
    frames = [int(f) for f in data.frames]
    errs, lens = mnd.assign(inp, outp, data.colors, "gp1", "gp2", frames, cal_f, trf_f)
    assert errs.shape == (2 * 4, 10) and lens.shape == (2 * 4, 10)
    assert sorted(os.listdir(outp)) == sorted(f"rods_df_{c}.csv" for c in data.colors)
    from oracle import port

    for ci, color in enumerate(data.colors):
        # round_trip: pandas' default float parser is not exact to the last bit, the file is
        res = pd.read_csv(os.path.join(outp, f"rods_df_{color}.csv"), index_col=0, float_precision="round_trip")
        src = pd.read_csv(os.path.join(inp, f"rods_df_{color}.csv"), index_col=0)  # what the reference would read
        assert list(res.frame.unique()) == frames  # reference test_assign
        # the oracle chain on the same detections: frame 0 by match2D, later frames by matchND
        prev = None
        for fn, f in enumerate(frames):
            a, b = port.select_rods(src, "gp1", f), port.select_rods(src, "gp2", f)
            if fn == 0:
                exp = port.match_arrays(a, b, *_args(rig))["rows"]
            else:
                exp = onp.match_frame_arrays(a, b, prev, *_args(rig))[0]
            got = res.loc[res.frame == f].iloc[:, :18].to_numpy()
            np.testing.assert_array_equal(got[:, 10:], exp[:, 10:])
            assert rel_err(got[:, :10], exp[:, :10]) < TOL
            prev = exp[:, :6].reshape(-1, 2, 3)
        # tracked: rod ids follow the same physical rod through the frames (ground truth of the generator)
        ends = res.sort_values(["frame", "particle"])[["x", "y", "z"]].to_numpy().reshape(len(frames), -1, 3)
        assert np.abs(np.diff(ends, axis=0)).max() < 2.0  # a rod moves ~0.2 per frame; a swap would jump by tens


def test_match_frame_unbalanced_and_empty(gpu, rig):
    from particletracking_b200 import synthetic as syn
    from particletracking_b200.reconstruct_3D import match2D as m2d
    from particletracking_b200.reconstruct_3D import matchND as mnd

    data = syn.make_stereo(2, 1, 6, seed=9)
    df = syn.to_dataframes(data)["black"]
    first = m2d.match_frame(df, "gp1", "gp2", 0, "black", *_args(rig), renumber=True)[0]
    fewer = df.drop(df.index[(df.frame == 1)][:2])  # frame 1 holds 4 rods, the previous frame 6
    with pytest.raises(ValueError, match="Unbalanced"):  # matchND.py:592-598
        mnd.match_frame(fewer, first, "gp1", "gp2", 1, "black", *_args(rig))
    assert mnd.match_frame(df, first, "gp1", "gp2", 99, "black", *_args(rig)) is None  # matchND.py:465-467


@pytest.mark.parametrize("name", ["matchnd_12.npz", "matchnd_25.npz"])
def test_device_sequence_vs_reference_golden(gpu, rig, name):
    from particletracking_b200 import rig as rigmod

    torch, engine = gpu
    g = np.load(os.path.join(GOLDEN, name))
    crig = rigmod.make_rig(*_args(rig))
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    rows, costs, n_rows, status, stats = engine.matchnd_sequence(crig, t(g["cam1"][:, None]), t(g["cam2"][:, None]),
                                                                 t(g["n1"][:, None].astype(np.int32)),
                                                                 t(g["n2"][:, None].astype(np.int32)))
    assert (status.cpu().numpy() == 0).all()
    rows, costs = rows.cpu().numpy()[:, 0], costs.cpu().numpy()[:, 0]
    N = g["rows"].shape[1]
    assert (n_rows.cpu().numpy() == N).all()
    np.testing.assert_array_equal(rows[:, :N, 10:], g["rows"][:, :, 10:])
    assert rel_err(rows[:, :N, :10], g["rows"][:, :, :10]) < TOL
    assert rel_err(costs[:, :N], g["costs"]) < TOL
    assert rel_err(rows[:, :N, 9], g["lens"]) < TOL
    assert int(stats.cpu().numpy()[0, 3]) == 0  # no problem needed the search


def test_assign_device_path_equals_per_frame_path(gpu, rig, tmp_path):
    from particletracking_b200 import synthetic as syn
    from particletracking_b200.reconstruct_3D import matchND as mnd

    data = syn.make_stereo(6, 3, 9, seed=91, sigma_px=0.5)
    inp = str(tmp_path / "in")
    syn.write_csvs(data, inp)
    cal_f, trf_f = str(tmp_path / "cal.json"), str(tmp_path / "trf.json")
    json.dump({k: np.asarray(v).tolist() for k, v in rig["calibration"].items()}, open(cal_f, "w"))
    json.dump({k: np.asarray(v).tolist() for k, v in rig["transformation"].items()}, open(trf_f, "w"))
    frames = [int(f) for f in data.frames]
    fast = mnd.assign(inp, str(tmp_path / "fast"), data.colors, "gp1", "gp2", frames, cal_f, trf_f)
    slow = mnd.assign(inp, str(tmp_path / "slow"), data.colors, "gp1", "gp2", frames, cal_f, trf_f,
                      solver=mnd.npartite_matching)  # an explicit solver takes the frame-by-frame Python path
    np.testing.assert_array_equal(fast[0], slow[0])
    np.testing.assert_array_equal(fast[1], slow[1])
    for color in data.colors:
        a = pd.read_csv(tmp_path / "fast" / f"rods_df_{color}.csv", index_col=0)
        b = pd.read_csv(tmp_path / "slow" / f"rods_df_{color}.csv", index_col=0)
        pd.testing.assert_frame_equal(a, b, check_exact=True)


def test_assign_device_path_errors(gpu, rig, tmp_path):
    from particletracking_b200 import synthetic as syn
    from particletracking_b200.reconstruct_3D import matchND as mnd

    data = syn.make_stereo(4, 1, 6, seed=5)
    df = syn.to_dataframes(data)["black"]
    cal_f, trf_f = str(tmp_path / "cal.json"), str(tmp_path / "trf.json")
    json.dump({k: np.asarray(v).tolist() for k, v in rig["calibration"].items()}, open(cal_f, "w"))
    json.dump({k: np.asarray(v).tolist() for k, v in rig["transformation"].items()}, open(trf_f, "w"))
    os.mkdir(tmp_path / "unb")
    df.drop(df.index[df.frame == 2][:2]).to_csv(tmp_path / "unb" / "rods_df_black.csv")
    with pytest.raises(ValueError, match="Unbalanced"):
        mnd.assign(str(tmp_path / "unb"), str(tmp_path / "o1"), ["black"], "gp1", "gp2", [0, 1, 2, 3], cal_f, trf_f)
    assert list(mnd.assign.last["status"][:, 0]) == [0, 0, 6, 8]
    os.mkdir(tmp_path / "emp")
    df.to_csv(tmp_path / "emp" / "rods_df_black.csv")
    with pytest.raises(TypeError):  # frame 7 does not exist: match_frame -> None -> unpacking fails
        mnd.assign(str(tmp_path / "emp"), str(tmp_path / "o2"), ["black"], "gp1", "gp2", [0, 1, 7], cal_f, trf_f)
