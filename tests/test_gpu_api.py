"""GPU tests of the reference-shaped Python API (particletracking_b200.reconstruct_3D):
same call, same objects back as the reference -- checked against the DataFrame-level oracle
port (itself pinned to the unmodified reference) and with the structural assertions the
reference's own tests make (ParticleDetection/tests/test_reconstruct_3D/test_match_2d.py)."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def api():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from particletracking_b200.reconstruct_3D import match2D, matchND, tracking

    return match2D, matchND, tracking


def _args(rig):
    return (rig["calibration"], rig["P1"], rig["P2"], rig["rot"], rig["trans"], rig["r1"], rig["r2"], rig["t1"], rig["t2"])


def _cmp_df(got: pd.DataFrame, exp: pd.DataFrame):
    assert list(got.columns) == list(exp.columns)
    assert (got.dtypes.values == exp.dtypes.values).all()
    assert len(got) == len(exp)
    a = got.drop(columns=["color"]).to_numpy(dtype=float)
    b = exp.drop(columns=["color"]).to_numpy(dtype=float)
    np.testing.assert_array_equal(a[:, 10:], b[:, 10:])  # 2D columns, frame, particle, seen: exact
    assert (got["color"].values == exp["color"].values).all()
    for s in (slice(0, 3), slice(3, 6), slice(6, 9), slice(9, 10)):
        d = np.abs(a[:, s] - b[:, s]).max(axis=1)
        assert (d / np.maximum(np.abs(b[:, s]).max(axis=1), 1.0)).max() < TOL


@pytest.mark.parametrize("n,drop,dummy", [(12, 0.0, 0.0), (25, 0.0, 0.0), (16, 0.25, 0.1), (1, 0.0, 0.0)])
def test_match_frame_renumber_true(api, rig, n, drop, dummy):
    from oracle import port
    from particletracking_b200 import synthetic as syn

    m2d = api[0]
    data = syn.make_stereo(3, 1, n, seed=200 + n, p_drop=drop, p_dummy=dummy)
    df = syn.to_dataframes(data)["black"]
    before = df.copy()
    for f in data.frames:
        got = m2d.match_frame(df, "gp1", "gp2", int(f), "black", *_args(rig), renumber=True)
        exp = port.match_frame(df, "gp1", "gp2", int(f), "black", *_args(rig))
        assert len(got) == 3  # reference test: three return values
        _cmp_df(got[0], exp[0])
        assert got[1].shape == exp[1].shape and rel_err(got[1], exp[1]) < TOL
        assert rel_err(got[2], exp[2]) < TOL
        assert len(got[0]) == len(got[1]) == len(got[2])
    pd.testing.assert_frame_equal(df, before)  # inputs are never mutated


def test_match_frame_renumber_false(api, rig):
    from oracle import port
    from particletracking_b200 import synthetic as syn

    m2d = api[0]
    data = syn.make_stereo(2, 1, 25, seed=8)
    inv = np.argsort(data.truth_perm, axis=-1)
    data.cam2 = np.take_along_axis(data.cam2, inv[..., None, None], axis=2)
    df = syn.to_dataframes(data)["black"]
    df["particle"] = df["particle"] + 7
    df.loc[df.index[3], ["x1_gp2", "y1_gp2", "x2_gp2", "y2_gp2"]] = np.nan
    for f in data.frames:
        got = m2d.match_frame(df, "gp1", "gp2", int(f), "black", *_args(rig), renumber=False)
        exp = port.match_frame(df, "gp1", "gp2", int(f), "black", *_args(rig), renumber=False)
        _cmp_df(got[0], exp[0])
        assert rel_err(got[1], exp[1]) < TOL and rel_err(got[2], exp[2]) < TOL
        # reference test_match_frame[False]: each particle's 2D values are a permutation of the input's
        src = df.loc[df.frame == f].set_index("particle")
        c2d = [c for c in df.columns if c.endswith("_gp1") and c[0] in "xy"] + [c for c in df.columns if c.endswith("_gp2") and c[0] in "xy"]
        for _, row in got[0].iterrows():
            assert sorted(row[c2d].to_numpy()) == sorted(src.loc[row["particle"], c2d].to_numpy())


def test_match_frame_empty_and_invalid(api, rig):
    from particletracking_b200 import synthetic as syn

    m2d = api[0]
    data = syn.make_stereo(2, 1, 6, seed=4)
    df = syn.to_dataframes(data)["black"]
    assert m2d.match_frame(df, "gp1", "gp2", 999, "black", *_args(rig)) is None  # no such frame
    df0 = df.copy()
    df0.loc[df0.frame == 0, ["x1_gp1", "y1_gp1", "x2_gp1", "y2_gp1"]] = 0.0      # all-zero rows are dropped
    assert m2d.match_frame(df0, "gp1", "gp2", 0, "black", *_args(rig)) is None
    dfn = df.copy()
    dfn.loc[dfn.index[1], "x1_gp1"] = np.nan  # partially-NaN rod survives the filter -> scipy's ValueError
    with pytest.raises(ValueError, match="invalid numeric entries"):
        m2d.match_frame(dfn, "gp1", "gp2", 0, "black", *_args(rig))
    with pytest.raises(TypeError):  # the reference's drivers unpack None
        m2d.match_complex(df0, [0, 1], "black", rig["calibration"], rig["transformation"])


def test_match_complex_and_threshold(api, rig):
    from oracle import port
    from particletracking_b200 import synthetic as syn

    m2d = api[0]
    data = syn.make_stereo(7, 1, 20, seed=5, p_drop=0.1, sigma_px=0.5)
    df = syn.to_dataframes(data)["black"]
    frames = [int(f) for f in data.frames[1:6]]
    got = m2d.match_complex(df, frames, "black", rig["calibration"], rig["transformation"])
    exp = port.match_complex(df, frames, "black", rig["calibration"], rig["transformation"])
    assert len(got) == 3
    _cmp_df(got[0], exp[0])
    assert isinstance(got[1], list) and len(got[1]) == len(frames)
    for a, b in zip(got[1], exp[1]):
        assert rel_err(a, b) < TOL
    for a, b in zip(got[2], exp[2]):
        assert rel_err(a, b) < TOL
    # threshold extension: a mask over the returned costs, default output unchanged
    thr = m2d.match_complex(df, frames, "black", rig["calibration"], rig["transformation"], max_repr_err=1.5)
    pd.testing.assert_frame_equal(thr[0], got[0])
    for keep, c in zip(thr[3], got[1]):
        np.testing.assert_array_equal(keep, c <= 1.5)


def test_match_csv_complex_roundtrip(api, rig, tmp_path):
    from oracle import port
    from particletracking_b200 import synthetic as syn

    m2d = api[0]
    data = syn.make_stereo(5, 3, 12, seed=6, first_frame=505)
    inp, outp = str(tmp_path / "in"), str(tmp_path / "out")
    syn.write_csvs(data, inp)
    cal_f, trf_f = str(tmp_path / "cal.json"), str(tmp_path / "trf.json")
    json.dump({k: np.asarray(v).tolist() for k, v in rig["calibration"].items()}, open(cal_f, "w"))
    json.dump({k: np.asarray(v).tolist() for k, v in rig["transformation"].items()}, open(trf_f, "w"))
    frames = [505, 506, 507]
    errs, lens = m2d.match_csv_complex(inp, outp, data.colors, "gp1", "gp2", frames, cal_f, trf_f, rematching=True)
    assert errs.shape == (len(data.colors) * len(frames), 12) and lens.shape == errs.shape
    assert sorted(os.listdir(outp)) == sorted(f"rods_df_{c}.csv" for c in data.colors)
    for color in data.colors:
        # pandas' default float parser is not exactly round-tripping; ask for the exact one so
        # that the written values themselves are compared
        res = pd.read_csv(os.path.join(outp, f"rods_df_{color}.csv"), index_col=0, float_precision="round_trip")
        assert list(res["frame"].unique()) == frames  # reference test_match_csv_complex
        src = pd.read_csv(os.path.join(inp, f"rods_df_{color}.csv"), index_col=0)
        exp = port.match_complex(src, frames, color, rig["calibration"], rig["transformation"])[0]
        _cmp_df(res, exp)


def test_tracking_global_assignment_api(api, rig):
    from oracle import port
    from particletracking_b200 import synthetic as syn

    m2d, _, trk = api
    data = syn.make_stereo(9, 1, 14, seed=7)
    df = syn.to_dataframes(data)["black"]
    matched = m2d.match_complex(df, list(data.frames), "black", rig["calibration"], rig["transformation"])[0]
    got, tot = trk.tracking_global_assignment(matched)
    exp, etot = port.tracking_global_assignment(matched)
    pd.testing.assert_frame_equal(got, exp, check_exact=True)   # track ids (positional, App. D.2): bit-exact
    assert rel_err(tot, etot) < 1e-13
    out, tot2, col, tid = trk.tracking_global_assignment_chained(matched)
    ends = matched[["x1", "y1", "z1", "x2", "y2", "z2"]].to_numpy().reshape(9, 14, 2, 3)
    _, ecol, _ = port.tracking_arrays(ends[:, :, 0], ends[:, :, 1])
    np.testing.assert_array_equal(col, ecol)
    np.testing.assert_array_equal(tid, port.chain_tracks(ecol))
    with pytest.raises(ValueError):  # unequal counts per frame: reshape fails in the reference too
        trk.tracking_global_assignment(matched.iloc[:-1])


def test_matchnd_weights_and_block(api, rig):
    from oracle import port
    from particletracking_b200 import synthetic as syn

    _, mnd, _ = api
    data = syn.make_stereo(2, 1, 9, seed=9)
    df = syn.to_dataframes(data)["black"]
    rods1, rods2, costs, p3, re = mnd.triangulate_frame(df, "gp1", "gp2", 1, *_args(rig))
    ref = port.match_arrays(rods1, rods2, *_args(rig), agg="sum")
    assert rel_err(costs, ref["cost"]) < TOL and rel_err(re, ref["rep_errs"]) < TOL
    assert rel_err(p3, ref["p_world"]) < TOL
    prev = port.match_arrays(data.cam1[0, 0], data.cam2[0, 0], *_args(rig))["rows"][:, :6].reshape(-1, 2, 3)
    w, ch = mnd.create_weights(p3, prev, re)
    ew, ech = port.create_weights(ref["p_world"], prev, ref["rep_errs"])
    np.testing.assert_array_equal(ch, ech)
    assert rel_err(w, ew) < 1e-8  # products of two 1e-9-accurate factors
    # a previous frame without 3D columns gives NaN weights: refused, never solved silently
    with pytest.raises(ValueError):
        mnd.match_frame(df, df, "gp1", "gp2", 1, "black", *_args(rig))


def test_reorder_endpoints_vs_reference_golden(api):
    """K7 against the output of the reference's reorder_endpoints_csv (golden) -- swap decisions
    and re-ordered values identical, costs within tolerance -- and through the CSV entry point."""
    import torch
    from conftest import GOLDEN
    from particletracking_b200 import engine

    g = np.load(os.path.join(GOLDEN, "reorder.npz"))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[None])).cuda()
    oe, o2, sw, mc = engine.reorder_endpoints(t(g["ends"]), t(g["pts2d"]))
    np.testing.assert_array_equal(oe.cpu().numpy()[0], g["out_ends"])
    np.testing.assert_array_equal(o2.cpu().numpy()[0], g["out_2d"])
    assert rel_err(mc.cpu().numpy()[0].T, g["mincosts_T"]) < 1e-14
    assert sw.cpu().numpy()[0, 0].sum() == 0


def test_reorder_endpoints_csv_api_and_long_sequence(api, tmp_path):
    import torch
    from oracle import port
    from particletracking_b200 import engine

    m2d = api[0]
    rng = np.random.default_rng(5)
    F, N = 40, 6
    base = rng.normal(size=(1, N, 2, 3)) * 30
    ends = base + np.cumsum(rng.normal(size=(F, N, 2, 3)) * 0.3, axis=0)
    pts = rng.uniform(100, 900, size=(F, N, 2, 2, 2))
    flip = rng.random((F, N)) < 0.5
    ends[flip] = ends[flip][:, ::-1]
    pts[flip] = pts[flip][:, :, ::-1]
    cols3 = ["x1", "y1", "z1", "x2", "y2", "z2"]
    c2d = [f"{k}_{c}" for c in ("gp1", "gp2") for k in ("x1", "y1", "x2", "y2")]
    df = pd.DataFrame(ends.reshape(F * N, 6), columns=cols3)
    df[c2d] = pts.reshape(F * N, 8)
    df["frame"] = np.repeat(np.arange(10, 10 + F), N)
    df["particle"] = np.tile(np.arange(N), F)
    df = df.sample(frac=1.0, random_state=1)  # row order must not matter: rows are sorted by particle per frame
    inp, outp = tmp_path / "in", tmp_path / "out"
    inp.mkdir()
    df.to_csv(inp / "rods_df_red.csv", sep=",")
    mc = m2d.reorder_endpoints_csv(str(inp), str(outp), ["red"], "gp1", "gp2", list(range(10, 10 + F)))
    res = pd.read_csv(outp / "rods_df_red.csv", index_col=0, float_precision="round_trip")
    src = pd.read_csv(inp / "rods_df_red.csv", index_col=0)
    src = src.sort_values(["frame", "particle"])
    oe, o2, sw, omc = port.reorder_endpoints_arrays(src[cols3].to_numpy().reshape(F, N, 2, 3), src[c2d].to_numpy().reshape(F, N, 2, 2, 2))
    np.testing.assert_array_equal(res[cols3].to_numpy().reshape(F, N, 2, 3), oe)
    np.testing.assert_array_equal(res[c2d].to_numpy().reshape(F, N, 2, 2, 2), o2)
    assert mc.shape == (N, F - 1) and rel_err(mc, omc.T) < 1e-14
    assert list(res["frame"]) == list(np.repeat(np.arange(10, 10 + F), N))
    # property at a length the Python loop would crawl through: idempotence (re-ordering the output changes nothing)
    Fl, Nl = 20000, 64
    big = rng.normal(size=(1, 1, Nl, 2, 3)) * 30 + np.cumsum(rng.normal(size=(1, Fl, Nl, 2, 3)) * 0.2, axis=1)
    fl = rng.random((1, Fl, Nl)) < 0.5
    big[fl] = big[fl][:, ::-1]
    o1, _, s1, _ = engine.reorder_endpoints(torch.from_numpy(big).cuda())
    o2_, _, s2, _ = engine.reorder_endpoints(o1)
    assert s1.sum().item() > 0 and s2.sum().item() == 0
    np.testing.assert_array_equal(o1.cpu().numpy(), o2_.cpu().numpy())


def test_project_and_reproject_points_api(api, rig):
    from conftest import GOLDEN
    from particletracking_b200.reconstruct_3D import calibrate_cameras as cc

    g = np.load(os.path.join(GOLDEN, "project.npz"))
    cal, trf = rig["calibration"], rig["transformation"]
    a, b = cc.reproject_points(g["X"], cal, None)
    assert rel_err(a, g["uv1"]) < 1e-12 and rel_err(b, g["uv2"]) < 1e-12
    aw, bw = cc.reproject_points(g["X"], cal, trf)
    assert rel_err(aw, g["uv1_w"]) < 1e-12 and rel_err(bw, g["uv2_w"]) < 1e-12
    p = cc.project_points(g["uv1"].T.copy(), g["uv2"].T.copy(), cal, None)
    assert p.shape == (3, 64) and rel_err(p, g["p3d"]) < TOL
    pw = cc.project_points(g["uv1"].T.copy(), g["uv2"].T.copy(), cal, trf)
    assert rel_err(pw, g["p3d_w"]) < TOL
    # reference test_calibrate_cameras.py: shapes
    assert a.shape == (64, 2) and b.shape == (64, 2)


def test_host_pipeline_compact_result_expands_to_the_full_rows(env_api=None):
    """ptk_reconstruct_host_compact + ptk_expand_rows_host == ptk_reconstruct_host, bit for bit (rows, costs,
    assignments, tracking), on ragged data with a threshold and on the benchmark's shape."""
    import torch
    from particletracking_b200 import engine, rig as rigmod, synthetic as syn

    for (F, C, N, kw) in ((300, 4, 16, dict(p_drop=0.05, p_dummy=0.02)), (260, 8, 32, {})):
        data = syn.make_stereo(F, C, N, seed=41, **kw)
        crig = rigmod.rig_from_dicts(data.calibration, data.transformation)
        track = not kw
        pipe = engine.HostPipeline(0)
        full = pipe.reconstruct(crig, data.cam1, data.cam2, data.n1, data.n2, track=track, max_repr_err=4.0)
        full = {k: (v.clone() if v is not None else None) for k, v in full.items()}
        comp = pipe.reconstruct(crig, data.cam1, data.cam2, data.n1, data.n2, track=track, max_repr_err=4.0, compact=True)
        assert comp["rows"] is None and full["ends"] is None
        rows = engine.expand_rows(comp, data.cam1, data.cam2)
        a, b = rows.numpy(), full["rows"].numpy()
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.array_equal(np.nan_to_num(a).view(np.int64), np.nan_to_num(b).view(np.int64))
        for k in ("match_cost", "row_ind", "col_ind", "n_out", "status", "keep", "track_col", "track_id", "track_total"):
            if full[k] is not None:
                x, y = comp[k].numpy(), full[k].numpy()
                assert np.array_equal(np.nan_to_num(x), np.nan_to_num(y)), k
        pipe.close()


def test_reproject_data_vs_port(api, rig):
    """Plotter.reproject_data (RodTracker backend/reconstruction.py:373-469): our K9-based function against
    the cv2 restatement, on a matched frame (world coordinates) incl. NaN rows and position scaling."""
    import pandas as pd
    from oracle import port
    from particletracking_b200 import synthetic as syn
    from particletracking_b200.reconstruct_3D import calibrate_cameras as cc
    from particletracking_b200.reconstruct_3D import match2D

    data = syn.make_stereo(3, 1, 14, seed=6)
    df = syn.to_dataframes(data)["black"]
    cal, trf = rig["calibration"], rig["transformation"]
    out, _, _ = match2D.match_complex(df, [int(f) for f in data.frames], "black", cal, trf)
    out = out.copy()
    out.loc[out.index[2], ["x1", "y1", "z1", "x2", "y2", "z2"]] = np.nan
    for scale in (1.0, 0.5):
        got = cc.reproject_data(out, ["gp1", "gp2"], cal, scale, trf)
        want = port.reproject_data(out, ["gp1", "gp2"], cal, scale, trf)
        assert got.shape == want.shape and got.shape[0] == 4
        assert rel_err(got, want) < 1e-9
    assert cc.reproject_data(out, ["gp1", "gp2"], None) is None
    empty = out.copy()
    empty[["x1", "y1", "z1"]] = np.nan
    assert cc.reproject_data(empty, ["gp1", "gp2"], cal, 1.0, trf) is None
    ragged = out.copy()  # end 1 missing in one row only: the reference's np.asarray of unequal lists raises, so do we
    ragged.loc[ragged.index[5], ["x1", "y1", "z1"]] = np.nan
    with pytest.raises(ValueError):
        port.reproject_data(ragged, ["gp1", "gp2"], cal, 1.0, trf)
    with pytest.raises(ValueError):
        cc.reproject_data(ragged, ["gp1", "gp2"], cal, 1.0, trf)
