"""Boundary behaviour of the C ABI on the GPU: error codes, re-entrancy from several host
threads (the reference's callers run one worker per colour), stream semantics."""
import ctypes
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(rig):
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from particletracking_b200 import _native as nat
    from particletracking_b200 import engine, rig as rigmod

    return engine, nat, rigmod.rig_from_dicts(rig["calibration"], rig["transformation"])


def test_error_codes(env):
    import torch

    engine, nat, crig = env
    L = nat.lib()
    P, N = 4, 8
    f = lambda *s: torch.zeros(s, dtype=torch.float64, device="cuda")
    i = lambda *s: torch.zeros(s, dtype=torch.int32, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    c1, c2, n1, n2 = f(P, N, 2, 2), f(P, N, 2, 2), i(P), i(P)
    row, col, n_out, st, rows, mc = i(P, N), i(P, N), i(P), i(P), f(P, N, 18), f(P, N)
    ws = torch.empty(64, dtype=torch.uint8, device="cuda")
    args = lambda wsb: (ctypes.byref(crig), P, N, p(c1), p(c2), p(n1), p(n2), None, None, p(row), p(col), p(n_out), p(st),
                        p(rows), p(mc), float("inf"), None, p(ws), wsb, None)
    assert L.ptk_match_batch(*args(64)) == -3                       # PTK_ERR_WORKSPACE
    assert b"workspace" in L.ptk_strerror(-3)
    bad = list(args(64))
    bad[2] = 257                                                    # Nmax > PTK_MAX_RODS
    assert L.ptk_match_batch(*bad) == -1
    bad = list(args(64))
    bad[9] = None                                                   # required output missing
    assert L.ptk_match_batch(*bad) == -1
    ctx = ctypes.c_void_p()
    assert L.ptk_ctx_create(99, ctypes.byref(ctx)) == -1
    # all-empty problems are not an error: status EMPTY, nothing written but the padding
    r = engine.match_batch(crig, c1, c2, n1, n2)
    assert (r.status.cpu().numpy() == nat.PROBLEM_EMPTY).all() and (r.n_out.cpu().numpy() == 0).all()
    with pytest.raises(ValueError):
        engine.match_batch(crig, c1.cpu(), c2, n1, n2)             # host tensor to a device entry point
    with pytest.raises(ValueError):
        engine.match_batch(crig, c1.float(), c2, n1, n2)


def test_reentrant_from_threads_and_streams(env):
    """Eight host threads (one per colour, like RodTracker's QThreadPool), each on its own CUDA
    stream, must get the same bits as a serial run."""
    import torch
    from particletracking_b200 import synthetic as syn

    engine, nat, crig = env
    data = syn.make_stereo(60, 8, 20, seed=77, p_drop=0.1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    per_colour = []
    for c in range(8):
        per_colour.append((t(data.cam1[:, c]), t(data.cam2[:, c]), t(data.n1[:, c]), t(data.n2[:, c])))
    serial = [engine.match_batch(crig, *x) for x in per_colour]
    torch.cuda.synchronize()
    out = [None] * 8
    errs = []

    def work(c):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(5):
                    out[c] = engine.match_batch(crig, *per_colour[c])
            s.synchronize()
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(c,)) for c in range(8)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs
    for a, b in zip(serial, out):
        np.testing.assert_array_equal(a.rows.cpu().numpy(), b.rows.cpu().numpy())
        np.testing.assert_array_equal(a.col_ind.cpu().numpy(), b.col_ind.cpu().numpy())


def test_host_pipeline_reports_invalid_like_scipy(env):
    from particletracking_b200 import synthetic as syn

    engine, nat, crig = env
    data = syn.make_stereo(5, 2, 6, seed=3)
    cam1 = data.cam1.copy()
    cam1[2, 1, 0, 0, 0] = np.nan  # partially-NaN rod -> NaN costs -> scipy's ValueError in the reference
    pipe = engine.HostPipeline(0)
    with pytest.raises(ValueError, match="invalid numeric entries"):
        pipe.reconstruct(crig, cam1, data.cam2, data.n1, data.n2)
    out = pipe.reconstruct(crig, cam1, data.cam2, data.n1, data.n2, raise_on_invalid=False)
    st = out["status"].numpy()
    assert st[2, 1] == nat.PROBLEM_INVALID and (np.delete(st.ravel(), 5) == 0).all()
    pipe.close()


def _ragged(F=70, C=3, N=10, seed=11):
    """Frames with dropped detections (n_out < Nmax -> NaN rows -> the tracking LSAP of the pairs
    touching them is INVALID) and one partially-NaN rod: the case ADVICE r1 flagged."""
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(F, C, N, seed=seed, p_drop=0.02)
    cam1 = data.cam1.copy()
    cam1[40, 1, 0, 0, 0] = np.nan
    return data, cam1


def test_chain_tolerates_failed_assignments_device(env):
    """track=True on ragged / NaN input must not fault: failed pairs have col = -1, the chain pass
    ignores them, marks unreached positions -1 and propagates (same rule as the oracle)."""
    import torch
    from oracle import c_oracle as co

    engine, nat, crig = env
    data, cam1 = _ragged()
    F, C, N = cam1.shape[:3]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    c1, c2 = t(cam1.reshape(F * C, N, 2, 2)), t(data.cam2.reshape(F * C, N, 2, 2))
    n1, n2 = t(data.n1.reshape(-1)), t(data.n2.reshape(-1))
    pipe = engine.DevicePipeline(crig, F, C, N, track=True)
    pipe.run(c1, c2, n1, n2)
    torch.cuda.synchronize()
    col = pipe.track_col.cpu().numpy()
    tst = pipe.track_status.cpu().numpy()
    assert (tst != 0).any() and (tst == 0).any()
    assert (col[tst != 0] == -1).all()
    tid = pipe.track_id.cpu().numpy()
    np.testing.assert_array_equal(tid, co.chain_tracks(col))
    assert (tid == -1).any() and tid.min() == -1 and tid.max() < N
    # the generic entry points agree
    m, col2, tot2, tid2, st2 = engine.reconstruct_device(crig, c1, c2, n1, n2, F, C, track=True)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(col2.cpu().numpy(), col)
    np.testing.assert_array_equal(tid2.cpu().numpy(), tid)
    # garbage maps straight into the chain kernels: nothing is dereferenced unchecked
    rng = np.random.default_rng(0)
    junk = rng.integers(-5, N + 5, size=(C, F - 1, N)).astype(np.int32)
    got = engine.chain_tracks(t(junk)).cpu().numpy()
    assert got.min() >= -1 and got.max() < N
    # (duplicates make the result order dependent; compare on a permutation-or-invalid input)
    perm = np.stack([[rng.permutation(N) for _ in range(F - 1)] for _ in range(C)]).astype(np.int32)
    perm[:, 13] = -1
    perm[1, 50, 3] = N + 2
    np.testing.assert_array_equal(engine.chain_tracks(t(perm)).cpu().numpy(), co.chain_tracks(perm))


def test_chain_tolerates_failed_assignments_host_and_sharded(env):
    import torch
    from oracle import c_oracle as co
    from particletracking_b200 import distributed as pdist

    engine, nat, crig = env
    data, cam1 = _ragged(seed=12)
    F, C, N = cam1.shape[:3]
    pipe = engine.HostPipeline(0)
    with pytest.raises(ValueError, match="invalid numeric entries"):
        pipe.reconstruct(crig, cam1, data.cam2, data.n1, data.n2, track=True)
    out = pipe.reconstruct(crig, cam1, data.cam2, data.n1, data.n2, track=True, raise_on_invalid=False)
    col = out["track_col"].numpy().copy()
    tid = out["track_id"].numpy().copy()
    np.testing.assert_array_equal(tid, co.chain_tracks(col))
    assert (tid == -1).any()
    pipe.close()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    rec = pdist.ShardedReconstructor(crig, F, C, N, torch.device("cuda", 0))
    o = rec.step(t(cam1.reshape(F * C, N, 2, 2)), t(data.cam2.reshape(F * C, N, 2, 2)), t(data.n1.reshape(-1)),
                 t(data.n2.reshape(-1)))
    rec.finish()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(o["track_id"].cpu().numpy(), tid)


def test_counts_beyond_nmax_and_tilt_are_rejected(env):
    import torch

    engine, nat, crig = env
    from particletracking_b200 import rig as rigmod, synthetic as syn

    data = syn.make_stereo(3, 2, 6, seed=5)
    c1, c2, n1, n2 = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in data.problems())
    n1 = n1.clone()
    n1[2] = 7  # > Nmax
    r = engine.match_batch(crig, c1, c2, n1, n2)
    st = r.status.cpu().numpy()
    assert st[2] == nat.PROBLEM_INVALID and (np.delete(st, 2) == 0).all()
    cal = dict(data.calibration)
    cal["dist1"] = np.concatenate([np.asarray(cal["dist1"], dtype=float).ravel(), np.zeros(14)])[:14]
    cal["dist1"][12] = 0.01
    with pytest.raises(ValueError, match="tauX"):
        rigmod.rig_from_dicts(cal, data.transformation)
    bad = rigmod.PtkRig.from_buffer_copy(bytes(crig))
    bad.dist2[13] = 1e-3
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    rc = nat.lib().ptk_undistort_batch(ctypes.byref(bad), 6, 6, p(c1), p(c2), p(n1), p(n2), p(ws), p(ws), None)
    assert rc == -1
    pipe = engine.DevicePipeline(crig, 3, 2, 6, track=True)
    with pytest.raises(ValueError, match="float64"):
        pipe.run(c1.float(), c2, n1, n2)
    with pytest.raises(ValueError, match="shape"):
        pipe.run(c1[:4], c2, n1, n2)


def test_workspace_is_per_stream(env):
    """Two streams of one thread must not share scratch: interleaved calls on two
    streams give the serial results."""
    import torch
    from particletracking_b200 import synthetic as syn

    engine, nat, crig = env
    d1 = syn.make_stereo(200, 4, 24, seed=21)
    d2 = syn.make_stereo(200, 4, 24, seed=22)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    a1 = tuple(t(x) for x in d1.problems())
    a2 = tuple(t(x) for x in d2.problems())
    r1 = engine.match_batch(crig, *a1)
    r2 = engine.match_batch(crig, *a2)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(4):
        with torch.cuda.stream(s1):
            q1 = engine.match_batch(crig, *a1)
        with torch.cuda.stream(s2):
            q2 = engine.match_batch(crig, *a2)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(q1.rows.cpu().numpy(), r1.rows.cpu().numpy())
    np.testing.assert_array_equal(q2.rows.cpu().numpy(), r2.rows.cpu().numpy())


def test_bounds_log_counts_a_provoked_violation(env):
    """The debug build's checker itself: an assignment outside the padded problem handed to ptk_rows_batch
    must be counted (and not dereferenced); the release build reports enabled = 0."""
    import torch
    from particletracking_b200 import synthetic as syn

    engine, nat, crig = env
    L = nat.lib()
    en, cnt, site = ctypes.c_int(0), ctypes.c_longlong(0), ctypes.c_int(0)
    idx, bnd = ctypes.c_longlong(0), ctypes.c_longlong(0)
    rep = lambda reset: L.ptk_bounds_report(ctypes.byref(en), ctypes.byref(cnt), ctypes.byref(site), ctypes.byref(idx),
                                            ctypes.byref(bnd), reset)
    assert rep(0) == 0
    if not en.value:
        assert cnt.value == 0
        return
    before = cnt.value
    assert before == 0, f"violations before this test: {before} (site {site.value})"
    data = syn.make_stereo(2, 1, 6, seed=1)
    c1, c2, n1, n2 = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in data.problems())
    col = torch.full((2, 6), 9, dtype=torch.int32, device="cuda")  # 9 >= Nmax = 6
    n_out = torch.full((2,), 6, dtype=torch.int32, device="cuda")
    engine.rows_batch(crig, c1, c2, n1, n2, col, n_out)
    torch.cuda.synchronize()
    assert rep(1) == 0
    assert cnt.value > 0 and site.value == 30 and idx.value == 9 and bnd.value == 6
    assert rep(0) == 0 and cnt.value == 0  # reset: the session-end check sees only real violations


def test_cost_matrix_does_not_depend_on_the_block_shape_of_k1(env, tmp_path):
    """K1 (canonical rigs) gives a CTA `rows_per` camera-1 rods x a chunk of camera-2 rods; the block shape is a
    tuning knob (PTK_KCOST_ROWS, PTK_KCOST_CFG: read once per process) and must not change a bit of the cost
    matrix -- ragged counts, partial row blocks (rows_per = 3, 5), several column chunks (80 rods)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "k1_cost.py"
    script.write_text(
        "import sys, numpy as np, torch\n"
        f"sys.path.insert(0, {root!r})\n"
        "from particletracking_b200 import engine, rig as rigmod, synthetic as syn\n"
        "out = []\n"
        "for n, seed in ((20, 5), (80, 6)):\n"
        "    d = syn.make_stereo(3, 2, n, seed=seed, p_drop=0.15)\n"
        "    r = rigmod.rig_from_dicts(d.calibration, d.transformation)\n"
        "    a = tuple(torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in d.problems())\n"
        "    c = engine.cost_batch(r, *a, want_choice=True)\n"
        "    e = engine.cost_batch(r, *a, want_choice=False)\n"
        "    torch.cuda.synchronize()\n"
        "    out += [c[0].cpu().numpy(), c[1].cpu().numpy(), e[0].cpu().numpy()]\n"
        "np.savez(sys.argv[1], *out)\n")
    res = []
    for k, (rows, cfg) in enumerate((("16", "0"), ("3", "0"), ("5", "6"), ("1", "4"))):
        f = tmp_path / f"c{k}.npz"
        e = dict(os.environ, PTK_KCOST_ROWS=rows, PTK_KCOST_CFG=cfg)
        subprocess.check_call([sys.executable, str(script), str(f)], env=e)
        res.append(np.load(f))
    for other in res[1:]:
        for key in res[0].files:
            a, b = res[0][key], other[key]
            if a.dtype.kind == "f":
                m = ~np.isnan(a)
                assert np.array_equal(np.isnan(a), np.isnan(b))
                assert np.array_equal(a[m].view(np.int64), b[m].view(np.int64))
            else:
                assert np.array_equal(a, b)
