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
