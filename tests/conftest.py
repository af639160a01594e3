"""Shared fixtures.  ``-m "not gpu"`` runs the oracle / host-logic tests on CPU;
``-m gpu`` runs the parity tests proper through the C-ABI on a B200."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def rig():
    """Synthetic rig in the form the reference's drivers pass to match_frame."""
    from scipy.spatial.transform import Rotation

    from particletracking_b200 import synthetic as syn

    cal = syn.synthetic_calibration()
    trf = syn.synthetic_transformation()
    P1, P2, r1, r2, t1, t2 = syn.projection_matrices(cal)
    return dict(
        calibration=cal, transformation=trf, P1=P1, P2=P2, r1=r1, r2=r2, t1=t1, t2=t2,
        rot=Rotation.from_matrix(trf["rotation"]), trans=trf["translation"],
    )


def rel_err(a, b):
    """Norm-relative error used for the 1e-9 bar:
    max|a-b| / max(max|b|, 1)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1.0))


def pytest_sessionfinish(session, exitstatus):
    """Bounds-checked debug build (PTK_LIB=.../libptk_dbg.so, ptk_bounds.cuh): the device-side log of index
    violations must be empty after the whole suite, except for what test_bounds_log_* provoked and reset."""
    import ctypes

    try:
        import torch

        if not torch.cuda.is_available():
            return
        from particletracking_b200 import _native as nat

        if nat._lib is None:
            return
        en, cnt, site = ctypes.c_int(0), ctypes.c_longlong(0), ctypes.c_int(0)
        idx, bnd = ctypes.c_longlong(0), ctypes.c_longlong(0)
        nat.lib().ptk_bounds_report(ctypes.byref(en), ctypes.byref(cnt), ctypes.byref(site), ctypes.byref(idx), ctypes.byref(bnd), 0)
        if en.value:
            msg = f"bounds-checked build: {cnt.value} violations (first: site {site.value}, index {idx.value}, bound {bnd.value})"
            print("\n" + msg)
            if cnt.value:
                session.exitstatus = 1
    except Exception as e:  # pragma: no cover
        print(f"\nbounds report unavailable: {e}")
