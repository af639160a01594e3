"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference from
``/root/reference`` (present in the build container, absent on the GPU box).

Used by ``tests/golden/make_golden.py`` to generate golden vectors and by the
CPU tests to pin ``oracle/port.py`` / ``oracle/ptk_oracle.c`` against the real
thing whenever the tree is there.  Nothing in ``particletracking_b200/`` may
import this module.

The reference imports two packages that are not installed here and that the
hot path never calls: ``trackpy`` (only
``tracking_trackpy``, tracking.py:36-67) and ``pulp`` (only
``npartite_matching``, matchND.py:45-138).  They are stubbed in
``sys.modules`` so that the modules import; calling the stubbed functions
fails, which is what we want.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_SRC = "/root/reference/ParticleDetection/src"
REFERENCE_EXAMPLES = "/root/reference/ParticleDetection/tests/example_data"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "ParticleDetection"))


def load():
    """Return ``(match2D, matchND, tracking, data_loading)`` reference modules."""
    if not available():
        raise RuntimeError("reference tree not present (expected on the GPU box)")
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    sys.modules.setdefault("trackpy", types.ModuleType("trackpy"))
    if "pulp" not in sys.modules:
        p = types.ModuleType("pulp")
        p.LpSolver = object
        p.PULP_CBC_CMD = lambda **k: None
        sys.modules["pulp"] = p
    import cv2

    cv2.setNumThreads(1)
    import ParticleDetection.reconstruct_3D.match2D as m2d
    import ParticleDetection.reconstruct_3D.matchND as mnd
    import ParticleDetection.reconstruct_3D.tracking as trk
    import ParticleDetection.utils.data_loading as dl

    return m2d, mnd, trk, dl
