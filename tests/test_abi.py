"""The C-ABI library loads and exports every symbol include/ptk.h declares; the ctypes
struct matches the C struct.  No compute calls (runs without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "ptk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ptk_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from particletracking_b200 import _native as nat

    L = nat.lib()
    declared = _declared()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), f"{name} declared in ptk.h but not exported by libptk.so"
    assert sorted(nat.SYMBOLS) == declared
    assert L.ptk_version() == 100
    assert L.ptk_strerror(0) == b"ok" and b"workspace" in L.ptk_strerror(-3)


def test_rig_struct_layout_matches_header():
    from particletracking_b200.rig import PtkRig

    n_doubles = 9 + 9 + 14 + 14 + 9 + 3 + 9 + 3 + 12 + 12 + 9 + 3
    assert ctypes.sizeof(PtkRig) == n_doubles * 8 + 8
    assert PtkRig.agg_sum.offset == n_doubles * 8


def test_argument_validation_without_gpu():
    """Invalid arguments are rejected before any CUDA call (error behaviour of the boundary)."""
    from particletracking_b200 import _native as nat
    from particletracking_b200.rig import PtkRig

    L = nat.lib()
    rig = PtkRig()
    assert L.ptk_match_batch(ctypes.byref(rig), 4, 300, *([None] * 12), 0.0, None, None, 0, None) == -1
    assert L.ptk_lsap_batch(1, 0, 4, *([None] * 7), None) == -1
    assert L.ptk_workspace_bytes(nat.WS_MATCH, 8000, 32) > 8000 * 32 * 8 * 2
    assert L.ptk_workspace_bytes(99, 1, 1) == 0
    assert L.ptk_launch_count(0) >= 0


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "particletracking_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "ptk_oracle" not in txt, f
