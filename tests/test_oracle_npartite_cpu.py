"""CPU tests of the matchND oracle (oracle/npartite.py): the HiGHS restatement of the reference's
0/1 programme against exhaustive enumeration, and the restated match_frame chain against goldens
produced by the reference's own matchND.match_frame (solver patched to the restatement, because
pulp/CBC cannot run here -- tests/golden/make_golden.py:golden_matchnd)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_err
from oracle import npartite as onp
from oracle import port


def _port_args(rig):
    return (rig["calibration"], rig["P1"], rig["P2"], rig["rot"], rig["trans"],
            rig["r1"], rig["r2"], rig["t1"], rig["t2"])


@pytest.mark.parametrize("dims", [(2, 3, 4), (3, 3, 3), (4, 3, 2), (1, 4, 4), (4, 4, 3)])
def test_milp_equals_enumeration(dims):
    rng = np.random.default_rng(sum(dims))
    for kind in ("uniform", "signed", "ties"):
        w = {"uniform": rng.random(dims), "signed": rng.normal(size=dims),
             "ties": rng.integers(0, 3, dims).astype(float)}[kind]
        sol = onp.npartite_matching(w)
        assert onp.is_feasible(dims, sol)
        assert abs(onp.objective(w, sol) - onp.brute_force(w)) < 1e-12
        assert all(np.all(np.diff(sol[0]) > 0) for _ in [0])  # np.where order: sorted by the first axis


def test_reference_test_shapes():
    # reference tests/test_reconstruct_3D/test_match_nd.py:36-47 (all group counts the ILP supports)
    rng = np.random.default_rng(0)
    for shape in [(2, 3, 4), (2, 3), (4, 3, 7), (3, 4, 5, 6), (7, 10, 1, 3, 5)]:
        res = onp.npartite_matching(rng.random(shape))
        assert len(res) == len(shape)
        for i, dim in enumerate(shape):
            assert res[i].shape == (min(shape),) and res[i].max() <= dim - 1


@pytest.mark.parametrize("name", ["matchnd_12.npz", "matchnd_25.npz"])
def test_restated_chain_reproduces_reference_goldens(rig, name):
    g = np.load(os.path.join(GOLDEN, name))
    F = len(g["frames"])
    for f in range(1, F):
        prev = g["rows"][f - 1][:, :6].reshape(-1, 2, 3)
        n1, n2 = g["n1"][f], g["n2"][f]
        rows, costs, lens, w, sol = onp.match_frame_arrays(g["cam1"][f, :n1], g["cam2"][f, :n2], prev, *_port_args(rig))
        assert rel_err(w, g["weights"][f - 1]) < 1e-12
        np.testing.assert_array_equal(np.stack(sol), g["sols"][f - 1])
        assert abs(onp.objective(w, sol) - g["values"][f - 1]) <= 1e-12 * abs(g["values"][f - 1])
        np.testing.assert_array_equal(rows[:, 10:], g["rows"][f][:, 10:])
        assert rel_err(rows[:, :10], g["rows"][f][:, :10]) < 1e-12
        assert rel_err(costs, g["costs"][f]) < 1e-12 and rel_err(lens, g["lens"][f]) < 1e-12


def test_bookkeeping_fill_in_and_unbalanced_error():
    # matchND.py:574-583: previous rods without a triple take the unused detections in index order
    rng = np.random.default_rng(3)
    p = rng.normal(size=(3, 3, 4, 3))
    r1, r2 = rng.normal(size=(3, 2, 2)), rng.normal(size=(3, 2, 2))
    pc = rng.integers(0, 4, (4, 3, 3)).astype(float)
    costs = rng.random((3, 3))
    with pytest.raises(ValueError, match="Unbalanced"):  # 4 previous rods, 3 detections: index 3 overflows
        onp.match_frame_rows(p, r1, r2, costs, pc, (np.array([0, 1, 2]), np.array([2, 0, 1]), np.array([1, 2, 0])), 4)
    out, c = onp.match_frame_rows(p, r1, r2, costs, pc[:3], (np.array([0, 2]), np.array([2, 0]), np.array([1, 2])), 3)
    # rod 1 was not matched: it receives cam1 index 1 and cam2 index 0 (the unused ones)
    np.testing.assert_array_equal(out[1, 10:14], r1[1].ravel())
    np.testing.assert_array_equal(out[1, 14:], r2[0].ravel())
    np.testing.assert_array_equal(c, costs[[2, 1, 0], [1, 0, 2]])


def test_bounding_loop_prototype_reaches_stored_optima():
    """tools/ap3_proto.py is the CPU model of k_ap3's bounding loop (dual descent, then deflected
    subgradient) that its constants were chosen with.  On cubes from frames with dummy / dropped
    detections it must close the bound at the stored HiGHS optimum (tests/golden/ap3_rod_optima.json)."""
    import json
    import os

    from conftest import GOLDEN
    from tools.ap3_check import rod_weights
    from tools.ap3_proto import run

    with open(os.path.join(GOLDEN, "ap3_rod_optima.json")) as f:
        table = json.load(f)
    for name in ("d16_1", "drop16_0", "e24_3"):
        w = rod_weights(**table[name]["args"])
        how, it, st = run(w)
        opt = table[name]["milp"]
        assert how != "gap", (name, how, it)
        assert abs(st["best"] - opt) <= 1e-9 * abs(opt), (name, st["best"] - opt)
        assert st["ub"] >= opt * (1 - 1e-12)
