"""Host-side CSV writer of libptk.so (no GPU needed): byte-identical to pandas' to_csv for the
rods_df layout, and the float formatter equals Python's repr()."""
import ctypes

import numpy as np
import pandas as pd


def test_format_double_equals_python_repr():
    from particletracking_b200 import _native as nat

    L = nat.lib()
    buf = ctypes.create_string_buffer(40)
    rng = np.random.default_rng(1)
    vals = list(rng.normal(size=5000) * 10.0 ** rng.integers(-25, 25, size=5000))
    vals += [0.0, -0.0, 1.0, -1.0, 100000.0, 1e15, 1e16, 9999999999999998.0, 1e-4, 9.999e-5, 5e-324,
             1.7976931348623157e308, 0.1, 1 / 3, 653.0, 2 ** 53, float("inf"), float("-inf")]
    for v in vals:
        n = L.ptk_format_double(float(v), buf)
        assert buf.raw[:n].decode() == repr(float(v))
    assert L.ptk_format_double(float("nan"), buf) == 0  # pandas writes NaN as an empty field


def test_csv_writer_identical_to_pandas(tmp_path):
    from particletracking_b200.reconstruct_3D import match2D as m2d

    rng = np.random.default_rng(2)
    n = 500
    cols = [*m2d.COLS_3D, *m2d._cam_cols("gp1"), *m2d._cam_cols("gp2")]
    rows = rng.normal(size=(n, 18)) * 10.0 ** rng.integers(-6, 7, size=(n, 18))
    rows[3, 4] = np.nan
    rows[7, :] = np.round(rows[7, :])
    frame = np.repeat(np.arange(500, 500 + n // 20), 20)
    particle = np.tile(np.arange(20), n // 20)
    m = dict(rows=rows, frame=frame, particle=particle, cols=cols, seen_cols=["seen_gp1", "seen_gp2"])
    p1, p2 = str(tmp_path / "ours.csv"), str(tmp_path / "pandas.csv")
    m2d._write_csv(p1, "blue", m)
    df = pd.DataFrame(rows, columns=cols)
    df["frame"] = frame
    df["color"] = "blue"
    df["particle"] = particle
    df[["seen_gp1", "seen_gp2"]] = 1
    df.to_csv(p2, sep=",")
    assert open(p1, "rb").read() == open(p2, "rb").read()
    back = pd.read_csv(p1, index_col=0)
    assert list(back.columns) == list(df.columns) and len(back) == n
