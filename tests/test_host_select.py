"""Host logic of the batched match path (no GPU): the vectorised row selection / packing equals the
per-frame form that follows match2D.py:824-844 line by line, for both ``renumber`` modes."""
import numpy as np
import pandas as pd
import pytest

from particletracking_b200 import synthetic as syn
from particletracking_b200.reconstruct_3D import match2D as m2d


def _frames_df(seed, p_drop, p_dummy, zeros=False):
    data = syn.make_stereo(30, 1, 9, seed=seed, p_drop=p_drop, p_dummy=p_dummy)
    df = syn.to_dataframes(data)[data.colors[0]]
    if zeros:  # all-zero rows are dropped per camera as well (match2D.py:829-830)
        df.loc[df.index[5], m2d._cam_cols("gp1")] = 0.0
        df.loc[df.index[17], m2d._cam_cols("gp2")] = 0.0
    return df.sample(frac=1.0, random_state=seed)  # shuffled rows: DataFrame order inside a frame matters


@pytest.mark.parametrize("renumber", [True, False])
@pytest.mark.parametrize("p_drop,p_dummy,zeros", [(0.0, 0.0, False), (0.15, 0.05, True)])
def test_select_pack_equals_per_frame_selection(renumber, p_drop, p_dummy, zeros):
    df = _frames_df(3, p_drop, p_dummy, zeros)
    frames = [7, 3, 29, 0, 12]  # any order, a subset
    v1 = df[m2d._cam_cols("gp1")].to_numpy(dtype=np.float64)
    v2 = df[m2d._cam_cols("gp2")].to_numpy(dtype=np.float64)
    c1, c2, n1, n2, kept = m2d._select_pack(df["frame"].to_numpy(), v1, v2, frames, renumber)
    fidx = m2d._FrameIndex(df)
    pos_all = []
    for p, f in enumerate(frames):
        a, b, k = m2d._select(df, fidx, v1, v2, f, renumber)
        assert (n1[p], n2[p]) == (len(a), len(b))
        np.testing.assert_array_equal(c1[p, : len(a)], a)
        np.testing.assert_array_equal(c2[p, : len(b)], b)
        assert np.isnan(c1[p, len(a):]).all() and np.isnan(c2[p, len(b):]).all()
        if not renumber:
            pos_all.append(k)
    if not renumber:
        np.testing.assert_array_equal(kept, np.concatenate(pos_all))
    else:
        assert kept is None


def test_select_pack_special_cases():
    df = _frames_df(4, 0.0, 0.0)
    v1 = df[m2d._cam_cols("gp1")].to_numpy(dtype=np.float64)
    v2 = df[m2d._cam_cols("gp2")].to_numpy(dtype=np.float64)
    fr = df["frame"].to_numpy()
    assert m2d._select_pack(fr, v1, v2, [1, 2, 1], True) is None  # repeated frame -> general path
    with pytest.raises(TypeError):  # a requested frame without rods: the drivers unpack None (match2D.py:622)
        m2d._select_pack(fr, v1, v2, [1, 999], True)


def test_table_equals_dataframe_columns(tmp_path):
    data = syn.make_stereo(5, 1, 6, seed=1)
    syn.write_csvs(data, str(tmp_path))
    path = str(tmp_path / f"rods_df_{data.colors[0]}.csv")
    tab, df = m2d._Table.read(path), pd.read_csv(path, index_col=0)
    assert tab.columns == list(df.columns) and len(tab) == len(df)
    np.testing.assert_array_equal(m2d._index(tab), m2d._index(df))
    for name in ("frame", "particle"):
        np.testing.assert_array_equal(m2d._col(tab, name), m2d._col(df, name))
    cols = m2d._cam_cols("gp1") + m2d._cam_cols("gp2")
    np.testing.assert_array_equal(m2d._cols(tab, cols), m2d._cols(df, cols))
    pd.testing.assert_frame_equal(tab.to_frame(), df, check_exact=True)
