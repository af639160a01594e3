"""Host-side CSV reader of libptk.so (no GPU needed): the DataFrame it yields equals
``pd.read_csv(path, sep=",", index_col=0)`` -- values bit for bit, dtypes, names, index -- and the
float parser equals pandas' default one on every string tried."""
import ctypes
import io
import os

import numpy as np
import pandas as pd
import pytest

from particletracking_b200.utils import csv_io

REF_EXAMPLES = "/root/reference/ParticleDetection/tests/example_data"


def _same(path):
    ours = csv_io.read_rods_csv(str(path))
    ref = pd.read_csv(str(path), sep=",", index_col=0)
    pd.testing.assert_frame_equal(ours, ref, check_exact=True)
    for c in ref.columns:  # assert_frame_equal treats -0.0 == 0.0; compare the bits of the floats
        if ref[c].dtype == np.float64:
            assert ours[c].to_numpy().tobytes() == ref[c].to_numpy().tobytes(), c
    return ours


def test_float_parser_equals_pandas():
    from particletracking_b200 import _native as nat

    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.normal(size=40000) * 10.0 ** rng.integers(-30, 30, size=40000),
                           rng.uniform(0, 2000, 40000), rng.integers(-10 ** 17, 10 ** 17, 1000).astype(float) / 7])
    strs = [repr(float(v)) for v in vals] + ["%.20g" % v for v in vals[:20000]] + ["%.3f" % v for v in vals[40000:60000]]
    strs += ["1e308", "1.7976931348623157e308", "4.9e-324", "2.2250738585072014e-308", "1e-320",
             "123456789012345678901234567890", "0.000000000000000000000000000001234567890123456789", "1E5", "1.e3",
             ".5", "5.", "-0", "+1.5", " 1.5", "1.5 ", "1e+05", "0e0", "00012.50", "9007199254740993", "0.1e-400"]
    ref = pd.read_csv(io.StringIO("a\n" + "\n".join(strs) + "\n"))["a"].to_numpy()
    assert ref.dtype == np.float64
    L = nat.lib()
    out = ctypes.c_double()
    for s, r in zip(strs, ref):
        b = s.encode()
        assert L.ptk_parse_double(b, len(b), ctypes.byref(out)) == 1, s
        assert np.float64(out.value).tobytes() == np.float64(r).tobytes(), (s, out.value, r)
    for s in ["", "abc", "1.5x", "--1", "e5", ".", "1e400"]:
        b = s.encode()
        assert L.ptk_parse_double(b, len(b), ctypes.byref(out)) == 0, s


def test_synthetic_rods_csv(tmp_path):
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(40, 2, 12, seed=5, p_drop=0.1, p_dummy=0.05)
    syn.write_csvs(data, str(tmp_path))
    for color in data.colors:
        df = _same(tmp_path / f"rods_df_{color}.csv")
        assert len(df) == 40 * 12
    # a matched file (3D columns filled) written by pandas and by the library's own writer
    rng = np.random.default_rng(1)
    full = pd.read_csv(tmp_path / f"rods_df_{data.colors[0]}.csv", index_col=0)
    for c in ["x1", "y1", "z1", "x2", "y2", "z2", "x", "y", "z", "l"]:
        full[c] = rng.normal(size=len(full)) * 10.0 ** rng.integers(-8, 8, size=len(full))
    full.to_csv(tmp_path / "full.csv")
    _same(tmp_path / "full.csv")


@pytest.mark.parametrize("threads", [1, 3, 0])
def test_threads_give_identical_tables(tmp_path, threads):
    rng = np.random.default_rng(threads)
    n = 20000
    df = pd.DataFrame({"a": rng.normal(size=n), "b": rng.integers(-5, 5, size=n), "c": rng.choice(["red", "green", "blue"], n),
                       "d": rng.uniform(size=n)})
    df.loc[rng.random(n) < 0.1, "d"] = np.nan
    df.loc[n - 3, "b"] = 7  # ints stay ints
    p = tmp_path / "t.csv"
    df.to_csv(p)
    ours = csv_io.read_rods_csv(str(p), n_threads=threads)
    pd.testing.assert_frame_equal(ours, pd.read_csv(p, index_col=0), check_exact=True)


def test_kind_inference_and_dialect(tmp_path):
    cases = {
        # ints widened to float by a late float / by an NA; NA spellings; strings with NA
        "widen.csv": ",i,f,na,s,b,m\n0,1,1,1,x,True,1\n1,2,2.5,NA,,False,\n2,3,3,,y,True,3\n3,4,nan,null,NaN,False,4\n",
        "crlf.csv": ",a,b\r\n0,1.5,u\r\n1,2.5,v\r\n",
        "blank_lines.csv": ",a,b\n\n0,1,2\n\n\n1,3,4\n",
        "quoted.csv": ',a,b\n0,"hello, world","1.5"\n1,"say ""hi""",2.5\n2,plain,"3"\n',
        "short_rows.csv": ",a,b,c\n0,1,2,3\n1,4\n2,5,6\n",
        "no_trailing_newline.csv": ",a,b\n0,1,2\n1,3,4",
        "index_strings.csv": "name,a\nfoo,1\nbar,2\n",
        "negzero_inf.csv": ",a,b\n0,-0.0,inf\n1,0.0,-inf\n2,1e-5,Infinity\n",
        "header_only.csv": ",a,b\n",
        "bigint.csv": ",a,b\n0,9223372036854775807,1\n1,-9223372036854775808,2\n",
        "mixed_to_string.csv": ",a\n0,1\n1,2.5\n2,abc\n3,\n",
        "all_empty_col.csv": ",a,b\n0,,1\n1,,2\n",
        "spaces.csv": ",a,b\n0, 1, 2.5\n1,3 ,4.5 \n",
    }
    for name, text in cases.items():
        p = tmp_path / name
        p.write_bytes(text.encode())
        _same(p)


def test_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        csv_io.read_rods_csv(str(tmp_path / "missing.csv"))
    p = tmp_path / "toomany.csv"
    p.write_text(",a,b\n0,1,2\n1,3,4,5,6\n")
    with pytest.raises(ValueError, match="tokenizing"):
        csv_io.read_rods_csv(str(p))
    with pytest.raises(Exception):
        pd.read_csv(p, index_col=0)
    (tmp_path / "empty.csv").write_text("")
    with pytest.raises(ValueError):
        csv_io.read_rods_csv(str(tmp_path / "empty.csv"))


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference tree not present")
def test_reference_example_files():
    n = 0
    for root in [REF_EXAMPLES, "/root/reference/RodTracker/src/RodTracker/resources/example_data/csv"]:
        if not os.path.isdir(root):
            continue
        for f in sorted(os.listdir(root)):
            if f.endswith(".csv"):
                _same(os.path.join(root, f))
                n += 1
    assert n >= 2
