"""``match2D`` of the reference (``reconstruct_3D/match2D.py``) on the B200.

Same signatures, return types and error behaviour as the reference's ``match_frame``
(:741-998), ``match_complex`` (:634-738) and ``match_csv_complex`` (:513-631).  The numeric
body (undistort -> pairs -> DLT -> reproject -> cost -> LSAP -> rows) runs in ``libptk.so``;
what stays here is the DataFrame/CSV glue, batched over frames (and colours) so that one
call to the library covers the whole request.

Keyword-only extras (default off): ``max_repr_err`` adds a ``keep`` mask over
the returned costs (north-star "threshold it"); nothing else changes.
"""
from __future__ import annotations

import logging
import os
import pathlib
import threading
import time
from typing import Iterable, List, Optional, Sequence

import numpy as np
import pandas as pd
import torch

from .. import engine
from ..rig import make_rig
from ..utils import csv_io
from ..utils import data_loading as dl

_logger = logging.getLogger(__name__)

_CSV_WORKERS = 4  # colours of match_csv_complex in flight
COLS_3D = ["x1", "y1", "z1", "x2", "y2", "z2", "x", "y", "z", "l"]


def _cam_cols(name: str) -> List[str]:
    return [f"x1_{name}", f"y1_{name}", f"x2_{name}", f"y2_{name}"]


class _FrameIndex:
    """Row positions of every frame, in DataFrame order (``data.loc[data.frame == f]``)."""

    def __init__(self, data: pd.DataFrame):
        fr = data["frame"].to_numpy()
        self.order = np.argsort(fr, kind="stable")
        self.sorted = fr[self.order]

    def rows(self, frame) -> np.ndarray:
        lo = np.searchsorted(self.sorted, frame, side="left")
        hi = np.searchsorted(self.sorted, frame, side="right")
        return self.order[lo:hi]


def _valid(v: np.ndarray) -> np.ndarray:
    """match2D.py:826-830: drop rows that are all NaN, then rows that are all zero
    (partially-NaN rows and the detector's -1 dummies survive)."""
    return ~np.isnan(v).all(axis=1) & (v != 0).any(axis=1)


def _pack(problems: Sequence[tuple]):
    """[(rods1[n1,4], rods2[n2,4]), ...] -> padded host tensors for the library."""
    P = len(problems)
    nmax = max(1, max(max(len(a), len(b)) for a, b in problems))
    if nmax > 256:
        raise ValueError("more than 256 rods in one (frame, colour) problem")
    c1 = np.full((P, nmax, 4), np.nan)
    c2 = np.full((P, nmax, 4), np.nan)
    n1 = np.zeros(P, dtype=np.int32)
    n2 = np.zeros(P, dtype=np.int32)
    for p, (a, b) in enumerate(problems):
        n1[p], n2[p] = len(a), len(b)
        c1[p, : len(a)] = a
        c2[p, : len(b)] = b
    dev = torch.device("cuda", torch.cuda.current_device())
    t = lambda x: torch.from_numpy(x).to(dev)
    return t(c1.reshape(P, nmax, 2, 2)), t(c2.reshape(P, nmax, 2, 2)), t(n1), t(n2)


class _Table:
    """Column store the batched path works on: what the native CSV reader returns, without the
    detour through a DataFrame (``index`` = the file's first column)."""

    def __init__(self, names: Sequence[str], cols: dict):
        self.index = cols[0]
        self.columns = list(names[1:])
        self._pos = {n: c for c, n in enumerate(names) if c > 0}
        self._cols = cols

    @classmethod
    def read(cls, path: str) -> "_Table":
        return cls(*csv_io.read_columns(path))

    def __len__(self):
        return len(self.index)

    def to_frame(self) -> pd.DataFrame:
        return pd.DataFrame({n: self._cols[c] for n, c in self._pos.items()}, index=pd.Index(self.index))

    def col(self, name: str) -> np.ndarray:
        try:
            return self._cols[self._pos[name]]
        except KeyError:
            raise KeyError(name) from None


def _col(data, name: str) -> np.ndarray:
    return data.col(name) if isinstance(data, _Table) else data[name].to_numpy()


def _cols(data, names: Sequence[str]) -> np.ndarray:
    if isinstance(data, _Table):
        return np.stack([np.asarray(data.col(n), dtype=np.float64) for n in names], axis=1)
    return data[list(names)].to_numpy(dtype=np.float64)


def _index(data) -> np.ndarray:
    return data.index if isinstance(data, _Table) else data.index.to_numpy()


def _select_pack(fr: np.ndarray, v1: np.ndarray, v2: np.ndarray, frames: Sequence[int], renumber: bool):
    """match2D.py:824-844 for all requested frames at once (no Python loop over frames): returns
    padded ``cam1, cam2 [P,nmax,4]``, ``n1, n2 [P]`` and, for ``renumber=False``, the row positions
    kept per problem (concatenated in problem order).  None if this form does not apply
    (frame numbers listed twice)."""
    frames_a = np.asarray(frames)
    P = len(frames_a)
    order_f = np.argsort(frames_a, kind="stable")
    fs = frames_a[order_f]
    if P > 1 and (fs[1:] == fs[:-1]).any():
        return None
    pos = np.minimum(np.searchsorted(fs, fr), P - 1)
    pidx = np.where(fs[pos] == fr, order_f[pos], -1)  # problem of every row, -1 = frame not requested
    k1, k2 = _valid(v1), _valid(v2)
    if not renumber:
        k1 = k2 = k1 & k2

    def pack(v, k):
        sel = np.flatnonzero(k & (pidx >= 0))
        o = np.argsort(pidx[sel], kind="stable")  # DataFrame order inside every frame
        sel = sel[o]
        ps = pidx[sel]
        n = np.bincount(ps, minlength=P)
        rank = np.arange(len(sel)) - np.repeat(np.cumsum(n) - n, n)
        return sel, ps, rank, n

    s1, p1, r1, n1 = pack(v1, k1)
    s2, p2, r2, n2 = pack(v2, k2)
    if P and ((n1 == 0).any() or (n2 == 0).any()):
        # the reference's drivers unpack match_frame's None here (match2D.py:622, 733)
        raise TypeError("cannot unpack non-iterable NoneType object")
    nmax = int(max(1, n1.max(initial=0), n2.max(initial=0)))
    if nmax > 256:
        raise ValueError("more than 256 rods in one (frame, colour) problem")
    c1 = np.full((P, nmax, 4), np.nan)
    c2 = np.full((P, nmax, 4), np.nan)
    c1[p1, r1] = v1[s1]
    c2[p2, r2] = v2[s2]
    return c1, c2, n1.astype(np.int32), n2.astype(np.int32), (s1 if not renumber else None)


_pipelines = threading.local()


def _host_pipeline() -> "engine.HostPipeline":
    """One ``ptk_ctx`` (streams, pinned staging, device buffers) per calling thread: the reference
    is driven from one thread per colour, so nothing is shared between calls."""
    dev = torch.cuda.current_device()
    hp = getattr(_pipelines, "hp", None)
    if hp is None or hp.device != dev:
        hp = engine.HostPipeline(dev)
        _pipelines.hp = hp
    return hp


def _raise_for_status(status: np.ndarray):
    # scipy.optimize.linear_sum_assignment raises these from match2D.py:933
    if (status == 1).any():
        raise ValueError("matrix contains invalid numeric entries")
    if (status == 2).any():
        raise ValueError("cost matrix is infeasible")


def _solve(problems, rig, renumber: bool, max_repr_err: Optional[float]):
    """Run the library on a list of (rods_cam1, rods_cam2) problems.  Returns per problem
    (rows[n,18], costs[n], keep[n] | None)."""
    engine._require_cuda()
    c1, c2, n1, n2 = _pack(problems)
    if renumber:
        r = engine.match_batch(rig, c1, c2, n1, n2, max_repr_err=max_repr_err)
        status = r.status.cpu().numpy()
        _raise_for_status(status)
        rows, costs, n_out = r.rows.cpu().numpy(), r.match_cost.cpu().numpy(), r.n_out.cpu().numpy()
        keep = r.keep.cpu().numpy().astype(bool) if r.keep is not None else None
    else:
        # match2D.py:876-884, 941-955: pairs are kept (zip), only end-point combos are evaluated
        P, nmax = c1.shape[:2]
        col = torch.arange(nmax, dtype=torch.int32, device=c1.device).expand(P, nmax).contiguous()
        rows_t, pc = engine.rows_batch(rig, c1, c2, n1, n2, col, n1)
        rows, costs, n_out = rows_t.cpu().numpy(), pc.cpu().numpy(), n1.cpu().numpy()
        keep = (costs <= max_repr_err) if max_repr_err is not None else None
    out = []
    for p in range(len(problems)):
        n = int(n_out[p])
        out.append((rows[p, :n], costs[p, :n], None if keep is None else keep[p, :n]))
    return out


def _select(data: pd.DataFrame, fidx: _FrameIndex, v1: np.ndarray, v2: np.ndarray, frame, renumber: bool):
    """match2D.py:824-844 for one frame -> (rods1[n1,4], rods2[n2,4], positions kept for
    renumber=False)."""
    pos = fidx.rows(frame)
    a, b = v1[pos], v2[pos]
    k1, k2 = _valid(a), _valid(b)
    if not renumber:
        # match2D.py:831-837: keep index labels present in both cameras.  Labels, not positions:
        # with a non-unique index a label survives if ANY of its rows is valid in both.
        both = k1 & k2
        lab = data.index.to_numpy()[pos]
        if len(np.unique(lab)) != len(lab):
            drop = set(lab[k1 ^ k2])
            both = np.array([(l not in drop) for l in lab]) & k1 & k2
        return a[both], b[both], pos[both]
    return a[k1], b[k2], None


def _frame_df(rows, frame, color, particle, cols, seen_cols):
    df = pd.DataFrame(rows, columns=cols)
    df["frame"] = frame
    df["color"] = color
    df["particle"] = particle
    if seen_cols:
        df[seen_cols] = 1
    return df


def match_frame(data: pd.DataFrame, cam1_name: str, cam2_name: str, frame: int, color: str, calibration: dict,
                P1: np.ndarray, P2: np.ndarray, rot, trans: np.ndarray, r1: np.ndarray, r2: np.ndarray,
                t1: np.ndarray, t2: np.ndarray, renumber: bool = True, *, max_repr_err: Optional[float] = None):
    """Matches and triangulates the rods of one frame (reference match2D.py:741-998).

    Returns ``(DataFrame, costs, lengths)`` -- or ``None`` when a camera has no rods in this
    frame (:838-840).  With ``max_repr_err`` a 4th element ``keep`` (bool per row) is appended."""
    cols = [*COLS_3D, *_cam_cols(cam1_name), *_cam_cols(cam2_name)]
    fidx = _FrameIndex(data)
    v1 = data[_cam_cols(cam1_name)].to_numpy(dtype=np.float64)
    v2 = data[_cam_cols(cam2_name)].to_numpy(dtype=np.float64)
    rods1, rods2, kept = _select(data, fidx, v1, v2, frame, renumber)
    if len(rods1) == 0 or len(rods2) == 0:
        return None
    rig = make_rig(calibration, P1, P2, rot, trans, r1, r2, t1, t2)
    rows, costs, keep = _solve([(rods1, rods2)], rig, renumber, max_repr_err)[0]
    if (not renumber) and ("particle" in data.columns):
        particle = data["particle"].to_numpy()[kept]  # match2D.py:990-993
    else:
        particle = list(range(len(rows)))
    seen_cols = [c for c in data.columns if "seen" in c]
    df = _frame_df(rows, frame, color, particle, cols, seen_cols)
    if max_repr_err is not None:
        return df, costs, rows[:, 9].copy(), keep
    return df, costs, rows[:, 9].copy()


def _derive_rig(calibration: dict, transform: dict):
    """match2D.py:584-596 / 689-710: P1, P2, r, t from the calibration; world rotation and
    translation from either transform format."""
    from ..rig import rig_from_dicts

    return rig_from_dicts(calibration, transform)


def _match_arrays_loop(data: pd.DataFrame, frames, rig, cam1_name, cam2_name, renumber: bool, max_repr_err=None):
    """Frame-by-frame selection (the general form: repeated frame numbers, non-unique index)."""
    fidx = _FrameIndex(data)
    v1 = data[_cam_cols(cam1_name)].to_numpy(dtype=np.float64)
    v2 = data[_cam_cols(cam2_name)].to_numpy(dtype=np.float64)
    problems, kepts = [], []
    for f in frames:
        a, b, kept = _select(data, fidx, v1, v2, f, renumber)
        if len(a) == 0 or len(b) == 0:
            # the reference's drivers unpack match_frame's None here (match2D.py:622, 733)
            raise TypeError("cannot unpack non-iterable NoneType object")
        problems.append((a, b))
        kepts.append(kept)
    res = _solve(problems, rig, renumber, max_repr_err)
    n_rows = [len(r[0]) for r in res]
    if (not renumber) and ("particle" in data.columns):
        pcol = data["particle"].to_numpy()
        particle = np.concatenate([pcol[k] for k in kepts])
    else:
        particle = np.concatenate([np.arange(n) for n in n_rows]) if n_rows else np.zeros(0, dtype=np.int64)
    return dict(
        rows=np.concatenate([r[0] for r in res], axis=0),
        frame=np.repeat(np.asarray(frames), n_rows),
        particle=np.asarray(particle),
        costs=[r[1] for r in res],
        lens=[r[0][:, 9].copy() for r in res],
        keeps=[r[2] for r in res],
        costs2d=None, lens2d=None,
    )


def _match_arrays(data, frame_numbers: Iterable[int], rig, cam1_name, cam2_name, renumber: bool,
                  max_repr_err=None, copy_rows: bool = True):
    """All requested frames in one library call.  ``data`` is a DataFrame or a ``_Table``.
    Returns a dict of flat arrays (rows of all frames concatenated) plus the per-frame lists the
    reference returns; with ``copy_rows=False`` ``rows`` may alias the pipeline's pinned output
    buffer, valid until this thread's next call."""
    frames = list(frame_numbers)
    if not frames:
        return None
    common = dict(cols=[*COLS_3D, *_cam_cols(cam1_name), *_cam_cols(cam2_name)],
                  seen_cols=[c for c in data.columns if "seen" in c])
    packed = None
    if renumber or len(np.unique(_index(data))) == len(data):
        v1, v2 = _cols(data, _cam_cols(cam1_name)), _cols(data, _cam_cols(cam2_name))
        packed = _select_pack(np.asarray(_col(data, "frame")), v1, v2, frames, renumber)
    if packed is None:
        if isinstance(data, _Table):
            data = data.to_frame()
        return {**_match_arrays_loop(data, frames, rig, cam1_name, cam2_name, renumber, max_repr_err), **common}
    engine._require_cuda()
    c1, c2, n1, n2, kept = packed
    P, nmax = c1.shape[:2]
    if renumber:
        r = _host_pipeline().reconstruct(rig, c1.reshape(P, 1, nmax, 2, 2), c2.reshape(P, 1, nmax, 2, 2), n1.reshape(P, 1),
                                         n2.reshape(P, 1), track=False, max_repr_err=max_repr_err, raise_on_invalid=False)
        _raise_for_status(r["status"].numpy().reshape(P))
        rows, costs = r["rows"].numpy().reshape(P, nmax, -1), r["match_cost"].numpy().reshape(P, nmax)
        n_out = r["n_out"].numpy().reshape(P)
        keep = r["keep"].numpy().reshape(P, nmax).astype(bool) if r["keep"] is not None else None
    else:
        # match2D.py:876-884, 941-955: pairs are kept (zip), only end-point combos are evaluated
        dev = torch.device("cuda", torch.cuda.current_device())
        t = lambda x: torch.from_numpy(x).to(dev)
        d1, d2, dn = t(c1.reshape(P, nmax, 2, 2)), t(c2.reshape(P, nmax, 2, 2)), t(n1)
        col = torch.arange(nmax, dtype=torch.int32, device=dev).expand(P, nmax).contiguous()
        rows_t, pc = engine.rows_batch(rig, d1, d2, dn, dn, col, dn)
        rows, costs, n_out = rows_t.cpu().numpy(), pc.cpu().numpy(), n1
        keep = (costs <= max_repr_err) if max_repr_err is not None else None
        copy_rows = False
    n0 = int(n_out[0])
    if (n_out == n0).all():  # the usual case: every frame holds the same number of rods
        flat = rows[:, :n0].reshape(P * n0, -1)
        costs2d, lens2d = costs[:, :n0].copy(), rows[:, :n0, 9].copy()
        out = dict(rows=flat.copy() if copy_rows and np.shares_memory(flat, rows) else flat,
                   costs=list(costs2d), lens=list(lens2d), costs2d=costs2d, lens2d=lens2d,
                   keeps=[None] * P if keep is None else list(keep[:, :n0]),
                   particle=np.tile(np.arange(n0), P))
    else:
        mask = np.arange(nmax)[None, :] < n_out[:, None]
        cuts = np.cumsum(n_out)[:-1]
        flat = rows[mask]
        out = dict(rows=flat, costs=np.split(costs[mask], cuts), lens=np.split(flat[:, 9].copy(), cuts), costs2d=None,
                   lens2d=None, keeps=[None] * P if keep is None else np.split(keep[mask], cuts),
                   particle=np.broadcast_to(np.arange(nmax), (P, nmax))[mask])
    if (not renumber) and ("particle" in data.columns):
        out["particle"] = np.asarray(_col(data, "particle"))[kept]  # match2D.py:990-993
    out["frame"] = np.repeat(np.asarray(frames), n_out)
    return {**out, **common}


def _match_many(data: pd.DataFrame, frame_numbers: Iterable[int], color: str, rig, cam1_name, cam2_name,
                renumber: bool, max_repr_err=None):
    m = _match_arrays(data, frame_numbers, rig, cam1_name, cam2_name, renumber, max_repr_err)
    if m is None:
        return pd.DataFrame(), [], [], []
    df = pd.DataFrame(m["rows"], columns=m["cols"])
    df["frame"] = m["frame"]
    df["color"] = color
    df["particle"] = m["particle"]
    if m["seen_cols"]:
        df[m["seen_cols"]] = 1
    return df, m["costs"], m["lens"], m["keeps"]


def _write_csv(path: str, color: str, m: dict):
    """``df_out.to_csv(path, sep=",")`` of the frame layout (match2D.py:626-629, 986-997) through
    the library's writer: same bytes as pandas, ~25x faster."""
    import ctypes

    from .. import _native as nat

    header = "," + ",".join([*m["cols"], "frame", "color", "particle", *m["seen_cols"]])
    rows = np.ascontiguousarray(m["rows"], dtype=np.float64)
    frame = np.ascontiguousarray(m["frame"], dtype=np.int64)
    particle = np.ascontiguousarray(m["particle"], dtype=np.int64)
    rc = nat.lib().ptk_csv_write_rods(path.encode(), header.encode(), len(rows), rows.shape[1] if rows.ndim == 2 else 0,
                                      rows.ctypes.data_as(ctypes.c_void_p), frame.ctypes.data_as(ctypes.c_void_p),
                                      str(color).encode(), particle.ctypes.data_as(ctypes.c_void_p), len(m["seen_cols"]))
    if rc != 0:
        raise OSError(f"could not write {path}")


def match_complex(data: pd.DataFrame, frame_numbers: Iterable[int], color: str, calibration: dict, transform: dict,
                  cam1_name="gp1", cam2_name="gp2", renumber: bool = True, *, max_repr_err: Optional[float] = None):
    """Matches and triangulates rods from a DataFrame (reference match2D.py:634-738): all
    requested frames go to the GPU in one batch.  Returns ``(DataFrame, costs, lengths)`` with
    per-frame lists, as the reference does."""
    rig = _derive_rig(calibration, transform)
    df, costs, lens, keeps = _match_many(data, frame_numbers, color, rig, cam1_name, cam2_name, renumber, max_repr_err)
    if max_repr_err is not None:
        return df, costs, lens, keeps
    return df, costs, lens


def match_csv_complex(input_folder, output_folder, colors, cam1_name="gp1", cam2_name="gp2", frame_numbers=None,
                      calibration_file=None, transformation_file=None, rematching=True):
    """Matches and triangulates rods from ``rods_df_{color}.csv`` files and writes files of
    the same layout (reference match2D.py:513-631).  Returns ``(reprojection errors, rod
    lengths)`` as ndarrays (``np.array`` of the per-frame lists, like the reference)."""
    if calibration_file is None:
        raise FileNotFoundError(
            "calibration_file: the reference's default (example_calibration/Matlab/gp12.json) is not part of this "
            "package; pass the calibration explicitly")
    if transformation_file is None:
        raise FileNotFoundError(
            "transformation_file: the reference's default (example_calibration/Matlab/world_transformation.json) "
            "does not exist in the reference tree either; pass the transformation explicitly")
    if not os.path.exists(output_folder):
        os.mkdir(output_folder)
    calibration = dl.load_camera_calibration(str(pathlib.Path(calibration_file)))
    transforms = dl.load_world_transformation(str(pathlib.Path(transformation_file)))
    rig = _derive_rig(calibration, transforms)
    engine._require_cuda()
    dev = torch.cuda.current_device()
    frames = list(frame_numbers)

    stages = []

    def one_color(color):
        # native reader -> vectorised selection -> one library call -> native writer; every stage
        # releases the GIL, so colours overlap (one ptk_ctx per worker thread)
        with torch.cuda.device(dev):
            t0 = time.perf_counter()
            data = _Table.read(input_folder + f"/rods_df_{color}.csv")
            t1 = time.perf_counter()
            m = _match_arrays(data, frames, rig, cam1_name, cam2_name, rematching, copy_rows=False)
            t2 = time.perf_counter()
            out_path = os.path.join(output_folder, f"rods_df_{color}.csv")
            if m is None:
                pd.DataFrame().to_csv(out_path, sep=",")
                return None
            _write_csv(out_path, color, m)
            stages.append((color, t1 - t0, t2 - t1, time.perf_counter() - t2))
            return (m["costs2d"], m["lens2d"]) if m["costs2d"] is not None else (m["costs"], m["lens"])

    colors = list(colors)
    if len(colors) > 1:
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=min(len(colors), _CSV_WORKERS)) as pool:
            results = list(pool.map(one_color, colors))
    else:
        results = [one_color(c) for c in colors]
    match_csv_complex.last_stage_seconds = {c: dict(read=a, match=b, write=w) for c, a, b, w in stages}
    results = [r for r in results if r is not None]
    if results and all(isinstance(r[0], np.ndarray) for r in results) and len({r[0].shape[1] for r in results}) == 1:
        return np.concatenate([r[0] for r in results]), np.concatenate([r[1] for r in results])
    # ragged frames: np.array over per-frame arrays, as the reference builds it (match2D.py:631)
    all_repr_errs = [c for r in results for c in list(r[0])]
    all_rod_lengths = [l for r in results for l in list(r[1])]
    return np.array(all_repr_errs), np.array(all_rod_lengths)


def _reorder_tensors(data: pd.DataFrame, frame_numbers, cam1_name: str, cam2_name: str):
    """Frames x particles tensors for the re-ordering kernel: per frame the rows sorted by
    ``particle`` (match2D.py:1057,1064); every frame must hold the same particle ids."""
    fidx = _FrameIndex(data)
    pcol = data["particle"].to_numpy()
    idx = []
    ref_ids = None
    for f in frame_numbers:
        rows = fidx.rows(f)
        rows = rows[np.argsort(pcol[rows], kind="stable")]
        ids = pcol[rows]
        if ref_ids is None:
            ref_ids = ids
        elif len(ids) != len(ref_ids) or (ids != ref_ids).any():
            raise ValueError("reorder_endpoints needs the same particle ids in every frame")
        idx.append(rows)
    idx = np.stack(idx)  # [F, N] positions into data
    ends = data[COLS_3D[:6]].to_numpy(dtype=np.float64)[idx].reshape(len(idx), -1, 2, 3)
    c2d = [*_cam_cols(cam1_name), *_cam_cols(cam2_name)]
    pts = data[c2d].to_numpy(dtype=np.float64)[idx].reshape(len(idx), -1, 2, 2, 2)
    return idx, ends, pts, c2d


def reorder_endpoints(data: pd.DataFrame, frame_numbers, cam1_name: str = "gp1", cam2_name: str = "gp2"):
    """DataFrame form of ``reorder_endpoints_csv``: returns ``(df_out, mincosts_T)``."""
    engine._require_cuda()
    frames = list(frame_numbers)
    idx, ends, pts, c2d = _reorder_tensors(data, frames, cam1_name, cam2_name)
    dev = torch.device("cuda", torch.cuda.current_device())
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[None])).to(dev)
    out_ends, out_2d, _, mincost = engine.reorder_endpoints(t(ends), t(pts))
    flat = idx.reshape(-1)
    df_out = data.iloc[flat].copy()
    df_out[COLS_3D[:6]] = out_ends[0].cpu().numpy().reshape(len(flat), 6)
    df_out[c2d] = out_2d[0].cpu().numpy().reshape(len(flat), 8)
    df_out.reset_index(drop=True, inplace=True)
    # mincosts_T[particle, frame - frame_numbers[1]] (match2D.py:1058,1100)
    return df_out, mincost[0].cpu().numpy().T.copy()


def reorder_endpoints_csv(input_folder: str, output_folder: str, colors: Iterable[str], cam1_name: str = "gp1",
                          cam2_name: str = "gp2", frame_numbers: Iterable[int] = None):
    """Reorders rod end points in ``rods_df_{color}.csv`` files such that their displacement
    between consecutive frames is minimal (reference match2D.py:1001-1146): one device scan per
    colour instead of O(frames x particles) pandas look-ups.  Writes files of the same layout and
    returns the costs of the last colour, ``mincosts_T[particle, frame]``, as the reference does.

    The reference hard-codes the 2D column names ``*_gp1`` / ``*_gp2`` (match2D.py:1083-1091);
    here ``cam1_name`` / ``cam2_name`` are honoured (identical for the default names)."""
    if not os.path.exists(output_folder):
        os.mkdir(output_folder)
    mincosts_T = None
    for color in colors:
        data = csv_io.read_rods_csv(input_folder + f"/rods_df_{color}.csv")
        df_out, mincosts_T = reorder_endpoints(data, frame_numbers, cam1_name, cam2_name)
        df_out.to_csv(os.path.join(output_folder, f"rods_df_{color}.csv"), sep=",")
    return mincosts_T
