"""``tracking.tracking_global_assignment`` of the reference (``reconstruct_3D/tracking.py``
:70-138) on the B200: the assignments of all consecutive frame pairs are solved in ``libptk.so``
straight from the end points (K2T: no cost matrix; its separable part lives in registers); the
DataFrame glue mirrors the reference, including its positional ``particle`` rewrite.

``tracking_trackpy`` is a thin wrapper over the third-party ``trackpy`` and is out of scope.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import pandas as pd
import torch

from .. import engine


def _ends(data: pd.DataFrame):
    frames = data["frame"].unique()
    F = len(frames)
    # tracking.py:97-98 -- raises ValueError (from reshape) when frames hold unequal counts
    p1 = data[["x1", "y1", "z1"]].to_numpy(dtype=np.float64).reshape((F, -1, 3))
    p2 = data[["x2", "y2", "z2"]].to_numpy(dtype=np.float64).reshape((F, -1, 3))
    return frames, np.stack([p1, p2], axis=2)  # [F,N,2,3]


def _track(ends: np.ndarray):
    engine._require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    col, tot, status, _ = engine.track_batch(torch.from_numpy(np.ascontiguousarray(ends[None])).to(dev))
    status = status.cpu().numpy()
    if (status == 1).any():
        raise ValueError("matrix contains invalid numeric entries")  # scipy, tracking.py:122
    if (status == 2).any():
        raise ValueError("cost matrix is infeasible")
    return col, tot


def _assign(data: pd.DataFrame):
    """The body shared by both entry points: one pass of the tracking kernels.  Returns the
    reference-shaped DataFrame, the per-pair costs and the device tensor col_ind[1,F-1,N] (None for
    fewer than two frames)."""
    frames, ends = _ends(data)
    F, N = ends.shape[:2]
    out = data.copy()
    if "particle" not in out.columns:
        out["particle"] = 0
        first = (out.frame == frames[0]).to_numpy()
        out.loc[first, "particle"] = np.arange(int(first.sum()))
    if F < 2:
        return out.sort_values(by=["frame", "particle"]).reset_index(drop=True), np.zeros(0), None, N
    col, tot = _track(ends)
    later = (out.frame != frames[0]).to_numpy()
    pos_in_frame = out.groupby("frame", sort=False).cumcount().to_numpy()
    part = out["particle"].to_numpy().copy()
    part[later] = pos_in_frame[later]
    out["particle"] = part
    out = out.sort_values(by=["frame", "particle"]).reset_index(drop=True)
    return out, tot.cpu().numpy()[0], col, N


def tracking_global_assignment(data: pd.DataFrame) -> Tuple[pd.DataFrame, np.ndarray]:
    """Tracks rods (one colour) over multiple frames with optimal assignment (reference
    tracking.py:70-138).  Returns the data with the reference's ``particle`` numbers (frame 1
    keeps its ids, later frames get their row position -- that is what tracking.py:130-133
    does) sorted by (frame, particle), and the assignment cost per frame pair.

    Deviation: two frames (one pair) work here; the reference raises IndexError there because
    of a ``squeeze`` (tracking.py:123)."""
    out, total, _, _ = _assign(data)
    return out, total


def tracking_global_assignment_chained(data: pd.DataFrame):
    """Extension (north star "chain pass"; no reference counterpart):
    like ``tracking_global_assignment`` but also returns ``col_ind[F-1,N]`` and
    ``track_id[F,N]`` with ids[0] = arange(N), ids[f+1][col_ind[f][a]] = ids[f][a].  One pass of the
    tracking kernels serves both results."""
    out, total, col, N = _assign(data)
    if col is None:
        return out, total, np.zeros((0, N), dtype=np.int32), np.arange(N, dtype=np.int32)[None]
    tid = engine.chain_tracks(col)
    return out, total, col.cpu().numpy()[0], tid.cpu().numpy()[0]
