"""Tensor-level API of the CUDA hot path: PyTorch owns device memory and streams,
``libptk.so`` does the work (``include/ptk.h``).  Everything the reference-shaped
functions in ``reconstruct_3D/`` do goes through here.

No CPU fallback: every function raises if the tensors are not CUDA tensors / the
library is missing.
"""
from __future__ import annotations

import ctypes
import dataclasses
from typing import Optional

import numpy as np
import torch

from . import _native as nat
from .rig import PtkRig

_f64, _i32, _u8 = torch.float64, torch.int32, torch.uint8


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("particletracking_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


class _Workspace:
    """Scratch per (thread, device, CUDA stream).  The callers run one worker per colour
    and a worker may use several torch streams: kernels enqueued on different streams must never
    share scratch, so the stream is part of the key.  A buffer that is replaced by a larger one is
    handed back to the caching allocator only after ``record_stream`` -- the kernels that still read
    it were enqueued on that stream, which need not be the one the tensor was allocated under."""

    def __init__(self):
        import threading

        self._tls = threading.local()

    def get(self, nbytes: int, device) -> torch.Tensor:
        stream = torch.cuda.current_stream(device)
        bufs = getattr(self._tls, "bufs", None)
        if bufs is None:
            bufs = self._tls.bufs = {}
        key = (str(device), stream.cuda_stream)
        buf = bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            if buf is not None:
                buf.record_stream(stream)
            with torch.cuda.stream(stream):
                buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
            bufs[key] = buf
        return buf


_ws = _Workspace()


@dataclasses.dataclass
class MatchResult:
    rows: torch.Tensor        # [P,Nmax,18] float64, NaN beyond n_out
    match_cost: torch.Tensor  # [P,Nmax]
    row_ind: torch.Tensor     # [P,Nmax] int32, -1 beyond n_out
    col_ind: torch.Tensor
    n_out: torch.Tensor       # [P] int32
    status: torch.Tensor      # [P] int32 (PTK_PROBLEM_*)
    cost: Optional[torch.Tensor] = None    # [P,Nmax,Nmax] if requested
    choice: Optional[torch.Tensor] = None  # [P,Nmax,Nmax] uint8 if requested
    keep: Optional[torch.Tensor] = None    # [P,Nmax] uint8 if max_repr_err given


def undistort_batch(rig: PtkRig, cam1, cam2, n1, n2):
    """K0: quirk-exact undistortion (reference match2D.py:846-863)."""
    _require_cuda()
    cam1, cam2 = _chk(cam1, _f64, "cam1"), _chk(cam2, _f64, "cam2")
    n1, n2 = _chk(n1, _i32, "n1"), _chk(n2, _i32, "n2")
    P, Nmax = cam1.shape[:2]
    u1 = torch.full_like(cam1, float("nan"))
    u2 = torch.full_like(cam2, float("nan"))
    with torch.cuda.device(cam1.device):
        nat.check(nat.lib().ptk_undistort_batch(ctypes.byref(rig), P, Nmax, _ptr(cam1), _ptr(cam2), _ptr(n1), _ptr(n2),
                                                _ptr(u1), _ptr(u2), _stream()), "ptk_undistort_batch")
    return u1, u2


def cost_batch(rig: PtkRig, cam1, cam2, n1, n2, want_choice=True, want_combos=False):
    """K0+K1: cost[P,Nmax,Nmax] (+choice, +per-combo world points / reprojection errors)."""
    _require_cuda()
    cam1, cam2 = _chk(cam1, _f64, "cam1"), _chk(cam2, _f64, "cam2")
    n1, n2 = _chk(n1, _i32, "n1"), _chk(n2, _i32, "n2")
    P, Nmax = cam1.shape[:2]
    dev = cam1.device
    cost = torch.full((P, Nmax, Nmax), float("nan"), dtype=_f64, device=dev)
    choice = torch.zeros((P, Nmax, Nmax), dtype=_u8, device=dev) if want_choice else None
    pw = torch.full((P, Nmax, Nmax, 4, 3), float("nan"), dtype=_f64, device=dev) if want_combos else None
    re = torch.full((P, Nmax, Nmax, 4, 2), float("nan"), dtype=_f64, device=dev) if want_combos else None
    nb = nat.lib().ptk_workspace_bytes(nat.WS_COST, P, Nmax)
    ws = _ws.get(nb, dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().ptk_cost_batch(ctypes.byref(rig), P, Nmax, _ptr(cam1), _ptr(cam2), _ptr(n1), _ptr(n2),
                                           _ptr(cost), _ptr(choice), _ptr(pw), _ptr(re), _ptr(ws), ws.numel(), _stream()),
                  "ptk_cost_batch")
    return cost, choice, pw, re


def lsap_batch(cost: torch.Tensor, nr: Optional[torch.Tensor] = None, nc: Optional[torch.Tensor] = None):
    """K2: batched scipy-exact linear_sum_assignment.  cost[P,R,C] -> row_ind, col_ind, n_out, status."""
    _require_cuda()
    cost = _chk(cost, _f64, "cost")
    P, R, C = cost.shape
    dev = cost.device
    k = min(R, C)
    row = torch.empty((P, k), dtype=_i32, device=dev)
    col = torch.empty((P, k), dtype=_i32, device=dev)
    n_out = torch.empty((P,), dtype=_i32, device=dev)
    status = torch.empty((P,), dtype=_i32, device=dev)
    if nr is not None:
        nr = _chk(nr, _i32, "nr")
    if nc is not None:
        nc = _chk(nc, _i32, "nc")
    with torch.cuda.device(dev):
        nat.check(nat.lib().ptk_lsap_batch(P, R, C, _ptr(cost), _ptr(nr), _ptr(nc), _ptr(row), _ptr(col), _ptr(n_out),
                                           _ptr(status), _stream()), "ptk_lsap_batch")
    return row, col, n_out, status


def match_batch(rig: PtkRig, cam1, cam2, n1, n2, want_cost=False, max_repr_err: Optional[float] = None,
                out: Optional[MatchResult] = None) -> MatchResult:
    """K0..K3: the whole ``match_frame`` body (match2D.py:843-983) for P problems.  ``out``: a result of
    an earlier call with the same shapes and options whose tensors are overwritten instead of allocating
    new ones (callers that keep several steps in flight reuse a fixed set of buffers this way)."""
    _require_cuda()
    cam1, cam2 = _chk(cam1, _f64, "cam1"), _chk(cam2, _f64, "cam2")
    n1, n2 = _chk(n1, _i32, "n1"), _chk(n2, _i32, "n2")
    P, Nmax = cam1.shape[:2]
    if cam2.shape[:2] != (P, Nmax) or n1.numel() != P or n2.numel() != P:
        raise ValueError("cam1/cam2 must be [P,Nmax,2,2] with n1/n2 [P]")
    dev = cam1.device
    if out is not None and not want_cost and tuple(out.rows.shape) == (P, Nmax, nat.ROW_COLS) and out.rows.device == dev \
            and (out.keep is not None) == (max_repr_err is not None):
        rows, mc, row, col, n_out, status, cost, choice, keep = (out.rows, out.match_cost, out.row_ind, out.col_ind,
                                                                 out.n_out, out.status, None, None, out.keep)
    else:
        rows = torch.empty((P, Nmax, nat.ROW_COLS), dtype=_f64, device=dev)
        mc = torch.empty((P, Nmax), dtype=_f64, device=dev)
        row = torch.empty((P, Nmax), dtype=_i32, device=dev)
        col = torch.empty((P, Nmax), dtype=_i32, device=dev)
        n_out = torch.empty((P,), dtype=_i32, device=dev)
        status = torch.empty((P,), dtype=_i32, device=dev)
        cost = torch.full((P, Nmax, Nmax), float("nan"), dtype=_f64, device=dev) if want_cost else None
        choice = torch.zeros((P, Nmax, Nmax), dtype=_u8, device=dev) if want_cost else None
        keep = torch.empty((P, Nmax), dtype=_u8, device=dev) if max_repr_err is not None else None
    nb = nat.lib().ptk_workspace_bytes(nat.WS_MATCH, P, Nmax)
    ws = _ws.get(nb, dev)
    thr = float("inf") if max_repr_err is None else float(max_repr_err)
    with torch.cuda.device(dev):
        nat.check(nat.lib().ptk_match_batch(ctypes.byref(rig), P, Nmax, _ptr(cam1), _ptr(cam2), _ptr(n1), _ptr(n2),
                                            _ptr(cost), _ptr(choice), _ptr(row), _ptr(col), _ptr(n_out), _ptr(status),
                                            _ptr(rows), _ptr(mc), thr, _ptr(keep), _ptr(ws), ws.numel(), _stream()),
                  "ptk_match_batch")
    return MatchResult(rows, mc, row, col, n_out, status, cost, choice, keep)


def rows_batch(rig: PtkRig, cam1, cam2, n1, n2, col_ind, n_out):
    """K0+K3 for given assignments (``renumber=False``, match2D.py:876-884,941-983):
    rows[P,Nmax,18] and pair_cost[P,Nmax] for the pairs (m, col_ind[p,m]), m < n_out[p]."""
    _require_cuda()
    cam1, cam2 = _chk(cam1, _f64, "cam1"), _chk(cam2, _f64, "cam2")
    n1, n2 = _chk(n1, _i32, "n1"), _chk(n2, _i32, "n2")
    col_ind, n_out = _chk(col_ind, _i32, "col_ind"), _chk(n_out, _i32, "n_out")
    P, Nmax = cam1.shape[:2]
    dev = cam1.device
    rows = torch.empty((P, Nmax, nat.ROW_COLS), dtype=_f64, device=dev)
    pc = torch.empty((P, Nmax), dtype=_f64, device=dev)
    nb = nat.lib().ptk_workspace_bytes(nat.WS_COST, P, Nmax)
    ws = _ws.get(nb, dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().ptk_rows_batch(ctypes.byref(rig), P, Nmax, _ptr(cam1), _ptr(cam2), _ptr(n1), _ptr(n2),
                                           _ptr(col_ind), _ptr(n_out), _ptr(rows), _ptr(pc), _ptr(ws), ws.numel(),
                                           _stream()), "ptk_rows_batch")
    return rows, pc


def rows_to_ends(rows: torch.Tensor, F: int, C: int, N: Optional[int] = None) -> torch.Tensor:
    """rows[F*C,Nmax,18] (problem = f*C+c) -> ends[C,F,N,2,3]: the six columns
    tracking_global_assignment reads (tracking.py:97-98)."""
    _require_cuda()
    rows = _chk(rows, _f64, "rows")
    Nmax = rows.shape[-2]
    N = Nmax if N is None else N
    ends = torch.empty((C, F, N, 2, 3), dtype=_f64, device=rows.device)
    with torch.cuda.device(rows.device):
        nat.check(nat.lib().ptk_rows_to_ends(F, C, N, Nmax, _ptr(rows), _ptr(ends), _stream()), "ptk_rows_to_ends")
    return ends


def reconstruct_device(rig: PtkRig, cam1, cam2, n1, n2, F: int, C: int, track: bool = True,
                       max_repr_err: Optional[float] = None):
    """The whole hot path on device-resident tensors: match (K0-K3) -> ends -> track (K4+K2) ->
    chain (K5).  cam1/cam2 [F*C,Nmax,2,2] with problem index f*C+c.  Returns
    (MatchResult, col_ind[C,F-1,N], total_cost[C,F-1], track_id[C,F,N], track_status)."""
    m = match_batch(rig, cam1, cam2, n1, n2, max_repr_err=max_repr_err)
    if not track:
        return m, None, None, None, None
    ends = rows_to_ends(m.rows, F, C)
    col, tot, st, _ = track_batch(ends)
    tid = chain_tracks(col)
    return m, col, tot, tid, st


class DevicePipeline:
    """The whole path on device-resident tensors through ONE library call
    (``ptk_reconstruct_device``) with outputs and workspace allocated once for a given
    (F, C, Nmax): no per-step allocation, a single ctypes call per step, capturable in a CUDA
    graph.  This is what bench.py times as `value`."""

    def __init__(self, rig: PtkRig, F: int, C: int, Nmax: int, device=None, track: bool = True,
                 max_repr_err: Optional[float] = None):
        _require_cuda()
        self.rig, self.F, self.C, self.N, self.track = rig, F, C, Nmax, track
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        P = F * C
        self.rows = torch.empty((P, Nmax, nat.ROW_COLS), dtype=_f64, device=dev)
        self.match_cost = torch.empty((P, Nmax), dtype=_f64, device=dev)
        self.row_ind = torch.empty((P, Nmax), dtype=_i32, device=dev)
        self.col_ind = torch.empty((P, Nmax), dtype=_i32, device=dev)
        self.n_out = torch.empty((P,), dtype=_i32, device=dev)
        self.status = torch.empty((P,), dtype=_i32, device=dev)
        self.keep = torch.empty((P, Nmax), dtype=_u8, device=dev) if max_repr_err is not None else None
        self.thr = float("inf") if max_repr_err is None else float(max_repr_err)
        self.track_col = torch.empty((C, max(F - 1, 0), Nmax), dtype=_i32, device=dev) if track else None
        self.track_total = torch.empty((C, max(F - 1, 0)), dtype=_f64, device=dev) if track else None
        self.track_id = torch.empty((C, F, Nmax), dtype=_i32, device=dev) if track else None
        self.track_status = torch.empty((C, max(F - 1, 0)), dtype=_i32, device=dev) if track else None
        nb = nat.lib().ptk_reconstruct_device_workspace_bytes(F, C, Nmax, 1 if track else 0)
        self.ws = torch.empty(nb, dtype=_u8, device=dev)

    def _check_inputs(self, cam1, cam2, n1, n2):
        P = self.F * self.C
        for t, dt, name, shape in ((cam1, _f64, "cam1", (P, self.N, 2, 2)), (cam2, _f64, "cam2", (P, self.N, 2, 2)),
                                   (n1, _i32, "n1", (P,)), (n2, _i32, "n2", (P,))):
            if not isinstance(t, torch.Tensor) or not t.is_cuda or t.device != self.device:
                raise ValueError(f"{name} must be a CUDA tensor on {self.device}")
            if t.dtype != dt:
                raise ValueError(f"{name} must be {dt}, got {t.dtype}")
            if tuple(t.shape) != shape:
                raise ValueError(f"{name} must have shape {shape}, got {tuple(t.shape)}")
            if not t.is_contiguous():
                raise ValueError(f"{name} must be contiguous")

    def run(self, cam1, cam2, n1, n2):
        """Enqueue one pass on the current stream; results are in the pipeline's tensors."""
        self._check_inputs(cam1, cam2, n1, n2)
        with torch.cuda.device(self.device):
            nat.check(nat.lib().ptk_reconstruct_device(
                ctypes.byref(self.rig), self.F, self.C, self.N, _ptr(cam1), _ptr(cam2), _ptr(n1), _ptr(n2),
                _ptr(self.rows), _ptr(self.match_cost), _ptr(self.row_ind), _ptr(self.col_ind), _ptr(self.n_out),
                _ptr(self.status), self.thr, _ptr(self.keep), 1 if self.track else 0, _ptr(self.track_col),
                _ptr(self.track_total), _ptr(self.track_id), _ptr(self.track_status), _ptr(self.ws), self.ws.numel(),
                _stream()), "ptk_reconstruct_device")
        return self


class ShardPipeline:
    """One rank's frame range of a sharded run through ONE library call per step
    (``ptk_reconstruct_shard``): ``F`` own frames followed by ``n_halo`` (0 or 1) frames of the next rank
    that are re-computed locally, so that no halo exchange is needed.  Outputs are caller-provided
    tensors (``out``: the views of the rank's gather block, see ``distributed.ShardedReconstructor``) or
    allocated here.  ``run`` is eager; ``run_graph`` replays a CUDA graph of the same call captured the
    second time a set of input tensors is seen (the step is then a single ``cudaGraphLaunch``)."""

    OUT_SPEC = (  # name, dtype, shape as a function of (F, h, C, N)
        ("rows", _f64, lambda F, h, C, N: ((F + h) * C, N, nat.ROW_COLS)),
        ("match_cost", _f64, lambda F, h, C, N: ((F + h) * C, N)),
        ("n_out", _i32, lambda F, h, C, N: ((F + h) * C,)),
        ("status", _i32, lambda F, h, C, N: ((F + h) * C,)),
        ("track_total", _f64, lambda F, h, C, N: (C, F + h - 1)),
        ("track_status", _i32, lambda F, h, C, N: (C, F + h - 1)),
        ("track_id", _i32, lambda F, h, C, N: (C, F, N)),
        ("halo_map", _i32, lambda F, h, C, N: (C, N)),
    )

    def __init__(self, rig: PtkRig, F: int, n_halo: int, C: int, Nmax: int, device=None, out=None,
                 max_repr_err: Optional[float] = None, workspace: Optional[torch.Tensor] = None):
        _require_cuda()
        if n_halo not in (0, 1) or F < 1:
            raise ValueError("ShardPipeline: F >= 1 and n_halo in (0, 1)")
        self.rig, self.F, self.h, self.C, self.N = rig, F, n_halo, C, Nmax
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        out = {} if out is None else out
        for name, dt, shp in self.OUT_SPEC:
            shape = shp(F, n_halo, C, Nmax)
            t = out.get(name)
            if t is None:
                t = torch.empty(shape, dtype=dt, device=dev)
            elif tuple(t.shape) != tuple(shape) or t.dtype != dt or t.device != dev or not t.is_contiguous():
                raise ValueError(f"ShardPipeline: out[{name!r}] must be a contiguous {dt} tensor of shape {shape} on {dev}")
            setattr(self, name, t)
        P = (F + n_halo) * C
        self.row_ind = torch.empty((P, Nmax), dtype=_i32, device=dev)
        self.col_ind = torch.empty((P, Nmax), dtype=_i32, device=dev)
        self.track_col = torch.empty((C, F + n_halo - 1, Nmax), dtype=_i32, device=dev)
        self.keep = torch.empty((P, Nmax), dtype=_u8, device=dev) if max_repr_err is not None else None
        self.thr = float("inf") if max_repr_err is None else float(max_repr_err)
        nb = nat.lib().ptk_reconstruct_shard_workspace_bytes(F, n_halo, C, Nmax)
        if workspace is not None and (workspace.numel() < nb or workspace.device != dev):
            raise ValueError("ShardPipeline: workspace too small")
        self.ws = workspace if workspace is not None else torch.empty(nb, dtype=_u8, device=dev)
        self._seen, self._graphs, self._keep = {}, {}, []

    def _check_inputs(self, cam1, cam2, n1, n2):
        P = (self.F + self.h) * self.C
        for t, dt, name, shape in ((cam1, _f64, "cam1", (P, self.N, 2, 2)), (cam2, _f64, "cam2", (P, self.N, 2, 2)),
                                   (n1, _i32, "n1", (P,)), (n2, _i32, "n2", (P,))):
            if not isinstance(t, torch.Tensor) or not t.is_cuda or t.device != self.device:
                raise ValueError(f"{name} must be a CUDA tensor on {self.device}")
            if t.dtype != dt or tuple(t.shape) != shape or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous {dt} tensor of shape {shape} (F + n_halo frames), "
                                 f"got {t.dtype} {tuple(t.shape)}")

    def _enqueue(self, cam1, cam2, n1, n2):
        nat.check(nat.lib().ptk_reconstruct_shard(
            ctypes.byref(self.rig), self.F, self.h, self.C, self.N, _ptr(cam1), _ptr(cam2), _ptr(n1), _ptr(n2),
            _ptr(self.rows), _ptr(self.match_cost), _ptr(self.row_ind), _ptr(self.col_ind), _ptr(self.n_out),
            _ptr(self.status), self.thr, _ptr(self.keep), _ptr(self.track_col), _ptr(self.track_total),
            _ptr(self.track_status), _ptr(self.track_id), _ptr(self.halo_map) if self.h else None, _ptr(self.ws),
            self.ws.numel(), _stream()), "ptk_reconstruct_shard")

    def run(self, cam1, cam2, n1, n2):
        """Enqueue one pass on the current stream (eager launches)."""
        self._check_inputs(cam1, cam2, n1, n2)
        with torch.cuda.device(self.device):
            self._enqueue(cam1, cam2, n1, n2)
        return self

    def run_graph(self, cam1, cam2, n1, n2):
        """As ``run``; from the second call with the same input tensors on, one CUDA-graph launch."""
        key = (cam1.data_ptr(), cam2.data_ptr(), n1.data_ptr(), n2.data_ptr())
        g = self._graphs.get(key)
        if g is not None:
            g.replay()
            return self
        n = self._seen.get(key, 0)
        self._seen[key] = n + 1
        if n == 0 or _timing_on:  # first sight (one-time attribute set-up runs eagerly) / event timing on
            return self.run(cam1, cam2, n1, n2)
        self._check_inputs(cam1, cam2, n1, n2)
        g = torch.cuda.CUDAGraph()
        cur = torch.cuda.current_stream(self.device)
        cap = torch.cuda.Stream(self.device)
        cap.wait_stream(cur)
        with torch.cuda.device(self.device), torch.cuda.graph(g, stream=cap):
            self._enqueue(cam1, cam2, n1, n2)
        cur.wait_stream(cap)
        self._graphs[key] = g
        self._keep.append((cam1, cam2, n1, n2))  # the graph holds raw pointers: keep the tensors alive
        g.replay()
        return self


def chain_rebase_gathered(blocks: torch.Tensor, frames_per_rank, C: int, N: int, tid_off: int, map_off: int,
                          scratch: Optional[torch.Tensor] = None):
    """Root side of the sharded chain pass: ``blocks[W, block_bytes]`` uint8 (the gathered result blocks);
    re-bases every rank's local ids in place (``ptk_chain_rebase_gathered``)."""
    _require_cuda()
    if blocks.dtype != _u8 or blocks.dim() != 2 or not blocks.is_contiguous() or not blocks.is_cuda:
        raise ValueError("blocks must be a contiguous CUDA uint8 tensor [W, block_bytes]")
    W = blocks.shape[0]
    fr = (ctypes.c_int32 * W)(*[int(x) for x in frames_per_rank])
    if scratch is None or scratch.numel() < W:
        scratch = torch.empty((max(W, 1),), dtype=_i32, device=blocks.device)
    with torch.cuda.device(blocks.device):
        nat.check(nat.lib().ptk_chain_rebase_gathered(W, C, N, fr, _ptr(blocks), blocks.stride(0), int(tid_off), int(map_off),
                                                      _ptr(scratch), _stream()), "ptk_chain_rebase_gathered")
    return blocks


def track_batch(ends: torch.Tensor, want_cost=False):
    """K4+K2: ``tracking_global_assignment`` (tracking.py:97-123,136-137) for C colours.
    ends[C,F,N,2,3] -> col_ind[C,F-1,N], total_cost[C,F-1], status[C,F-1] (, cost[C,F-1,N,N])."""
    _require_cuda()
    ends = _chk(ends, _f64, "ends")
    C, F, N = ends.shape[:3]
    dev = ends.device
    col = torch.empty((C, max(F - 1, 0), N), dtype=_i32, device=dev)
    tot = torch.empty((C, max(F - 1, 0)), dtype=_f64, device=dev)
    status = torch.empty((C, max(F - 1, 0)), dtype=_i32, device=dev)
    cost = torch.empty((C, max(F - 1, 0), N, N), dtype=_f64, device=dev) if want_cost else None
    nb = nat.lib().ptk_workspace_bytes(nat.WS_TRACK, max(C * (F - 1), 1), N)
    ws = _ws.get(nb, dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().ptk_track_batch(C, F, N, _ptr(ends), _ptr(cost), _ptr(col), _ptr(tot), _ptr(status),
                                            _ptr(ws), ws.numel(), _stream()), "ptk_track_batch")
    return col, tot, status, cost


def chain_tracks(col_ind: torch.Tensor, ids_in: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K5 (extension): col_ind[C,F-1,N] -> track_id[C,F,N]."""
    _require_cuda()
    col_ind = _chk(col_ind, _i32, "col_ind")
    C, Fm1, N = col_ind.shape
    out = torch.empty((C, Fm1 + 1, N), dtype=_i32, device=col_ind.device)
    if ids_in is not None:
        ids_in = _chk(ids_in, _i32, "ids_in")
    with torch.cuda.device(col_ind.device):
        nat.check(nat.lib().ptk_chain_tracks(C, Fm1 + 1, N, _ptr(col_ind), _ptr(ids_in), _ptr(out), _stream()),
                  "ptk_chain_tracks")
    return out


def chain_rebase(maps: torch.Tensor, rank: int, track_id: torch.Tensor) -> torch.Tensor:
    """Multi-GPU chain pass: compose the all-gathered boundary maps[W,C,N] up to ``rank`` and
    re-base the local ids track_id[C,F,N] in place; returns the rank's start ids [C,N]."""
    _require_cuda()
    maps, track_id = _chk(maps, _i32, "maps"), _chk(track_id, _i32, "track_id")
    W, C, N = maps.shape
    F = track_id.shape[1]
    start = torch.empty((C, N), dtype=_i32, device=maps.device)
    with torch.cuda.device(maps.device):
        nat.check(nat.lib().ptk_chain_rebase(W, C, F, N, int(rank), _ptr(maps), _ptr(start), _ptr(track_id), _stream()),
                  "ptk_chain_rebase")
    return start


def create_weights(p_3D: torch.Tensor, p_prev: torch.Tensor, rep_errs: torch.Tensor):
    """K6: ``matchND.create_weights`` (matchND.py:141-217)."""
    _require_cuda()
    p_3D, p_prev, rep_errs = _chk(p_3D, _f64, "p_3D"), _chk(p_prev, _f64, "p_prev"), _chk(rep_errs, _f64, "rep_errs")
    N1, N2 = p_3D.shape[:2]
    Np = p_prev.shape[0]
    dev = p_3D.device
    w = torch.empty((Np, N1, N2), dtype=_f64, device=dev)
    ch = torch.empty((Np, N1, N2), dtype=_f64, device=dev)
    nb = nat.lib().ptk_workspace_bytes(nat.WS_WEIGHTS, Np * N1 * N2, 0)
    ws = _ws.get(nb, dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().ptk_create_weights(Np, N1, N2, _ptr(p_3D), _ptr(p_prev), _ptr(rep_errs), _ptr(w), _ptr(ch),
                                               _ptr(ws), ws.numel(), _stream()), "ptk_create_weights")
    return w, ch


class Npartite3Result:
    """sol[B,3,Dmin] (-1 padded), n_sol[B], value[B], bound[B], status[B], stats[B,4]."""

    def __init__(self, sol, n_sol, value, bound, status, stats):
        self.sol, self.n_sol, self.value, self.bound, self.status, self.stats = sol, n_sol, value, bound, status, stats


def npartite3_batch(weights: torch.Tensor, n0=None, n1=None, n2=None, max_steps: int = 0, node_limit: int = 0) -> Npartite3Result:
    """K10: the 3-group ``matchND.npartite_matching`` (matchND.py:45-138) for ``weights[B,D0,D1,D2]``."""
    _require_cuda()
    weights = _chk(weights, _f64, "weights")
    if weights.dim() != 4:
        raise ValueError("weights must be [B, D0, D1, D2]")
    B, D0, D1, D2 = weights.shape
    if max(D0, D1, D2) > nat.AP3_MAXN:
        raise ValueError(f"npartite3: more than {nat.AP3_MAXN} members in a group")
    dev = weights.device
    dmin = min(D0, D1, D2)
    sol = torch.empty((B, 3, dmin), dtype=_i32, device=dev)
    n_sol = torch.empty((B,), dtype=_i32, device=dev)
    value = torch.empty((B,), dtype=_f64, device=dev)
    bound = torch.empty((B,), dtype=_f64, device=dev)
    status = torch.empty((B,), dtype=_i32, device=dev)
    stats = torch.empty((B, 4), dtype=_i32, device=dev)
    ns = [None if x is None else _chk(x, _i32, "n") for x in (n0, n1, n2)]
    nb = nat.lib().ptk_workspace_bytes(nat.WS_AP3, B, 0)
    ws = _ws.get(nb, dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().ptk_npartite3_batch(B, D0, D1, D2, _ptr(weights), _ptr(ns[0]), _ptr(ns[1]), _ptr(ns[2]), _ptr(sol),
                                                _ptr(n_sol), _ptr(value), _ptr(bound), _ptr(status), _ptr(stats),
                                                int(max_steps), int(node_limit), _ptr(ws), ws.numel(), _stream()),
                  "ptk_npartite3_batch")
    return Npartite3Result(sol, n_sol, value, bound, status, stats)


def matchnd_sequence(rig: PtkRig, cam1, cam2, n1, n2):
    """``matchND.assign``'s frame loop (matchND.py:318-360) for ``cam1, cam2 [F,C,Nmax,2,2]``,
    ``n1, n2 [F,C]`` on the device: returns ``rows[F,C,Nmax,18], costs[F,C,Nmax], n_rows[F,C],
    status[F,C], solver_stats[C,4]``."""
    _require_cuda()
    cam1, cam2 = _chk(cam1, _f64, "cam1"), _chk(cam2, _f64, "cam2")
    n1, n2 = _chk(n1, _i32, "n1"), _chk(n2, _i32, "n2")
    F, C, Nmax = cam1.shape[:3]
    if Nmax > nat.AP3_MAXN:
        raise ValueError(f"matchND: more than {nat.AP3_MAXN} rods per (frame, colour)")
    dev = cam1.device
    rows = torch.empty((F, C, Nmax, nat.ROW_COLS), dtype=_f64, device=dev)
    costs = torch.empty((F, C, Nmax), dtype=_f64, device=dev)
    n_rows = torch.empty((F, C), dtype=_i32, device=dev)
    status = torch.empty((F, C), dtype=_i32, device=dev)
    stats = torch.zeros((C, 4), dtype=_i32, device=dev)
    nb = nat.lib().ptk_matchnd_sequence_workspace_bytes(F, C, Nmax)
    ws = _ws.get(nb, dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().ptk_matchnd_sequence(ctypes.byref(rig), F, C, Nmax, _ptr(cam1), _ptr(cam2), _ptr(n1), _ptr(n2),
                                                 _ptr(rows), _ptr(costs), _ptr(n_rows), _ptr(status), _ptr(stats), _ptr(ws),
                                                 ws.numel(), _stream()), "ptk_matchnd_sequence")
    return rows, costs, n_rows, status, stats


def reorder_endpoints(ends: torch.Tensor, pts2d: Optional[torch.Tensor] = None):
    """K7: ``match2D.reorder_endpoints_csv``'s numeric body (match2D.py:1055-1137).
    ends[C,F,N,2,3] (, pts2d[C,F,N,2,2,2]) -> (out_ends, out_2d | None, swapped[C,F,N] uint8,
    mincost[C,F-1,N])."""
    _require_cuda()
    ends = _chk(ends, _f64, "ends")
    C, F, N = ends.shape[:3]
    dev = ends.device
    out_ends = torch.empty_like(ends)
    out_2d = None
    if pts2d is not None:
        pts2d = _chk(pts2d, _f64, "pts2d")
        out_2d = torch.empty_like(pts2d)
    swapped = torch.zeros((C, F, N), dtype=_u8, device=dev)
    mincost = torch.empty((C, max(F - 1, 0), N), dtype=_f64, device=dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().ptk_reorder_endpoints(C, F, N, _ptr(ends), _ptr(pts2d), _ptr(out_ends), _ptr(out_2d),
                                                  _ptr(swapped), _ptr(mincost), _stream()), "ptk_reorder_endpoints")
    return out_ends, out_2d, swapped, mincost


def triangulate_points(rig: PtkRig, p1: torch.Tensor, p2: torch.Tensor, to_world: bool = False) -> torch.Tensor:
    """K8: ``calibrate_cameras.project_points`` (calibrate_cameras.py:137-199). p1, p2 [n,2] -> [n,3]."""
    _require_cuda()
    p1, p2 = _chk(p1, _f64, "p1"), _chk(p2, _f64, "p2")
    n = p1.shape[0]
    out = torch.empty((n, 3), dtype=_f64, device=p1.device)
    with torch.cuda.device(p1.device):
        nat.check(nat.lib().ptk_triangulate_points(ctypes.byref(rig), n, _ptr(p1), _ptr(p2), 1 if to_world else 0,
                                                   _ptr(out), _stream()), "ptk_triangulate_points")
    return out


def project_points(rig: PtkRig, X: torch.Tensor, from_world: bool = False):
    """K9: ``calibrate_cameras.reproject_points`` (calibrate_cameras.py:202-262). X [n,3] -> uv1, uv2 [n,2]."""
    _require_cuda()
    X = _chk(X, _f64, "X")
    n = X.shape[0]
    uv1 = torch.empty((n, 2), dtype=_f64, device=X.device)
    uv2 = torch.empty((n, 2), dtype=_f64, device=X.device)
    with torch.cuda.device(X.device):
        nat.check(nat.lib().ptk_project_points(ctypes.byref(rig), n, _ptr(X), 1 if from_world else 0, _ptr(uv1), _ptr(uv2),
                                               _stream()), "ptk_project_points")
    return uv1, uv2


def measure_fp64_peak(iters: int = 20000) -> float:
    """FLOP/s of a dependency-free DFMA loop on the current device (roofline denominator)."""
    _require_cuda()
    out = ctypes.c_double(0.0)
    nat.check(nat.lib().ptk_measure_fp64_peak(int(iters), ctypes.byref(out), _stream()), "ptk_measure_fp64_peak")
    return out.value


_timing_on = False


def kernel_timing(enable) -> None:
    """Switch CUDA-event timing around kernel launches: False / 0 off, True / 1 every launch, 2 the dominant
    kernel (K1) only (graph replays are not timed: ``ShardPipeline.run_graph`` launches eagerly while it is on)."""
    global _timing_on
    _timing_on = bool(enable)
    nat.lib().ptk_timing_enable(2 if enable == 2 else (1 if enable else 0))


TIMING_KINDS = ["undistort", "cost", "lsap_match", "rows", "rows_to_ends", "track_cost", "lsap_track", "chain", "weights", "npartite3"]


def kernel_timing_read(reset: bool = True):
    """{kind: (total ms, launches)} since the last reset; synchronise first."""
    n = len(TIMING_KINDS)
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_longlong * n)()
    nat.check(nat.lib().ptk_timing_read(ms, cnt, 1 if reset else 0), "ptk_timing_read")
    return {k: (ms[i], cnt[i]) for i, k in enumerate(TIMING_KINDS)}


def launch_count(reset: bool = False) -> int:
    return int(nat.lib().ptk_launch_count(1 if reset else 0))


# ----------------------------------------------------------------------------------------
class HostPipeline:
    """Host-buffer entry point (``ptk_reconstruct_host``): numpy / pinned-torch arrays in,
    arrays out, copies chunked over frames and overlapped with compute inside the library.
    This is what ``match_csv_complex`` and bench.py's ``e2e`` use."""

    def __init__(self, device: int = 0):
        _require_cuda()
        self._ctx = ctypes.c_void_p()
        nat.check(nat.lib().ptk_ctx_create(int(device), ctypes.byref(self._ctx)), "ptk_ctx_create")
        self.device = device
        self._out = {}

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            nat.lib().ptk_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _buf(self, name, shape, dtype):
        """Pinned output buffers, reused across calls of the same shape."""
        t = self._out.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, pin_memory=True)
            self._out[name] = t
        return t

    def reconstruct(self, rig: PtkRig, cam1, cam2, n1, n2, track: bool = False, max_repr_err: Optional[float] = None,
                    raise_on_invalid: bool = True, compact: bool = False):
        """cam1, cam2: [F,C,Nmax,2,2] float64 host (numpy or CPU torch, ideally pinned);
        n1, n2: [F,C] int32.  Returns a dict of pinned CPU torch tensors (views reused by the
        next call -- copy what you keep).

        ``compact``: instead of ``rows[F,C,Nmax,18]`` the result carries ``ends[F,C,Nmax,6]`` (x1..z2) and
        ``straight[F,C,Nmax]``: 12.5 MB come back instead of 36.9 MB per 8000 x 32 problems, which is what
        bounds this path on a multi-GPU host.  ``expand_rows(result, cam1, cam2)`` rebuilds the 18 columns
        bit for bit (``ptk_reconstruct_host_compact`` / ``ptk_expand_rows_host``)."""
        c1, c2 = _as_host(cam1, _f64), _as_host(cam2, _f64)
        m1, m2 = _as_host(n1, _i32), _as_host(n2, _i32)
        F, C, Nmax = c1.shape[:3]
        rows = ends = straight = None
        if compact:
            ends = self._buf("ends", (F, C, Nmax, 6), _f64)
            straight = self._buf("straight", (F, C, Nmax), _u8)
        else:
            rows = self._buf("rows", (F, C, Nmax, nat.ROW_COLS), _f64)
        mc = self._buf("mc", (F, C, Nmax), _f64)
        row = self._buf("row", (F, C, Nmax), _i32)
        col = self._buf("col", (F, C, Nmax), _i32)
        n_out = self._buf("n_out", (F, C), _i32)
        status = self._buf("status", (F, C), _i32)
        keep = self._buf("keep", (F, C, Nmax), _u8) if max_repr_err is not None else None
        tcol = ttot = tid = tst = None
        if track:
            tcol = self._buf("tcol", (C, max(F - 1, 0), Nmax), _i32)
            ttot = self._buf("ttot", (C, max(F - 1, 0)), _f64)
            tid = self._buf("tid", (C, F, Nmax), _i32)
            tst = self._buf("tst", (C, max(F - 1, 0)), _i32)
        thr = float("inf") if max_repr_err is None else float(max_repr_err)
        tail = (_ptr(mc), _ptr(row), _ptr(col), _ptr(n_out), _ptr(status), thr, _ptr(keep), 1 if track else 0, _ptr(tcol),
                _ptr(ttot), _ptr(tid), _ptr(tst))
        head = (self._ctx, ctypes.byref(rig), F, C, Nmax, _ptr(c1), _ptr(c2), _ptr(m1), _ptr(m2))
        if compact:
            rc = nat.lib().ptk_reconstruct_host_compact(*head, _ptr(ends), _ptr(straight), *tail)
        else:
            rc = nat.lib().ptk_reconstruct_host(*head, _ptr(rows), *tail)
        if rc == nat.PTK_ERR_NUMERIC:
            if raise_on_invalid:
                # scipy raises ValueError from linear_sum_assignment (match2D.py:933, tracking.py:122)
                raise ValueError("matrix contains invalid numeric entries")
        else:
            nat.check(rc, "ptk_reconstruct_host")
        return dict(rows=rows, ends=ends, straight=straight, match_cost=mc, row_ind=row, col_ind=col, n_out=n_out, status=status,
                    keep=keep, track_col=tcol, track_total=ttot, track_id=tid, track_status=tst)


def expand_rows(result: dict, cam1, cam2, out: Optional[torch.Tensor] = None, n_threads: int = 0) -> torch.Tensor:
    """The 18-column rows ``[F,C,Nmax,18]`` of a ``compact`` result (host arrays; ``ptk_expand_rows_host``)."""
    c1, c2 = _as_host(cam1, _f64), _as_host(cam2, _f64)
    ends, straight, col, n_out = result["ends"], result["straight"], result["col_ind"], result["n_out"]
    F, C, Nmax = c1.shape[:3]
    if out is None:
        out = torch.empty((F, C, Nmax, nat.ROW_COLS), dtype=_f64)
    elif tuple(out.shape) != (F, C, Nmax, nat.ROW_COLS) or out.dtype != _f64 or not out.is_contiguous() or out.is_cuda:
        raise ValueError("expand_rows: out must be a contiguous float64 host tensor [F,C,Nmax,18]")
    nat.check(nat.lib().ptk_expand_rows_host(F * C, Nmax, _ptr(c1), _ptr(c2), _ptr(ends), _ptr(straight), _ptr(col), _ptr(n_out),
                                            _ptr(out), int(n_threads)), "ptk_expand_rows_host")
    return out


def _as_host(a, dtype) -> torch.Tensor:
    if isinstance(a, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(a))
    else:
        t = a
    if t.is_cuda:
        raise ValueError("HostPipeline takes host buffers")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()
