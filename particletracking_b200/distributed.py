"""Frame sharding across the GPUs of one box.

One process per GPU (``torch.distributed``, backend ``nccl``; ``gloo`` in the CPU tests).
The path shards naturally:

* matching: every (frame, colour) problem is independent -> contiguous frame ranges per
  rank, no exchange;
* tracking: the pair (f, f+1) needs only those two frames' rows -> same partition plus a
  1-frame halo: rank r receives the first frame's end points of rank r+1 (``[C,N,2,3]``,
  a few KB) and solves the boundary pair itself;
* chain pass: permutation composition is associative -> every rank chains locally from the
  identity, the per-rank boundary maps (``[C,N]`` int32) are all-gathered, composed
  (``compose_rank_maps``), and the local ids are re-based with the rank's start ids;
* final collective: gather of the 18-column rows and track ids to rank 0.

There is no data-path collective inside matching or the per-pair assignment; NCCL carries
only the halo, the boundary maps and the final gather.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_frames(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous frame range [f0, f1) of ``rank``; sizes differ by at most one."""
    base, rem = divmod(n_frames, world)
    f0 = rank * base + min(rank, rem)
    return f0, f0 + base + (1 if rank < rem else 0)


def compose_rank_maps(maps: torch.Tensor) -> torch.Tensor:
    """maps[W,C,N]: for rank r, position x in rank r+1's first frame -> local id (position in
    rank r's first frame).  Returns starts[W,C,N]: global track id of local id x on rank r:
    starts[0] = arange(N), starts[r+1][c][x] = starts[r][c][maps[r][c][x]]."""
    W, C, N = maps.shape
    starts = torch.empty_like(maps)
    cur = torch.arange(N, dtype=maps.dtype, device=maps.device).expand(C, N).contiguous()
    for r in range(W):
        starts[r] = cur
        if r + 1 < W:
            cur = torch.gather(cur, 1, maps[r].long())
    return starts


class _CudaOps:
    """The product's compute: libptk.so kernels through particletracking_b200.engine."""

    def __init__(self, rig):
        from . import engine

        self.e = engine
        self.rig = rig
        self._out = {}

    def match(self, cam1, cam2, n1, n2, slot=None):
        # the rows of step k are still being gathered while step k+1 computes: two fixed sets of
        # output tensors (one per gather slot) instead of a fresh 37 MB allocation per step
        prev = self._out.get(slot) if slot is not None else None
        m = self.e.match_batch(self.rig, cam1, cam2, n1, n2, out=prev)
        if slot is not None:
            self._out[slot] = m
        return m.rows, m.col_ind, m.status

    def rows_to_ends(self, rows, F, C):
        return self.e.rows_to_ends(rows, F, C)

    def track(self, ends):
        col, tot, st, _ = self.e.track_batch(ends)
        return col, tot, st

    def chain(self, col, ids_in=None):
        return self.e.chain_tracks(col, ids_in)

    def rebase(self, maps, rank, local):
        """compose the gathered boundary maps and re-base ``local`` [C,F',N] in place"""
        self.e.chain_rebase(maps, rank, local)
        return local


class ShardedReconstructor:
    """match -> track -> chain for one rank's frame range, with the halo / boundary-map /
    gather exchanges described in the module docstring.

    ``ops`` is the compute back end (default: the CUDA kernels).  The CPU tests inject an
    oracle-backed one to exercise the exchange logic under ``gloo`` -- test infrastructure,
    not a fallback: the default never touches it.
    """

    def __init__(self, rig, n_frames_local: int, n_colors: int, n_rods: int, device, ops=None,
                 gather_to_root: bool = True, group=None, n_frames_total: Optional[int] = None):
        self.F, self.C, self.N = n_frames_local, n_colors, n_rods
        self.device = device
        self.ops = ops if ops is not None else _CudaOps(rig)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.gather_to_root = gather_to_root
        # frames held by every rank (shards may differ by one frame; gathers are padded to the max)
        total = n_frames_total if n_frames_total is not None else n_frames_local * self.world
        self.shard_sizes = [b - a for a, b in (shard_frames(total, self.world, r) for r in range(self.world))]
        if self.shard_sizes[self.rank] != n_frames_local:
            raise ValueError("n_frames_local does not match shard_frames(n_frames_total, world, rank)")
        # final-gather pipelining: two sets of receive buffers; the gather of step k runs on
        # NCCL's stream while step k+1 computes, and is waited for before its buffers are reused
        self._rows_all0 = self._rows_all1 = None
        self._tid_all0 = self._tid_all1 = None
        self._pending = [None, None]
        self._step_idx = 0
        # host-side time per segment of step() (seconds, summed), filled when PTK_STEP_PROFILE=1
        import os as _os

        self._prof = {} if _os.environ.get("PTK_STEP_PROFILE") == "1" else None

    # -- exchanges ------------------------------------------------------------------------
    def _halo(self, first_frame_ends: torch.Tensor) -> Optional[torch.Tensor]:
        """Send my first frame's ends to rank-1, receive rank+1's.  Returns None on the last rank."""
        if self.world == 1:
            return None
        ops = []
        recv = None
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, first_frame_ends, self.rank - 1, self.group))
        if self.rank < self.world - 1:
            recv = torch.empty_like(first_frame_ends)
            ops.append(dist.P2POp(dist.irecv, recv, self.rank + 1, self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return recv

    def _gather(self, t: torch.Tensor, cache_name: str, frame_dim: int, async_op: bool = False):
        """Gather ``t`` (frames along ``frame_dim``) to rank 0; returns the list of per-rank
        shards there, None elsewhere.  Shards of unequal length are padded for the collective.
        With ``async_op`` returns ``(result, work)``: the collective runs on NCCL's stream while
        the caller keeps enqueueing compute; ``work.wait()`` orders the current stream after it."""
        if self.world == 1:
            return ([t], None) if async_op else [t]
        fmax = max(self.shard_sizes)
        if t.shape[frame_dim] < fmax:
            pad_shape = list(t.shape)
            pad_shape[frame_dim] = fmax - t.shape[frame_dim]
            t = torch.cat([t, t.new_zeros(pad_shape)], dim=frame_dim)
        t = t.contiguous()
        if self.rank == 0:
            buf = getattr(self, cache_name)
            if buf is None or buf[0].shape != t.shape:
                buf = [torch.empty_like(t) for _ in range(self.world)]
                setattr(self, cache_name, buf)
            work = dist.gather(t, buf, dst=0, group=self.group, async_op=async_op)
            res = [b.narrow(frame_dim, 0, n) for b, n in zip(buf, self.shard_sizes)]
            return (res, work) if async_op else res
        work = dist.gather(t, None, dst=0, group=self.group, async_op=async_op)
        return (None, work) if async_op else None

    # -- one pass over this rank's frames ----------------------------------------------------
    def step(self, cam1, cam2, n1, n2):
        """cam1/cam2 [F*C,Nmax,2,2] (problem = f*C+c), n1/n2 [F*C] for this rank's frames.
        Returns a dict; on rank 0 ``rows_all`` / ``track_id_all`` hold every rank's shard.  With
        ``gather_to_root`` the matching outputs (``rows``, ``col_ind``, ``status``) live in two alternating
        sets of buffers: they are overwritten by the step after next -- copy what must outlive that."""
        F, C, N = self.F, self.C, self.N
        prof = self._prof
        if prof is not None:
            import time as _t

            t_last = [_t.perf_counter()]

            def mark(name):
                now = _t.perf_counter()
                prof[name] = prof.get(name, 0.0) + (now - t_last[0])
                t_last[0] = now
        else:
            def mark(name):
                pass
        rows_all = rows_work = None
        slot = self._step_idx % 2
        self._step_idx += 1
        if self.gather_to_root:
            self._wait(slot)  # this slot's send and receive buffers are free again
        if isinstance(self.ops, _CudaOps):
            rows, col_m, status = self.ops.match(cam1, cam2, n1, n2, slot=slot if self.gather_to_root else None)
        else:
            rows, col_m, status = self.ops.match(cam1, cam2, n1, n2)
        mark("match")
        if self.gather_to_root:
            # the big gather (rows) starts as soon as matching is enqueued and overlaps tracking
            rows_all, rows_work = self._gather(rows.view(F, C, -1, rows.shape[-1]), f"_rows_all{slot}", 0, async_op=True)
        mark("wait+gather_rows")
        ends = self.ops.rows_to_ends(rows, F, C)  # [C,F,N,2,3]
        mark("rows_to_ends")
        halo = self._halo(ends[:, 0].contiguous())
        mark("halo")
        if halo is not None:
            ends = torch.cat([ends, halo[:, None]], dim=1)  # boundary pair solved here
        mark("cat")
        col, total, tstatus = self.ops.track(ends)
        mark("track")
        local = self.ops.chain(col)  # ids relative to this rank's first frame
        mark("chain")
        if self.world > 1:
            if halo is not None:
                my_map = local[:, -1].contiguous()
            else:
                my_map = torch.arange(N, dtype=local.dtype, device=local.device).expand(C, N).contiguous()
            maps = torch.empty((self.world * C, N), dtype=local.dtype, device=local.device)
            dist.all_gather_into_tensor(maps, my_map, group=self.group)
            maps = maps.view(self.world, C, N)
            if hasattr(self.ops, "rebase"):
                track_id = self.ops.rebase(maps, self.rank, local)[:, :F].contiguous()
            else:  # back ends without a fused kernel: same composition with torch ops
                start = compose_rank_maps(maps)[self.rank].contiguous()
                track_id = self.ops.chain(col, start)[:, :F].contiguous()
        else:
            track_id = local
        mark("maps+rebase")
        out = dict(rows=rows, col_ind=col_m, status=status, track_col=col, track_total=total,
                   track_status=tstatus, track_id=track_id)
        if self.gather_to_root:
            tid_all, tid_work = self._gather(track_id, f"_tid_all{slot}", 1, async_op=True)
            self._pending[slot] = [w for w in (rows_work, tid_work) if w is not None]
            out["rows_all"] = rows_all        # [F,C,Nmax,18] per rank   } valid after finish()
            out["track_id_all"] = tid_all     # [C,F,N] per rank         } (or the next-but-one step)
        mark("gather_tid")
        return out

    def _wait(self, slot: int):
        if self._pending[slot]:
            for w in self._pending[slot]:
                w.wait()
        self._pending[slot] = None

    def finish(self):
        """Order the current stream after every outstanding gather (call before reading
        ``rows_all`` / ``track_id_all`` of the latest steps)."""
        self._wait(0)
        self._wait(1)
