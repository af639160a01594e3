"""Frame sharding across the GPUs of one box.

One process per GPU (``torch.distributed``, backend ``nccl``; ``gloo`` in the CPU tests).
The path shards naturally, and a step of one rank is ONE library call plus ONE collective:

* matching: every (frame, colour) problem is independent -> contiguous frame ranges per rank;
* tracking: the pair (f, f+1) needs only those two frames' rows -> same partition plus a 1-frame
  halo.  The halo is RE-COMPUTED, not exchanged: rank r is handed the detections of its own F
  frames plus the first frame of rank r+1 and matches that frame again (the matching is a pure
  function of its inputs: identical bits on both ranks), so the boundary pair is tracked locally
  and there is no send/recv on the critical path (``ptk_reconstruct_shard``);
* chain pass: permutation composition is associative -> every rank chains locally from the
  identity and emits, besides its local ids, its boundary map (the local id of every position of
  the halo frame);
* final collective: every rank's result block -- rows, costs, local track ids, boundary map, status
  words, one contiguous buffer -- is gathered to rank 0 by ONE NCCL gather per step, issued on a
  side stream so that it overlaps the next step's kernels; rank 0 then composes the boundary maps
  and re-bases the gathered ids in place (``ptk_chain_rebase_gathered``), also on the side stream.

There is no data-path collective inside matching or the per-pair assignment; NCCL carries only the
final gather (the north star's "NCCL only for the final gather of 3D rods and track IDs").  Nothing
in step k+1 depends on the gather of step k except the reuse of its buffers two steps later, which
is ordered by CUDA events, never by the host.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

ROW_COLS = 18


def shard_frames(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous frame range [f0, f1) of ``rank``; sizes differ by at most one."""
    base, rem = divmod(n_frames, world)
    f0 = rank * base + min(rank, rem)
    return f0, f0 + base + (1 if rank < rem else 0)


def shard_input_range(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """Frames whose detections ``rank`` must be given: its own range plus the 1-frame halo."""
    f0, f1 = shard_frames(n_frames, world, rank)
    return f0, min(f1 + 1, n_frames) if rank < world - 1 else f1


def _pick(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """table[c, idx[c, x]] with -1 for idx outside 0..N-1 (a broken chain stays broken)."""
    N = table.shape[-1]
    ok = (idx >= 0) & (idx < N)
    out = torch.gather(table, -1, idx.clamp(0, N - 1).long())
    return torch.where(ok, out, torch.full_like(out, -1))


def compose_rank_maps(maps: torch.Tensor) -> torch.Tensor:
    """maps[W,C,N]: for rank r, position x in rank r+1's first frame -> local id (position in
    rank r's first frame).  Returns starts[W,C,N]: global track id of local id x on rank r:
    starts[0] = arange(N), starts[r+1][c][x] = starts[r][c][maps[r][c][x]] (-1 where the chain broke)."""
    W, C, N = maps.shape
    starts = torch.empty_like(maps)
    cur = torch.arange(N, dtype=maps.dtype, device=maps.device).expand(C, N).contiguous()
    for r in range(W):
        starts[r] = cur
        if r + 1 < W:
            cur = _pick(cur, maps[r])
    return starts


class _BlockLayout:
    """Byte layout of one rank's result block (every region 256-byte aligned):
    ``track_id | halo_map | status | n_out | track_status | track_total | match_cost | rows``.
    ``rows`` is last and has room for the halo frame; only ``send_bytes`` (everything up to the rows of
    ``Fmax`` own frames) travel."""

    def __init__(self, Fmax: int, C: int, N: int):
        self.Fmax, self.C, self.N = Fmax, C, N
        self.off: Dict[str, int] = {}
        o = 0
        for name, nbytes in (("track_id", C * Fmax * N * 4), ("halo_map", C * N * 4), ("status", (Fmax + 1) * C * 4),
                             ("n_out", (Fmax + 1) * C * 4), ("track_status", C * Fmax * 4), ("track_total", C * Fmax * 8),
                             ("match_cost", (Fmax + 1) * C * N * 8), ("rows", (Fmax + 1) * C * N * ROW_COLS * 8)):
            self.off[name] = o
            o += (nbytes + 255) // 256 * 256
        self.capacity = o
        self.send_bytes = self.off["rows"] + Fmax * C * N * ROW_COLS * 8

    def views(self, block: torch.Tensor, F: int, h: int, own_only: bool = False) -> Dict[str, torch.Tensor]:
        """Typed views of ``block`` (uint8) for a rank with F own frames and h halo frames.  ``own_only``:
        the per-problem arrays cover the F own frames only (a received block is ``send_bytes`` long: the
        halo frame's rows did not travel); the per-pair arrays keep their [C, F+h-1] shape."""
        C, N = self.C, self.N
        f64, i32 = torch.float64, torch.int32

        def v(name, dtype, shape):
            n = 1
            for d in shape:
                n *= d
            nb = n * (8 if dtype == f64 else 4)
            return block[self.off[name]: self.off[name] + nb].view(dtype).view(shape)

        Fp = F if own_only else F + h
        return dict(track_id=v("track_id", i32, (C, F, N)), halo_map=v("halo_map", i32, (C, N)),
                    status=v("status", i32, (Fp * C,)), n_out=v("n_out", i32, (Fp * C,)),
                    track_status=v("track_status", i32, (C, F + h - 1)), track_total=v("track_total", f64, (C, F + h - 1)),
                    match_cost=v("match_cost", f64, (Fp * C, N)), rows=v("rows", f64, (Fp * C, N, ROW_COLS)))


class ShardedReconstructor:
    """match -> track -> chain for one rank's frame range; see the module docstring.

    ``step`` takes the detections of ``F + n_halo`` frames (``n_halo`` = 1 on every rank but the last:
    ``shard_input_range``) and returns views of the rank's result block; on rank 0 ``rows_all`` /
    ``track_id_all`` (global ids) / ``match_cost_all`` / ``status_all`` hold every rank's shard once the
    gather has completed (``finish()`` orders the current stream after it).  Results live in two
    alternating sets of buffers: they are overwritten by the step after next.

    ``ops`` is the compute back end (default: the CUDA kernels through ``engine.ShardPipeline``).  The
    CPU tests inject an oracle-backed one to exercise the exchange logic under ``gloo`` -- test
    infrastructure, not a fallback: the default never touches it.
    """

    def __init__(self, rig, n_frames_local: int, n_colors: int, n_rods: int, device, ops=None,
                 gather_to_root: bool = True, group=None, n_frames_total: Optional[int] = None,
                 max_repr_err: Optional[float] = None, use_graph: bool = True):
        self.F, self.C, self.N = n_frames_local, n_colors, n_rods
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.gather_to_root = gather_to_root and self.world > 1
        self.use_graph = use_graph
        total = n_frames_total if n_frames_total is not None else n_frames_local * self.world
        self.shard_sizes = [b - a for a, b in (shard_frames(total, self.world, r) for r in range(self.world))]
        if self.shard_sizes[self.rank] != n_frames_local:
            raise ValueError("n_frames_local does not match shard_frames(n_frames_total, world, rank)")
        self.h = 1 if self.rank < self.world - 1 else 0
        self.layout = _BlockLayout(max(self.shard_sizes), n_colors, n_rods)
        self.cuda = self.device.type == "cuda"
        self.ops = ops
        if ops is None and not self.cuda:
            raise RuntimeError("ShardedReconstructor: the product path needs a CUDA device (there is no CPU fallback)")
        u8 = torch.uint8
        self.blocks = [torch.zeros(self.layout.capacity, dtype=u8, device=self.device) for _ in range(2)]
        self.views = [self.layout.views(b, self.F, self.h) for b in self.blocks]
        self.recv = [None, None]
        if self.gather_to_root and self.rank == 0:
            self.recv = [torch.zeros((self.world, self.layout.send_bytes), dtype=u8, device=self.device) for _ in range(2)]
        self._step_idx = 0
        F, C = self.F, self.C
        self._out_base = [dict(rows=v["rows"][: F * C].view(F, C, self.N, ROW_COLS), match_cost=v["match_cost"][: F * C],
                               status=v["status"][: F * C], n_out=v["n_out"][: F * C], track_total=v["track_total"],
                               track_status=v["track_status"], track_id=v["track_id"], halo_map=v["halo_map"] if self.h else None)
                          for v in self.views]
        # typed views of the gathered blocks, built once (a view costs microseconds of Python; 8 ranks x 8 arrays per step did show)
        self._recv_views = [None, None]
        if self.gather_to_root and self.rank == 0:
            W = self.world
            for s in range(2):
                per = [self.layout.views(self.recv[s][r], self.shard_sizes[r], 1 if r < W - 1 else 0, own_only=True) for r in range(W)]
                Fs = self.shard_sizes
                self._recv_views[s] = dict(
                    per=per, lists=[self.recv[s][r] for r in range(W)],
                    rows_all=[per[r]["rows"].view(Fs[r], self.C, self.N, ROW_COLS) for r in range(W)],
                    track_id_all=[per[r]["track_id"] for r in range(W)], match_cost_all=[per[r]["match_cost"] for r in range(W)],
                    status_all=[per[r]["status"] for r in range(W)], track_total_all=[per[r]["track_total"] for r in range(W)],
                    track_status_all=[per[r]["track_status"] for r in range(W)])
        if self.cuda:
            from . import engine

            self._engine = engine
            ws = None
            self.pipes = []
            for s in range(2):
                if ops is None:
                    p = engine.ShardPipeline(rig, self.F, self.h, self.C, self.N, self.device, out=self.views[s],
                                             max_repr_err=max_repr_err, workspace=ws)
                    ws = p.ws  # both slots run on the same stream, in order: one workspace
                    self.pipes.append(p)
            self.side = torch.cuda.Stream(self.device)
            self.ev_comp = [torch.cuda.Event() for _ in range(2)]
            self.ev_done = [torch.cuda.Event() for _ in range(2)]
            self._done_recorded = [False, False]
            self._scratch = torch.empty((max(self.world, 1),), dtype=torch.int32, device=self.device)
            self._maps_all = torch.empty((self.world, self.C, self.N), dtype=torch.int32, device=self.device)

    # -- compute ----------------------------------------------------------------------------------
    def _compute_with_ops(self, v, cam1, cam2, n1, n2):
        """Generic back end (tests): same outputs as ptk_reconstruct_shard, written into the block views."""
        F, h, C = self.F, self.h, self.C
        rows, col_m, status = self.ops.match(cam1, cam2, n1, n2)
        v["rows"].copy_(rows.view(v["rows"].shape))
        v["status"].copy_(status.view(-1))
        ends = self.ops.rows_to_ends(v["rows"], F + h, C)  # [C,F+h,N,2,3]
        if F + h > 1:
            col, total, tstatus = self.ops.track(ends)
            v["track_total"].copy_(total)
            v["track_status"].copy_(tstatus)
        else:
            col = torch.empty((C, 0, self.N), dtype=torch.int32)
        loc = self.ops.chain(col)  # ids relative to this rank's first frame, [C,F+h,N]
        v["track_id"].copy_(loc[:, :F])
        if h:
            v["halo_map"].copy_(loc[:, F])
        return col

    # -- one pass over this rank's frames -----------------------------------------------------------
    def step(self, cam1, cam2, n1, n2):
        """cam1/cam2 [(F+n_halo)*C,Nmax,2,2] (problem = f*C+c), n1/n2 [(F+n_halo)*C]: this rank's frames
        followed by the next rank's first frame (not on the last rank)."""
        slot = self._step_idx % 2
        self._step_idx += 1
        v = self.views[slot]
        out = dict(self._out_base[slot])
        if self.cuda:
            main = torch.cuda.current_stream(self.device)
            if self._done_recorded[slot]:
                main.wait_event(self.ev_done[slot])  # the slot's block has been sent (two steps ago)
            if self.ops is None:
                p = self.pipes[slot]
                (p.run_graph if self.use_graph else p.run)(cam1, cam2, n1, n2)
                out["track_col"], out["col_ind"] = p.track_col, p.col_ind
            else:
                out["track_col"] = self._compute_with_ops(v, cam1, cam2, n1, n2)
            if self.world > 1:
                self.ev_comp[slot].record(main)
                with torch.cuda.stream(self.side):
                    self.side.wait_event(self.ev_comp[slot])
                    self._exchange(slot, out)
                    self.ev_done[slot].record(self.side)
                self._done_recorded[slot] = True
        else:
            out["track_col"] = self._compute_with_ops(v, cam1, cam2, n1, n2)
            if self.world > 1:
                self._exchange(slot, out)
        if self.world == 1:
            out["track_id_global"] = out["track_id"]
        return out

    def _exchange(self, slot: int, out: dict):
        """The step's only collective (+ the re-base).  Runs on the side stream on CUDA."""
        L, W = self.layout, self.world
        send = self.blocks[slot][: L.send_bytes]
        if self.gather_to_root:
            if self.rank == 0:
                recv, rv = self.recv[slot], self._recv_views[slot]
                dist.gather(send, rv["lists"], dst=0, group=self.group)
                per = rv["per"]
                if self.cuda and self.ops is None:
                    self._engine.chain_rebase_gathered(recv, self.shard_sizes, self.C, self.N, L.off["track_id"],
                                                       L.off["halo_map"], self._scratch)
                else:
                    maps = torch.stack([p["halo_map"] for p in per])
                    starts = compose_rank_maps(maps)
                    for r in range(1, W):
                        per[r]["track_id"].copy_(_pick(starts[r][:, None, :].expand_as(per[r]["track_id"]), per[r]["track_id"]))
                for k in ("rows_all", "track_id_all", "match_cost_all", "status_all", "track_total_all", "track_status_all"):
                    out[k] = rv[k]
                out["track_id_global"] = per[0]["track_id"]
            else:
                dist.gather(send, None, dst=0, group=self.group)
                out["rows_all"] = out["track_id_all"] = None
        else:
            # every rank wants its own global ids: all-gather the boundary maps (a few KB), re-base locally
            v = self.views[slot]
            my_map = v["halo_map"] if self.h else torch.arange(self.N, dtype=torch.int32, device=self.device).expand(
                self.C, self.N).contiguous()
            maps = self._maps_all if self.cuda else torch.empty((W, self.C, self.N), dtype=torch.int32)
            dist.all_gather_into_tensor(maps.view(W * self.C, self.N), my_map.contiguous(), group=self.group)
            if self.cuda and self.ops is None:
                gid = v["track_id"].clone()
                self._engine.chain_rebase(maps, self.rank, gid)
            else:
                start = compose_rank_maps(maps)[self.rank]
                gid = _pick(start[:, None, :].expand_as(v["track_id"]), v["track_id"])
            out["track_id_global"] = gid

    def finish(self):
        """Order the current stream after every outstanding gather / re-base (call before reading
        ``rows_all`` / ``track_id_all`` / ``track_id_global`` of the latest steps)."""
        if self.cuda and self.world > 1:
            main = torch.cuda.current_stream(self.device)
            for s in range(2):
                if self._done_recorded[s]:
                    main.wait_event(self.ev_done[s])
