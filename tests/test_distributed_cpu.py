"""Multi-rank host logic on CPU: world_size-2 `gloo` run of ShardedReconstructor with an
oracle-backed compute back end, against a single-rank run over the whole frame range."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from particletracking_b200 import distributed as pdist


def test_shard_frames_partition():
    for F in (1, 7, 1000, 50000):
        for W in (1, 2, 3, 8):
            spans = [pdist.shard_frames(F, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == F
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_compose_rank_maps_equals_sequential_chain():
    rng = np.random.default_rng(0)
    W, C, N = 5, 3, 9
    maps = np.stack([[rng.permutation(N) for _ in range(C)] for _ in range(W)]).astype(np.int32)
    starts = pdist.compose_rank_maps(torch.from_numpy(maps)).numpy()
    cur = np.tile(np.arange(N), (C, 1))
    for r in range(W):
        np.testing.assert_array_equal(starts[r], cur)
        cur = np.take_along_axis(cur, maps[r].astype(np.int64), axis=1)


class OracleOps:
    """Test-only compute back end: the C oracle on CPU tensors."""

    def __init__(self, rig):
        from oracle import c_oracle as co

        self.co, self.rig = co, rig

    def match(self, cam1, cam2, n1, n2):
        o = self.co.match_batch(self.rig, cam1.numpy(), cam2.numpy(), n1.numpy(), n2.numpy())
        return torch.from_numpy(o["rows"]), torch.from_numpy(o["col_ind"]), torch.from_numpy(o["status"])

    def rows_to_ends(self, rows, F, C):
        N = rows.shape[-2]
        return rows[..., :6].reshape(F, C, N, 2, 3).permute(1, 0, 2, 3, 4).contiguous()

    def track(self, ends):
        o = self.co.track_batch(ends.numpy())
        return torch.from_numpy(o["col_ind"]), torch.from_numpy(o["total_cost"]), torch.from_numpy(o["status"])

    def chain(self, col, ids_in=None):
        return torch.from_numpy(self.co.chain_tracks(col.numpy(), None if ids_in is None else ids_in.numpy()))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, F, C, N, out_path, to_root):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from particletracking_b200 import rig as rigmod
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(F, C, N, seed=5)
    rig = rigmod.rig_from_dicts(data.calibration, data.transformation)
    f0, f1 = pdist.shard_frames(F, world, rank)
    i0, i1 = pdist.shard_input_range(F, world, rank)  # own frames + the re-computed halo frame
    assert (i0, i1) == (f0, f1 + (1 if rank < world - 1 else 0))
    sl = slice(i0, i1)
    Fin = i1 - i0
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    rec = pdist.ShardedReconstructor(rig, f1 - f0, C, N, "cpu", ops=OracleOps(rig), n_frames_total=F, gather_to_root=to_root)
    args = (t(data.cam1[sl].reshape(Fin * C, N, 2, 2)), t(data.cam2[sl].reshape(Fin * C, N, 2, 2)),
            t(data.n1[sl].reshape(-1)), t(data.n2[sl].reshape(-1)))
    for _ in range(3):  # both buffer slots, and a slot's reuse
        out = rec.step(*args)
    rec.finish()
    if to_root:
        if rank == 0:
            rows = torch.cat(list(out["rows_all"]), dim=0).numpy()
            tid = torch.cat(list(out["track_id_all"]), dim=1).numpy()
            tot = torch.cat(list(out["track_total_all"]), dim=1).numpy()
            assert all(int((s != 0).sum()) == 0 for s in out["status_all"])
            np.savez(out_path, rows=rows, tid=tid, tot=tot)
        else:
            assert out["rows_all"] is None
    else:
        np.savez(out_path + f".{rank}.npz", tid=out["track_id_global"].numpy(), f0=f0, f1=f1)
    dist.destroy_process_group()


def _single_rank_oracle(F, C, N):
    from oracle import c_oracle as co
    from particletracking_b200 import rig as rigmod
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(F, C, N, seed=5)
    rig = rigmod.rig_from_dicts(data.calibration, data.transformation)
    c1, c2, n1, n2 = data.problems()
    o = co.match_batch(rig, c1, c2, n1, n2)
    rows = o["rows"].reshape(F, C, N, 18)
    ends = rows[..., :6].reshape(F, C, N, 2, 3).transpose(1, 0, 2, 3, 4).copy()
    tr = co.track_batch(ends)
    return rows, co.chain_tracks(tr["col_ind"]), tr["total_cost"]


@pytest.mark.parametrize("world,F", [(2, 11), (3, 10)])
def test_sharded_equals_single_rank(tmp_path, world, F):
    """One gather of the result blocks + root-side re-base == a single-rank run over the whole range
    (halo frame re-computed on the rank before it, no send/recv)."""
    C, N = 2, 6
    out_path = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(world, _free_port(), F, C, N, out_path, True), nprocs=world, join=True)
    got = np.load(out_path)
    rows, tid, tot = _single_rank_oracle(F, C, N)
    np.testing.assert_array_equal(got["rows"], rows)
    np.testing.assert_array_equal(got["tid"], tid)
    np.testing.assert_array_equal(got["tot"], tot)  # every pair tracked exactly once, boundary pairs included


def test_sharded_global_ids_on_every_rank(tmp_path):
    """gather_to_root=False: the boundary maps are all-gathered and every rank re-bases its own ids."""
    world, F, C, N = 3, 10, 2, 6
    out_path = str(tmp_path / "out")
    mp.spawn(_worker, args=(world, _free_port(), F, C, N, out_path, False), nprocs=world, join=True)
    _, tid, _ = _single_rank_oracle(F, C, N)
    for r in range(world):
        g = np.load(out_path + f".{r}.npz")
        np.testing.assert_array_equal(g["tid"], tid[:, int(g["f0"]):int(g["f1"])])


def test_block_layout_and_broken_maps():
    L = pdist._BlockLayout(7, 3, 5)
    assert all(o % 256 == 0 for o in L.off.values()) and L.send_bytes <= L.capacity
    blk = torch.zeros(L.capacity, dtype=torch.uint8)
    v = L.views(blk, 6, 1)
    assert v["rows"].shape == (7 * 3, 5, 18) and v["track_id"].shape == (3, 6, 5) and v["track_total"].shape == (3, 6)
    v["rows"].fill_(1.5)
    assert float(L.views(blk, 7, 0)["rows"].sum()) == 1.5 * 7 * 3 * 5 * 18
    w = L.views(blk[: L.send_bytes], 6, 1, own_only=True)
    assert w["rows"].shape == (6 * 3, 5, 18) and w["track_total"].shape == (3, 6)
    maps = torch.tensor([[[1, 0, -1]], [[2, 1, 0]], [[0, 1, 2]]], dtype=torch.int32)  # W=3, C=1, N=3
    st = pdist.compose_rank_maps(maps)
    assert st[1].tolist() == [[1, 0, -1]] and st[2].tolist() == [[-1, 0, 1]]
