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


def _worker(rank, world, port, F, C, N, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from particletracking_b200 import rig as rigmod
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(F, C, N, seed=5)
    rig = rigmod.rig_from_dicts(data.calibration, data.transformation)
    f0, f1 = pdist.shard_frames(F, world, rank)
    sl = slice(f0, f1)
    Fl = f1 - f0
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    rec = pdist.ShardedReconstructor(rig, Fl, C, N, "cpu", ops=OracleOps(rig), n_frames_total=F)
    out = rec.step(t(data.cam1[sl].reshape(Fl * C, N, 2, 2)), t(data.cam2[sl].reshape(Fl * C, N, 2, 2)),
                   t(data.n1[sl].reshape(-1)), t(data.n2[sl].reshape(-1)))
    rec.finish()
    if rank == 0:
        rows = torch.cat(list(out["rows_all"]), dim=0).numpy()
        tid = torch.cat(list(out["track_id_all"]), dim=1).numpy()
        np.savez(out_path, rows=rows, tid=tid)
    else:
        assert out["rows_all"] is None
    dist.destroy_process_group()


@pytest.mark.parametrize("world,F", [(2, 11), (3, 10)])
def test_sharded_equals_single_rank(tmp_path, world, F):
    C, N = 2, 6
    out_path = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(world, _free_port(), F, C, N, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    # single-rank oracle over the whole range
    from oracle import c_oracle as co
    from particletracking_b200 import rig as rigmod
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(F, C, N, seed=5)
    rig = rigmod.rig_from_dicts(data.calibration, data.transformation)
    c1, c2, n1, n2 = data.problems()
    o = co.match_batch(rig, c1, c2, n1, n2)
    rows = o["rows"].reshape(F, C, N, 18)
    np.testing.assert_array_equal(got["rows"], rows)
    ends = rows[..., :6].reshape(F, C, N, 2, 3).transpose(1, 0, 2, 3, 4).copy()
    tid = co.chain_tracks(co.track_batch(ends)["col_ind"])
    np.testing.assert_array_equal(got["tid"], tid)
