"""World-size-2 `gloo` test (CPU) of the multi-GPU scheme: row shards on cone-block boundaries, A*x slices
all-gathered, A^T*y partials all-reduced (north_star / SURVEY.md §8e).  Each rank regenerates only its own rows
from the counter-based generator and runs the ORACLE solver with a sharded operator; the iterates must equal the
single-process oracle's.  This covers the host-side N>1 logic; the NCCL path itself runs under `-m gpu`/bench."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from totsu_b200 import shard, synth


def test_row_shards_align_with_cone_blocks():
    blocks = [(H.SOC, 64)] * 1024
    sh = shard.row_shards(blocks, 8)
    assert sh == [(g * 8192, 8192) for g in range(8)]
    assert shard.row_shards([(H.RPOS, 10), (H.ZERO, 6)], 4) == [(0, 4), (4, 4), (8, 4), (12, 4)]
    with pytest.raises(ValueError):
        shard.row_shards([(H.SOC, 6), (H.SOC, 6)], 4)        # would cut a second-order block
    with pytest.raises(ValueError):
        shard.row_shards([(H.SOC, 5), (H.SOC, 6)], 2)


class _ShardedOp:
    """Oracle-side twin of the CUDA DenseOp's sharded mode: local rows only + gloo collectives."""
    def __init__(self, a_loc, row_off, m_total):
        self.a, self.off, self.m = a_loc, row_off, m_total

    def size(self):
        return (self.m, self.a.shape[1])

    def op(self, alpha, x, beta, y):
        loc = alpha * (self.a @ x) + (beta * y[self.off:self.off + self.a.shape[0]] if beta != 0.0 else 0.0)
        parts = [torch.zeros(self.a.shape[0], dtype=torch.float64) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(loc)))
        y[...] = torch.cat(parts).numpy()

    def trans_op(self, alpha, x, beta, y):
        part = torch.from_numpy(self.a.T @ x[self.off:self.off + self.a.shape[0]])
        dist.all_reduce(part)
        y[...] = alpha * part.numpy() + (beta * y if beta != 0.0 else 0.0)

    def absadd_cols(self, tau):
        part = torch.from_numpy(np.abs(self.a).sum(0))
        dist.all_reduce(part)
        tau += part.numpy()

    def absadd_rows(self, sigma):
        loc = torch.from_numpy(np.abs(self.a).sum(1))
        parts = [torch.zeros_like(loc) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, loc)
        sigma += torch.cat(parts).numpy()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blocks = [(H.SOC, 8)] * 6 + [(H.RPOS, 16)]
    m, n = 64, 24
    a, b, c = H.make_instance(m, n, blocks, seed=3, dtype=np.float32)
    off, rows = shard.row_shards(blocks, world)[rank]
    a_loc = synth.uniform_matrix(rows, n, 3, np.float32(1.0 / np.sqrt(n)), row_offset=off, dtype=np.float32).astype(np.float64)
    assert np.array_equal(a_loc, a[off:off + rows].astype(np.float64))          # a shard regenerates its rows exactly
    O = H.O
    s = O.Solver()
    s.par.max_iter = 32; s.par.eps_acc = 0.0; s.par.eps_inf = 0.0
    s.snapshots = {30: None}
    prob = (O.MatOp(O.MatType.General(n, 1), c.astype(np.float64)), _ShardedOp(a_loc, off, m),
            O.MatOp(O.MatType.General(m, 1), b.astype(np.float64)), H.oracle_cone(blocks), np.zeros(O.Solver.query_worklen((m, n))))
    try:
        s.solve(prob)
    except O.SolverError:
        pass
    snaps, _ = H.oracle_iterates(a, b, c, blocks, [30])
    ex = H.rel_linf(s.snapshots[30][0], snaps[30][0])
    ey = H.rel_linf(s.snapshots[30][1], snaps[30][1])
    q.put((rank, ex, ey))
    dist.destroy_process_group()


def test_sharded_iterates_match_single_process_gloo():
    world = 2
    sock = socket.socket(); sock.bind(("127.0.0.1", 0)); port = sock.getsockname()[1]; sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ex, ey in res:
        assert ex <= 1e-10 and ey <= 1e-10, (rank, ex, ey)
