"""Worker for tests/test_dist_gpu.py::test_c5_full_size: BASELINE config C5 (large dense LP, m = 262144, n = 65536,
A row-sharded across the ranks, f32, generated in HBM per shard: 68.7 GB total).  The oracle cannot hold this matrix,
so the checks are size-independent properties plus sampled rows/columns against the numpy twin of the generator:
(A x)[rows], (A^T u)[cols], the adjoint identity, bit-identical replicas, and 20 solver iterations whose residuals
must be finite, identical on every rank and identical between the peer-store and the paired/unpaired orders."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import helpers as H  # noqa: E402
from totsu_b200 import capi, host, shard, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    m, n = int(os.environ.get("C5_M", 262144)), int(os.environ.get("C5_N", 65536))
    torch.cuda.set_device(local)
    capi.init(local)
    L = capi.lib()
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idbuf = (C.c_ubyte * capi.NCCL_ID_BYTES)()
    if rank == 0:
        capi.check(L.tb_dist_unique_id(idbuf))
    t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    idbuf = (C.c_ubyte * capi.NCCL_ID_BYTES)(*t.cpu().tolist())
    capi.check(L.tb_dist_init(rank, world, idbuf))
    dt = np.float32
    blocks = [(H.RPOS, m)]
    row_off, m_loc = shard.row_shards(blocks, world)[rank]
    scale = dt(1.0 / np.sqrt(n))
    abuf = capi.Buf(dtype=dt, length=m_loc * n)
    capi.check(L.tb_fill_uniform_f32(abuf.view(), m_loc, n, row_off, 5, scale))
    hop = C.c_int64()
    capi.check(L.tb_denseop_create(capi.TB_F32, abuf.view(), m_loc, n, row_off, m, C.byref(hop)))
    rng = np.random.default_rng(5)
    x = rng.standard_normal(n).astype(dt); u = rng.standard_normal(m).astype(dt)
    xb, ub = capi.Buf(x.copy(), mutable=False), capi.Buf(u.copy(), mutable=False)
    ax, atu = np.zeros(m, dt), np.zeros(n, dt)
    b1, b2 = capi.Buf(ax), capi.Buf(atu)
    capi.check(L.tb_denseop_apply_f32(hop.value, 0, 1.0, xb.view(), 0.0, b1.view()))
    capi.check(L.tb_denseop_apply_f32(hop.value, 1, 1.0, ub.view(), 0.0, b2.view()))
    for b in (b1, b2, xb, ub):
        b.release()
    lhs, rhs = np.dot(ax.astype(np.float64), u.astype(np.float64)), np.dot(x.astype(np.float64), atu.astype(np.float64))
    assert abs(lhs - rhs) <= 1e-4 * (np.linalg.norm(ax) * np.linalg.norm(u)), ("adjoint", lhs, rhs)
    rows = np.array([0, 1, m // world - 1, m // world, m // 2 + 17, m - 1])
    sub = synth.uniform_matrix(len(rows), n, 5, scale, dtype=dt, rows=rows).astype(np.float64)
    assert H.rel_linf(ax[rows], sub @ x.astype(np.float64)) <= 5e-5
    cols = np.array([0, 7, 8, n // 2 - 1, n - 1])
    subc = synth.uniform_matrix(m, len(cols), 5, scale, dtype=dt, cols=cols).astype(np.float64)
    assert H.rel_linf(atu[cols], subc.T @ u.astype(np.float64)) <= 5e-5
    mine = torch.from_numpy(np.concatenate([ax, atu]).astype(np.float64)).cuda()
    ref = mine.clone(); dist.broadcast(ref, 0)
    assert torch.equal(mine, ref), "replicas diverged"
    capi.check(L.tb_denseop_destroy(hop.value))

    # 20 iterations of the LP (cone = RPos(m)), pairing on vs off: identical iterates, finite residuals
    x0 = rng.standard_normal(n) / np.sqrt(m)
    s0 = (np.abs(rng.standard_normal(m)) + 0.1) / np.sqrt(m)
    bvec = s0.astype(dt); cvec = (rng.standard_normal(n) / np.sqrt(n)).astype(dt)
    res = {}
    for fuse in (1, 0):
        capi.check(L.tb_set_pair_fusion(fuse))
        s = host.Session.dense(dt, abuf.view(), m_loc, n, cvec, bvec, blocks, fused_op=True, fused_cone=True, row_offset=row_off, m_total=m)
        assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0, device_precond=True) == "None"
        s.step(20)
        xh, yh = s.xy()
        res[fuse] = (xh, yh, (s.last.c0, s.last.c1, s.last.c2))
        s.close()
    capi.check(L.tb_set_pair_fusion(1))
    assert np.isfinite(res[1][0]).all() and np.isfinite(res[1][1]).all()
    assert np.array_equal(res[1][0], res[0][0]) and np.array_equal(res[1][1], res[0][1])
    mine = torch.from_numpy(np.concatenate([res[1][0], res[1][1]]).astype(np.float64)).cuda()
    ref = mine.clone(); dist.broadcast(ref, 0)
    assert torch.equal(mine, ref), "solver replicas diverged"
    abuf.release()
    capi.check(L.tb_device_sync())
    dist.barrier()
    capi.check(L.tb_dist_finalize())
    dist.destroy_process_group()
    if rank == 0:
        print("dist_c5_worker ok: world=%d m=%d n=%d residuals=%s" % (world, m, n, res[1][2]))


if __name__ == "__main__":
    main()
