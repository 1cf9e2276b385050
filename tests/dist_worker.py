"""Worker for tests/test_dist_gpu.py: one process per GPU (torchrun), A row-sharded across the ranks.  Checks the
sharded Operator (A*x all-gather, A^T*y all-reduce, the paired pass, absadd_*) against numpy on the full matrix and
the sharded solver iterates against the CPU oracle, through whichever collective path TB_P2P selects."""
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
    p2p = C.c_int()
    capi.check(L.tb_dist_p2p_enabled(C.byref(p2p)))
    want_p2p = os.environ.get("TB_P2P", "1") != "0"
    assert bool(p2p.value) == want_p2p, "peer exchange state %d, expected %s" % (p2p.value, want_p2p)

    rng = np.random.default_rng(7)      # same stream on every rank: replicated vectors
    for dt in (np.float32, np.float64):
        tol = 2e-5 if dt == np.float32 else 1e-12
        # (m, n): the first shards are large enough for the TMA streaming kernel, the last goes down the generic path
        for m, n in ((4096 * world, 1024), (1024 * world, 2048), (64 * world, 48)):
            blocks = [(H.SOC, 64)] * (m // 64)
            row_off, m_loc = shard.row_shards(blocks, world)[rank]
            scale = dt(1.0 / np.sqrt(n))
            a_full = synth.uniform_matrix(m, n, 3, scale, dtype=dt).astype(np.float64)
            abuf = capi.Buf(dtype=dt, length=m_loc * n)
            capi.check(capi.fn("tb_fill_uniform", dt)(abuf.view(), m_loc, n, row_off, 3, scale))
            hop = C.c_int64()
            capi.check(L.tb_denseop_create(capi.dtype_id(dt), abuf.view(), m_loc, n, row_off, m, C.byref(hop)))
            x = rng.standard_normal(n).astype(dt); y0 = rng.standard_normal(m).astype(dt)
            u = rng.standard_normal(m).astype(dt); v0 = rng.standard_normal(n).astype(dt)
            bx, by, bu, bv = capi.Buf(x.copy()), capi.Buf(y0.copy()), capi.Buf(u.copy()), capi.Buf(v0.copy())
            for rep in range(3):          # repeated: exercises both stage parities and flag reuse
                capi.check(capi.fn("tb_denseop_apply", dt)(hop.value, 0, 1.5, bx.view(), -0.5, by.view()))
                capi.check(capi.fn("tb_denseop_apply", dt)(hop.value, 1, -2.0, bu.view(), 0.25, bv.view()))
            yw, vw = y0.astype(np.float64), v0.astype(np.float64)
            for rep in range(3):
                yw = 1.5 * (a_full @ x.astype(np.float64)) - 0.5 * yw
                vw = -2.0 * (a_full.T @ u.astype(np.float64)) + 0.25 * vw
            ey, ev = H.rel_linf(by.download(), yw), H.rel_linf(bv.download(), vw)
            assert ey < tol and ev < tol, ("apply", dt, m, n, ey, ev)
            # paired pass, beta = 0 on both outputs
            capi.check(capi.fn("tb_denseop_apply_pair", dt)(hop.value, 1.0, bx.view(), 0.0, by.view(), 1.0, bu.view(), 0.0, bv.view()))
            ey = H.rel_linf(by.download(), a_full @ x.astype(np.float64)); ev = H.rel_linf(bv.download(), a_full.T @ u.astype(np.float64))
            assert ey < tol and ev < tol, ("pair", dt, m, n, ey, ev)
            # replicas must agree bit for bit (fixed summation order on every rank)
            mine = torch.from_numpy(np.concatenate([by.download(), bv.download()]).astype(np.float64)).cuda()
            ref = mine.clone(); dist.broadcast(ref, 0)
            assert torch.equal(mine, ref), ("replicas diverged", dt, m, n)
            # absadd
            tau = np.ones(n, dtype=dt); sig = np.ones(m, dtype=dt)
            bt, bs = capi.Buf(tau), capi.Buf(sig)
            capi.check(capi.fn("tb_denseop_absadd_cols", dt)(hop.value, bt.view()))
            capi.check(capi.fn("tb_denseop_absadd_rows", dt)(hop.value, bs.view()))
            et = H.rel_linf(bt.download(), 1 + np.abs(a_full).sum(0)); es = H.rel_linf(bs.download(), 1 + np.abs(a_full).sum(1))
            assert et < tol and es < tol, ("absadd", dt, m, n, et, es)
            for bf in (bx, by, bu, bv, bt, bs):
                bf.release()
            capi.check(L.tb_denseop_destroy(hop.value))
            abuf.release()

        # sharded solver iterates vs the oracle (fused op + fused cone), incl. a shard that takes the streaming kernel
        for (blocks, n) in (([(H.SOC, 16)] * (8 * world), 40), ([(H.SOC, 64)] * (16 * world) + [(H.RPOS, 1024 * world)], 1024)):
            m = sum(l for _, l in blocks)
            a, b, c = H.make_instance(m, n, blocks, seed=11, dtype=dt)
            row_off, m_loc = shard.row_shards(blocks, world)[rank]
            abuf, av = H.device_matrix(np.asfortranarray(a[row_off:row_off + m_loc, :]))
            s = host.Session.dense(dt, av, m_loc, n, c, b, blocks, fused_op=True, fused_cone=True, row_offset=row_off, m_total=m)
            assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0, device_precond=True) == "None"
            ks = [1, 10, 50]
            snaps, _ = H.oracle_iterates(a, b, c, blocks, ks)
            done = 0
            for k in ks:
                s.step(k - done); done = k
                xh, yh = s.xy()
                tk = {np.float32: {1: 5e-6, 10: 5e-5, 50: 1e-4}, np.float64: {1: 1e-12, 10: 1e-11, 50: 1e-9}}[dt][k]
                ex, ey = H.rel_linf(xh, snaps[k][0]), H.rel_linf(yh, snaps[k][1])
                assert ex <= tk and ey <= tk, ("iterates", dt, m, n, k, ex, ey)
            # replicas stay bit-identical through the solver loop (fixed summation order, same decisions on every rank)
            mine = torch.from_numpy(np.concatenate([xh, yh]).astype(np.float64)).cuda()
            ref = mine.clone(); dist.broadcast(ref, 0)
            assert torch.equal(mine, ref), ("solver replicas diverged", dt, m, n)
            if want_p2p and m_loc * n >= (1 << 20):
                # steady state on a streaming-size shard: 2 peer exchanges per iteration (the speculated criteria_conv pair
                # rides in the exchange of the pair before it), and every speculated pair is served
                e0, e1 = C.c_uint64(), C.c_uint64()
                sp0 = [C.c_uint64() for _ in range(3)]; sp1 = [C.c_uint64() for _ in range(3)]
                capi.check(L.tb_dist_exchanges(C.byref(e0))); capi.check(L.tb_spec_stats(*[C.byref(v) for v in sp0]))
                s.step(20)
                capi.check(L.tb_dist_exchanges(C.byref(e1))); capi.check(L.tb_spec_stats(*[C.byref(v) for v in sp1]))
                assert sp1[1].value - sp0[1].value == 20, ("speculated pairs served", sp1[1].value - sp0[1].value)
                assert e1.value - e0.value == 2 * 20, ("peer exchanges per iteration", (e1.value - e0.value) / 20)
            s.close()
            abuf.release()
    capi.check(L.tb_device_sync())
    dist.barrier()
    capi.check(L.tb_dist_finalize())
    dist.destroy_process_group()
    if rank == 0:
        print("dist_worker ok: world=%d p2p=%d" % (world, p2p.value))


if __name__ == "__main__":
    main()
