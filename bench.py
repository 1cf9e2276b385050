#!/usr/bin/env python
"""bench.py - solver iterations/second of the Totsu first-order conic iteration on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's arm (hand-written sm_100a kernels)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port of totsu_f64lapack)

A "step" is one solver iteration = update_vecs + criteria_conv (solver.rs:382-386) = 3 A.op + 3 A.trans_op +
2 cone projections + ~30 vector ops + 6 host-visible scalars, on the workload BASELINE.json's metric is quoted
on: config C3, random SOCP with 1024 ConeSOC blocks of dim 64, dense A 65536 x 16384, fp32, generated in HBM.
For N > 1 (torchrun, one process per GPU) A is row-sharded on cone-block boundaries; A*x slices are all-gathered
and A^T*y partials all-reduced over NCCL; total work is fixed ("strong" scaling).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_blocks, block_dim, n)          m = n_blocks * block_dim; block_dim == 1: ConeRPos(m) (an LP)
    "c3_socp_1024x64_A65536x16384": (1024, 64, 16384),
    "socp_small_128x64_A8192x4096": (128, 64, 4096),
    "c5_lp_A262144x65536": (262144, 1, 65536),          # BASELINE config C5: 68.7 GB of A (f32), meant for 8 GPUs
    "lp_A32768x65536": (32768, 1, 65536),               # one C5 shard on one GPU
}
DEFAULT_WORKLOAD = "c3_socp_1024x64_A65536x16384"
SEED = 0


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------
def instance_vectors(nblk, bdim, n, seed):
    """x0, s0 in int K, y0 in int K* for the feasible-by-construction recipe."""
    rng = np.random.default_rng(seed + 12345)
    m = nblk * bdim
    sc = 1.0 / math.sqrt(m)        # keeps ||b||, ||c|| = O(1): tau stays > 0 and criteria_conv runs every iteration
    x0 = rng.standard_normal(n) * sc

    def interior():
        if bdim == 1:
            return (np.abs(rng.standard_normal(m)) + 0.1) * sc
        v = rng.standard_normal((nblk, bdim))
        v[:, 0] = np.linalg.norm(v[:, 1:], axis=1) + 1.0
        return v.reshape(m) * sc
    return x0, interior(), interior()


class Clocks:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out["sm_mhz"] = float(np.median(sm)); out["sm_max_mhz"] = float(max(mx)); out["reasons"] = sorted(reasons)
            out["samples"] = len(sm)
        return out


# ------------------------------------------------------------------------------------------------------------
def cpu_reference_leg(nblk, bdim, n, steps, warmup, sample_blocks=None, threads=None):
    """The reference's CPU path for this workload: the oracle's port of ProbSOCP + F64LAPACK (per-block dgemv, f64,
    OpenBLAS instead of MKL) on a bounded row sample of the same A (the first `sample_blocks` cone blocks), timed
    per iteration and scaled linearly in rows (the iteration is dgemv-bound)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import totsu_oracle as O
    from totsu_b200 import synth
    cores = threads or os.cpu_count() or 1
    if sample_blocks is None:
        sample_blocks = max(1, min(nblk, 64 if bdim > 1 else 2048))
    ms = sample_blocks * bdim
    scale = np.float32(1.0 / math.sqrt(n))
    a32 = synth.uniform_matrix(ms, n, SEED, scale, dtype=np.float32)
    x0, s0, y0 = instance_vectors(nblk, bdim, n, SEED)
    a = a32.astype(np.float64)
    b = (a @ x0 + s0[:ms]).astype(np.float32).astype(np.float64)
    c = (-(a.T @ y0[:ms])).astype(np.float32).astype(np.float64)
    if bdim == 1:
        # ProbLP's shape (lp.rs:222-338): one MatOp G (ms x n) + ConeRPos(ms), no equalities
        prob = O.ProbLP(O.MatBuild(O.MatType.General(n, 1), c), O.MatBuild(O.MatType.General(ms, n), np.asfortranarray(a).reshape(-1, order="F")),
                        O.MatBuild(O.MatType.General(ms, 1), b), O.MatBuild(O.MatType.General(0, n)), O.MatBuild(O.MatType.General(0, 1)))
        del a
        return _time_oracle(O, prob, steps, warmup, cores, sample_blocks / nblk,
                            "first %d of %d rows of the same A (f64, one dgemv per op like ProbLP, OpenBLAS via numpy instead of MKL)" % (ms, nblk))
    # ProbSOCP's shape (socp.rs:359-366): rows of block i = [-c_i^T; -G_i], h_i, d_i from b
    gs, hs, cs, ds = [], [], [], []
    for i in range(sample_blocks):
        blk = a[i * bdim:(i + 1) * bdim, :]
        cs.append(O.MatBuild(O.MatType.General(n, 1), -blk[0, :]))
        gs.append(O.MatBuild(O.MatType.General(bdim - 1, n), np.asfortranarray(-blk[1:, :]).reshape(-1, order="F")))
        ds.append(float(b[i * bdim]))
        hs.append(O.MatBuild(O.MatType.General(bdim - 1, 1), b[i * bdim + 1:(i + 1) * bdim]))
    del a
    prob = O.ProbSOCP(O.MatBuild(O.MatType.General(n, 1), c), gs, hs, cs, ds,
                      O.MatBuild(O.MatType.General(0, n)), O.MatBuild(O.MatType.General(0, 1)))
    return _time_oracle(O, prob, steps, warmup, cores, sample_blocks / nblk,
                        "first %d of %d SOC blocks (%d x %d rows of the same A, f64, per-block dgemv like ProbSOCP, OpenBLAS via numpy instead of MKL)"
                        % (sample_blocks, nblk, ms, n))


def _time_oracle(O, prob, steps, warmup, cores, frac, what):
    s = O.Solver()
    s.par.max_iter = warmup + steps + 1
    s.par.eps_acc = 0.0
    s.par.eps_inf = 0.0
    times = []
    orig = s._update_vecs

    def timed_update(*args):
        times.append(time.perf_counter())
        return orig(*args)
    s._update_vecs = timed_update
    try:
        s.solve(prob.problem())
    except O.SolverError:
        pass
    times.append(time.perf_counter())
    t = times[warmup:warmup + steps + 1]
    per_iter = (t[-1] - t[0]) / max(1, len(t) - 1)
    value = (1.0 / per_iter) * frac
    return {"value": value, "unit": "iterations/s", "cores": cores, "kind": "port",
            "sample": "%s; %d iterations timed, %.4f s/iter on the sample, scaled x%g linearly in rows" % (what, len(t) - 1, per_iter, frac)}


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS.keys()))
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-blocks", type=int, default=None)
    ap.add_argument("--pair-fusion", type=int, default=1, help="serve op/trans_op pairs with one read of A when the backend can")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nblk, bdim, n = WORKLOADS[args.workload]
    m = nblk * bdim
    steps, warmup = args.steps, max(args.warmup, 3)
    config = {"workload": args.workload, "cone": ("%d x ConeSOC(%d)" % (nblk, bdim)) if bdim > 1 else "ConeRPos(%d)" % m, "A": "%d x %d dense column-major" % (m, n),
              "l2": "A (%.2f GB) is larger than L2; no explicit flush" % (m * n * (4 if args.dtype == "f32" else 8) / 1e9)}

    if args.impl == "reference":
        if rank != 0:
            return
        ref_steps = min(steps, 20)
        ref = cpu_reference_leg(nblk, bdim, n, ref_steps, min(warmup, 3), args.cpu_sample_blocks)
        line = {"impl": "reference", "metric": "solver iterations/sec", "value": ref["value"], "unit": "iterations/s", "n_gpus": 0,
                "steps": ref_steps, "warmup": min(warmup, 3), "ms_per_step": 1e3 / ref["value"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": ref,
                "e2e": {"value": ref["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    from totsu_b200 import capi, host
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: totsu_b200 has no CPU fallback")
    dt = np.float32 if args.dtype == "f32" else np.float64
    esize = np.dtype(dt).itemsize
    torch.cuda.set_device(local_rank)
    capi.init(local_rank)
    L = capi.lib()
    capi.check(L.tb_set_pair_fusion(1 if args.pair_fusion else 0))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idbuf = (C.c_ubyte * capi.NCCL_ID_BYTES)()
        if rank == 0:
            capi.check(L.tb_dist_unique_id(idbuf))
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        idbuf = (C.c_ubyte * capi.NCCL_ID_BYTES)(*t.cpu().tolist())
        capi.check(L.tb_dist_init(rank, world, idbuf))
        p2p = C.c_int()
        capi.check(L.tb_dist_p2p_enabled(C.byref(p2p)))
        config["collectives"] = ("peer stores fused into the matvec epilogue (cudaIpc staging over NVLink)" if p2p.value
                                 else "ncclAllGather / ncclAllReduce")
        config["parallelism"] = "A row-sharded x%d on cone-block boundaries, vectors replicated" % world
    from totsu_b200 import shard
    blocks = [(capi.CONE_SOC, bdim)] * nblk if bdim > 1 else [(capi.CONE_RPOS, m)]
    row_off, m_loc = shard.row_shards(blocks, world)[rank]

    # ---- instance: A generated in HBM (shard), b = A x0 + s0, c = -A^T y0 through the backend itself
    abuf = capi.Buf(dtype=dt, length=m_loc * n)
    scale = dt(1.0 / math.sqrt(n))
    capi.check(capi.fn("tb_fill_uniform", dt)(abuf.view(), m_loc, n, row_off, SEED, scale))
    hop = C.c_int64()
    capi.check(L.tb_denseop_create(capi.dtype_id(dt), abuf.view(), m_loc, n, row_off, m, C.byref(hop)))
    x0, s0, y0 = instance_vectors(nblk, bdim, n, SEED)
    b = s0.astype(dt); c = np.zeros(n, dtype=dt)
    bx, by, bb, bc = capi.Buf(x0.astype(dt), mutable=False), capi.Buf(y0.astype(dt), mutable=False), capi.Buf(b), capi.Buf(c)
    capi.check(capi.fn("tb_denseop_apply", dt)(hop.value, 0, 1.0, bx.view(), 1.0, bb.view()))
    capi.check(capi.fn("tb_denseop_apply", dt)(hop.value, 1, -1.0, by.view(), 0.0, bc.view()))
    for bf in (bx, by, bb, bc):
        bf.release()
    capi.check(L.tb_denseop_destroy(hop.value))

    stream = torch.cuda.ExternalStream(capi.stream_ptr(), device=torch.device("cuda", local_rank))

    def barrier():
        capi.check(L.tb_device_sync())
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    def new_session():
        s = host.Session.dense(dt, abuf.view(), m_loc, n, c, b, blocks, fused_op=True, fused_cone=True, row_offset=row_off, m_total=m)
        return s

    # ---- device-resident timing: K iterations between two events on the library's stream
    s = new_session()
    assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0, device_precond=True) == "None"
    if os.environ.get("BENCH_DEBUG"):
        for _ in range(min(warmup, 5)):
            s.step(1)
            print("dbg iter %d tau %.4e res %.4e %.4e %.4e" % (s.last.i, s.last.val_tau, s.last.c0, s.last.c1, s.last.c2), file=sys.stderr)
        print("dbg b[:4]", b[:4], "c[:4]", c[:4], "norms", s.norms(), file=sys.stderr)
        warmup = max(0, warmup - 5)
    s.step(warmup)
    barrier()
    clocks = Clocks(local_rank) if rank == 0 else None
    l0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    s.step(steps)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = capi.launch_count() - l0
    clk = clocks.stop() if clocks else None
    last = s.last
    # ---- the same region again with per-launch events around the streaming matvec (roofline numerator)
    capi.check(L.tb_prof_enable(1))
    prof_iters = min(steps, 20)
    pf0 = capi.pairs_fused()
    s.step(prof_iters)
    pairs_per_iter = (capi.pairs_fused() - pf0) / prof_iters
    nl, kms, kbytes = C.c_uint64(), C.c_double(), C.c_double()
    capi.check(L.tb_prof_read(C.byref(nl), C.byref(kms), C.byref(kbytes)))
    capi.check(L.tb_prof_enable(0))
    s.close()

    # ---- end to end: one whole Solver::solve through the public API with host buffers (work, c, b on the host;
    # scalars cross the boundary every iteration; the solution is read back)
    barrier()
    s = new_session()
    t0 = time.perf_counter()
    st = s.begin(max_iter=steps, eps_acc=0.0, eps_inf=0.0, device_precond=True)
    assert st == "None"
    t1 = time.perf_counter()
    st, _ = s.run()
    t2 = time.perf_counter()
    s.end()
    xs, ys = s.solution()
    capi.check(L.tb_device_sync())
    t_e2e = time.perf_counter() - t0
    t_e2e_local = t_e2e
    if os.environ.get("BENCH_DEBUG"):
        print("dbg e2e: begin %.4f s, run %.4f s, end+readback %.4f s" % (t1 - t0, t2 - t1, t0 + t_e2e - t2), file=sys.stderr)
    s.close()
    worklen = 4 * (n + 2 * m + 1) + 2 * (n + m + 1)
    h2d = worklen * esize / steps + 3 * esize          # work upload amortised + tau/kappa/unit scalars per iteration
    d2h = (n + m) * esize / steps + 6 * esize          # solution readback amortised + tau, kappa, g_x, g_y, |p|, |d|

    # ---- max over ranks
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([ms_total, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, t_e2e = float(tt[0]), float(tt[1])
    value = steps / (ms_total * 1e-3)
    e2e_value = steps / t_e2e
    peak, peak_src = measured_peaks()
    roof = None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.dtype == "f32" and world == 1:
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this workload
        tj = json.load(open(tpath))
        want = "stream_kernel<float, 1, 1>" if pairs_per_iter else "stream_kernel<float, "
        vals = [v["dram_bytes_per_launch"] for k, v in tj.items() if k.startswith(args.workload + "|") and want in k]
        if vals:
            traffic = sum(vals) / len(vals)
    if nl.value:
        # ALGORITHMIC bytes per launch: one read of this rank's A per op / trans_op served.  With the
        # lazy pairing a launch serves an op AND a trans_op from ONE read of A, so the algorithmic figure is twice the
        # bytes actually streamed and `frac` may exceed 1; `streamed_*` is the un-doubled DRAM-side figure.
        avg_ms = kms.value / nl.value
        alg_per_launch = 6.0 * prof_iters * m_loc * n * esize / nl.value
        ach = alg_per_launch / (avg_ms * 1e-3) / 1e9
        streamed = (kbytes.value / nl.value) / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "stream_kernel (TMA bulk-copy matvec)", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full capture)" if traffic else None, "peak_source": peak_src, "launches_timed": int(nl.value), "avg_launch_ms": avg_ms,
                "algorithmic_bytes_per_launch": alg_per_launch, "matvecs_per_launch": 6.0 * prof_iters / nl.value,
                "streamed_bytes_per_launch": kbytes.value / nl.value, "streamed_gbs": streamed, "streamed_frac": streamed / peak,
                "note": ("op/trans_op pairs share one read of A (lazy pairing behind tb_denseop_apply): achieved/frac use the un-fused "
                         "algorithmic bytes and can exceed 1; streamed_frac is bytes actually read / peak") if pairs_per_iter else None}
    abytes_iter = 6.0 * m * n * esize
    line = {"metric": "solver iterations/sec", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "Solver::solve (begin + %d iterations + end) through the host layer, work/c/b in host memory" % steps,
                    "breakdown_s": {"begin": t1 - t0, "iterate": t2 - t1, "end_and_readback": t0 + t_e2e_local - t2}},
            "gpu_launches": int(launches), "roofline": roof,
            "hbm_frac_whole_iteration": abytes_iter * value / (world * peak * 1e9),
            "algorithmic_bytes_per_iteration": abytes_iter, "pair_fusion": bool(args.pair_fusion), "pairs_fused_per_iteration": pairs_per_iter,
            "clocks": clk, "last_residuals": [last.c0, last.c1, last.c2], "status_e2e": st}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_leg(nblk, bdim, n, 6, 2, args.cpu_sample_blocks)
        print(json.dumps(line))
    abuf.release()
    if world > 1:
        capi.check(L.tb_dist_finalize())
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
