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

# cone: ("soc", n_blocks, block_dim) | ("rpos", m) | ("psd", k).  kind "dense": one stacked dense A (fused DenseOp +
# ProductCone route); kind "qp": the ProbQP front-end (stock MatOp route: packed P^(1/2) via transform_sp + G, A_eq via
# transform_ge; cone RotSOC(n+2) x RPos(m) x Zero(p), qp.rs:325-338).
WORKLOADS = {
    "c3_socp_1024x64_A65536x16384": {"kind": "dense", "cone": ("soc", 1024, 64), "n": 16384},       # BASELINE config C3 (the headline)
    "socp_small_128x64_A8192x4096": {"kind": "dense", "cone": ("soc", 128, 64), "n": 4096},
    "c5_lp_A262144x65536": {"kind": "dense", "cone": ("rpos", 262144), "n": 65536},                  # C5: 68.7 GB of A (f32), meant for 8 GPUs
    "lp_A32768x65536": {"kind": "dense", "cone": ("rpos", 32768), "n": 65536},                       # one C5 shard on one GPU
    "c4_sdp_psd512_A131328x1024": {"kind": "dense", "cone": ("psd", 512), "n": 1024},                # C4: one ConePSD block 512 x 512
    "sdp_small_psd64_A2080x128": {"kind": "dense", "cone": ("psd", 64), "n": 128},
    "c2_qp_n8192_m8192_p1024": {"kind": "qp", "n": 8192, "m": 8192, "p": 1024},                      # C2: A is 17410 x 8193 through ProbQP
    "qp_small_n512_m512_p64": {"kind": "qp", "n": 512, "m": 512, "p": 64},
}
DEFAULT_WORKLOAD = "c3_socp_1024x64_A65536x16384"
SEED = 0


def cone_rows(cone):
    return cone[1] * cone[2] if cone[0] == "soc" else cone[1] if cone[0] == "rpos" else cone[1] * (cone[1] + 1) // 2


def cone_text(cone):
    if cone[0] == "soc":
        return "%d x ConeSOC(%d)" % (cone[1], cone[2])
    if cone[0] == "rpos":
        return "ConeRPos(%d)" % cone[1]
    return "ConePSD(%d x %d, sk = %d)" % (cone[1], cone[1], cone_rows(cone))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------
def svec(mat):
    """Symmetric matrix -> packed upper triangle by columns, off-diagonals scaled by sqrt(2) (cone_psd.rs:18)."""
    k = mat.shape[0]
    r, c = np.triu_indices(k)
    order = np.lexsort((r, c))                     # by column, then row: index c(c+1)/2 + r
    r, c = r[order], c[order]
    return mat[r, c] * np.where(r == c, 1.0, math.sqrt(2.0))


def instance_vectors(cone, n, seed):
    """x0, s0 in int K, y0 in int K* for the feasible-by-construction recipe."""
    rng = np.random.default_rng(seed + 12345)
    m = cone_rows(cone)
    sc = 1.0 / math.sqrt(m)        # keeps ||b||, ||c|| = O(1): tau stays > 0 and criteria_conv runs every iteration
    x0 = rng.standard_normal(n) * sc

    def interior():
        if cone[0] == "rpos":
            return (np.abs(rng.standard_normal(m)) + 0.1) * sc
        if cone[0] == "soc":
            v = rng.standard_normal((cone[1], cone[2]))
            v[:, 0] = np.linalg.norm(v[:, 1:], axis=1) + 1.0
            return v.reshape(m) * sc
        k = cone[1]
        g = rng.standard_normal((k, k))
        return svec(g.T @ g / k + np.eye(k)) * sc
    return x0, interior(), interior()


def qp_instance(n, m, p, dt):
    """ProbQP data of the benchmark_qp shape (experimental/benchmark_qp/src/main.rs:14-54) plus p equality rows:
    diagonal P ~ U(0,1) (so P^(1/2) is known in closed form), G, A_eq ~ U(-1,1)/sqrt(n) from the counter-based
    generator, h = G x0 + slack, b = A_eq x0: feasible and bounded.  Matrices are returned flat column-major."""
    from totsu_b200 import synth
    rng = np.random.default_rng(SEED + 777)
    scale = dt(1.0 / math.sqrt(n))
    g = synth.uniform_matrix(m, n, SEED, scale, dtype=dt)
    a = synth.uniform_matrix(p, n, SEED + 1, scale, dtype=dt)
    x0 = rng.standard_normal(n) / math.sqrt(n)
    h = (g.astype(np.float64) @ x0 + np.abs(rng.standard_normal(m)) * 0.1 + 0.01).astype(dt)
    b = (a.astype(np.float64) @ x0).astype(dt)
    q = (rng.standard_normal(n) / math.sqrt(n)).astype(dt)
    pdiag = rng.uniform(0.05, 1.0, n)
    psqrt = np.zeros(n * (n + 1) // 2, dtype=dt)
    idx = np.arange(n, dtype=np.int64)
    psqrt[idx * (idx + 1) // 2 + idx] = np.sqrt(pdiag).astype(dt)
    return psqrt, q, g.reshape(-1, order="F"), h, a.reshape(-1, order="F"), b


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


def qp_stacked(qn, qm, qp_, dt, qdata=None):
    """ProbQP's composite operator written out as one dense matrix (qp.rs:98-140, rows [0; q^T,-1; -P^(1/2); G; A_eq],
    columns (x, t)) with b = [1; 0; 0_n; h; b_eq] (qp.rs:196-215) and c = [0_n; 1] (qp.rs:20-45); zero rows are appended
    to the ConeZero block so the row count is a multiple of 4 (16-byte aligned columns for the streaming kernel).
    Returns (A column-major 2-D, b, c, n_pad_rows)."""
    psqrt, q, g, h, a_eq, b_eq = qdata if qdata is not None else qp_instance(qn, qm, qp_, dt)
    m0, n = (2 + qn) + qm + qp_, qn + 1
    pad = (-m0) % 4
    m = m0 + pad
    stacked = np.zeros((m, n), dtype=dt, order="F")
    stacked[1, :qn] = q; stacked[1, qn] = -1.0
    idx = np.arange(qn, dtype=np.int64)
    stacked[2 + idx, idx] = -psqrt[idx * (idx + 1) // 2 + idx]           # P is diagonal in these workloads
    stacked[2 + qn:2 + qn + qm, :qn] = np.asarray(g).reshape(qm, qn, order="F")
    stacked[2 + qn + qm:2 + qn + qm + qp_, :qn] = np.asarray(a_eq).reshape(qp_, qn, order="F")
    b = np.zeros(m, dtype=dt); b[0] = 1.0; b[2 + qn:2 + qn + qm] = h; b[2 + qn + qm:2 + qn + qm + qp_] = b_eq
    c = np.zeros(n, dtype=dt); c[qn] = 1.0
    return stacked, b, c, pad


# ------------------------------------------------------------------------------------------------------------
def cpu_reference_leg(spec, steps, warmup, sample_blocks=None, threads=None):
    """The reference's CPU path for this workload: the oracle's port of the matching front-end + F64LAPACK (f64,
    OpenBLAS instead of MKL).  SOCP / LP: a bounded row sample of the same A (the first `sample_blocks` cone blocks),
    timed per iteration and scaled linearly in rows (the iteration is dgemv-bound).  SDP and QP: the full workload
    (a PSD block cannot be row-sampled; the QP fits)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import totsu_oracle as O
    from totsu_b200 import synth
    cores = threads or os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to every rank: give the BLAS pool all host cores back and report what it really uses
    try:
        from threadpoolctl import threadpool_limits, threadpool_info
        threadpool_limits(limits=cores)
        blas = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        if blas:
            cores = max(blas)
    except Exception:
        env = os.environ.get("OMP_NUM_THREADS")
        if env and env.isdigit():
            cores = min(cores, int(env))
    if spec["kind"] == "qp":
        n, m, p = spec["n"], spec["m"], spec["p"]
        psqrt, q, g, h, a, b = qp_instance(n, m, p, np.float32)
        f8 = lambda v: np.asarray(v, dtype=np.float64)
        prob = O.ProbQP(O.MatBuild(O.MatType.SymPack(n), f8(psqrt)), O.MatBuild(O.MatType.General(n, 1), f8(q)),
                        O.MatBuild(O.MatType.General(m, n), f8(g)), O.MatBuild(O.MatType.General(m, 1), f8(h)),
                        O.MatBuild(O.MatType.General(p, n), f8(a)), O.MatBuild(O.MatType.General(p, 1), f8(b)), 1e-12, p_is_sqrt=True)
        del g, a, psqrt
        return _time_oracle(O, prob, steps, warmup, cores, 1.0,
                            "the full workload (f64; dspmv on the packed P^(1/2) + dgemv on G and A_eq like ProbQP, OpenBLAS via numpy/scipy instead of MKL)")
    cone, n = spec["cone"], spec["n"]
    m = cone_rows(cone)
    scale = np.float32(1.0 / math.sqrt(n))
    x0, s0, y0 = instance_vectors(cone, n, SEED)
    if cone[0] == "psd":
        k = cone[1]
        a = np.empty((m, n), dtype=np.float64, order="F")
        for c0 in range(0, n, 128):          # column panels keep the generator's temporaries small
            a[:, c0:c0 + 128] = synth.uniform_matrix(m, min(128, n - c0), SEED, scale, dtype=np.float32, cols=np.arange(c0, min(n, c0 + 128)))
        b = (a @ x0 + s0).astype(np.float32).astype(np.float64)
        c = (-(a.T @ y0)).astype(np.float32).astype(np.float64)

        class _Dense:      # ProbSDP's operator tuple with p = 0 (sdp.rs:75-97: one MatOp symmat_f, sk x n) over the same A
            def problem(self):
                op_c = O.MatOp(O.MatType.General(n, 1), c)
                op_a = O.MatOp(O.MatType.General(m, n), a.reshape(-1, order="F"))
                op_b = O.MatOp(O.MatType.General(m, 1), b)
                cone_o = O._ProductCone([(O.ConePSD(np.zeros(O.ConePSD.query_worklen(m)), 1e-12), m)])
                return op_c, op_a, op_b, cone_o, np.zeros(O.Solver.query_worklen((m, n)))
        return _time_oracle(O, _Dense(), steps, warmup, cores, 1.0,
                            "the full workload (f64; one dgemv per op like ProbSDP's symmat_f, ConePSD::proj = LAPACK dsyevr + dsyr loop "
                            "on %d x %d, OpenBLAS via numpy/scipy instead of MKL)" % (k, k))
    nblk, bdim = (cone[1], cone[2]) if cone[0] == "soc" else (cone[1], 1)
    if sample_blocks is None:
        sample_blocks = max(1, min(nblk, 64 if bdim > 1 else 2048))
    ms = sample_blocks * bdim
    a32 = synth.uniform_matrix(ms, n, SEED, scale, dtype=np.float32)
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
    ap.add_argument("--speculation", type=int, default=1, help="compute the next pair's products in the current read of A when its inputs are already final (csrc/gemv.cu)")
    ap.add_argument("--vprog", type=int, default=1, help="run the small vector commands between streaming launches as one launch per batch (csrc/vprog.cu)")
    ap.add_argument("--route", default="fused", choices=["fused", "stock"],
                    help="QP workloads: 'fused' = ProbQP's stacked operator as one dense A (DenseOp + ProductCone), "
                         "'stock' = the ProbQP front-end itself (MatOp per block, stock cones)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    spec = WORKLOADS[args.workload]
    is_qp = spec["kind"] == "qp"
    qp_stock = is_qp and args.route == "stock"
    qp_fused = is_qp and args.route == "fused"
    esize = 4 if args.dtype == "f32" else 8
    if is_qp:
        qn, qm, qp_ = spec["n"], spec["m"], spec["p"]
        m, n = (2 + qn) + qm + qp_, qn + 1                   # the stacked operator ProbQP builds (qp.rs:325-331)
        qpad = (-m) % 4 if qp_fused else 0                   # zero rows (extra ConeZero coordinates, b = 0) so columns start 16-byte aligned
        dense_elems = qm * qn + qp_ * qn + qn * (qn + 1) // 2     # G + A_eq + packed P^(1/2): what one op / trans_op reads in the reference's formulation
        config = {"workload": args.workload, "cone": "ConeRotSOC(%d) x ConeRPos(%d) x ConeZero(%d)" % (qn + 2, qm, qp_ + qpad),
                  "A": "%d x %d through ProbQP: G %d x %d + A_eq %d x %d dense column-major, P^(1/2) upper-packed %d x %d" % (m, n, qm, qn, qp_, qn, qn, qn),
                  "route": "stock MatOp route (the ProbQP front-end itself: transform_ge + transform_sp per block, stock cones, no pair fusion)" if qp_stock else
                           "fused DenseOp + ProductCone handed to the unmodified Solver: ProbQP's rows [0; q^T,-1; -P^(1/2); G; A_eq] (qp.rs:325-331) stacked into "
                           "one dense %d x %d A (%d zero rows appended for alignment); algorithmic bytes stay the reference formulation's (packed P^(1/2))" % (m + qpad, n, qpad)}
        m += qpad
    else:
        cone, n = spec["cone"], spec["n"]
        m = cone_rows(cone)
        dense_elems = m * n
        config = {"workload": args.workload, "cone": cone_text(cone), "A": "%d x %d dense column-major" % (m, n),
                  "route": "fused DenseOp + ProductCone handed to the unmodified Solver"}
    config["l2"] = "matrices (%.2f GB) are larger than L2; no explicit flush" % (dense_elems * esize / 1e9) if dense_elems * esize > 200e6 else \
                   "matrices (%.3f GB) fit in the 126 MB L2: numbers are L2-resident, not HBM" % (dense_elems * esize / 1e9)
    steps, warmup = args.steps, max(args.warmup, 3)

    if args.impl == "reference":
        if rank != 0:
            return
        ref_steps = min(steps, 20)
        ref = cpu_reference_leg(spec, ref_steps, min(warmup, 3), args.cpu_sample_blocks)
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
    torch.cuda.set_device(local_rank)
    capi.init(local_rank)
    L = capi.lib()
    capi.check(L.tb_set_pair_fusion(1 if args.pair_fusion else 0))
    capi.check(L.tb_set_vprog(1 if args.vprog else 0))
    capi.check(L.tb_set_speculation(1 if args.speculation else 0))
    if world > 1:
        if is_qp or spec["cone"][0] == "psd":
            raise SystemExit("%s is a single-GPU configuration (a PSD block / the QP front-end does not shard)" % args.workload)
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

    abuf = None
    if qp_stock:
        qdata = qp_instance(qn, qm, qp_, dt)
        m_loc = m
        matrix_h2d = dense_elems * esize            # wrapped host arrays: uploaded on first use inside Solver::solve

        def new_session():
            return host.Session.qp(dt, qdata[0], qdata[1], qdata[2], qdata[3], qdata[4], qdata[5], 1e-12, p_is_sqrt=True, col_major=True)
    elif qp_fused:
        m_loc, matrix_h2d = m, 0                    # the stacked A is uploaded once into a backend buffer before the timed regions
        stacked, b, c, _ = qp_stacked(qn, qm, qp_, dt)
        assert stacked.shape == (m, n)
        abuf = capi.Buf(dtype=dt, length=m * n)
        abuf.upload(stacked.reshape(-1, order="F"))
        del stacked
        blocks = [(capi.CONE_ROTSOC, qn + 2), (capi.CONE_RPOS, qm), (capi.CONE_ZERO, qp_ + qpad)]

        def new_session():
            return host.Session.dense(dt, abuf.view(), m, n, c, b, blocks, fused_op=True, fused_cone=True)
    else:
        from totsu_b200 import shard
        if cone[0] == "soc":
            blocks = [(capi.CONE_SOC, cone[2])] * cone[1]
        elif cone[0] == "rpos":
            blocks = [(capi.CONE_RPOS, m)]
        else:
            blocks = [(capi.CONE_PSD, m)]
        row_off, m_loc = shard.row_shards(blocks, world)[rank]
        matrix_h2d = 0                              # A is generated in HBM
        # ---- instance: A generated in HBM (shard), b = A x0 + s0, c = -A^T y0 through the backend itself
        abuf = capi.Buf(dtype=dt, length=m_loc * n)
        scale = dt(1.0 / math.sqrt(n))
        capi.check(capi.fn("tb_fill_uniform", dt)(abuf.view(), m_loc, n, row_off, SEED, scale))
        hop = C.c_int64()
        capi.check(L.tb_denseop_create(capi.dtype_id(dt), abuf.view(), m_loc, n, row_off, m, C.byref(hop)))
        x0, s0, y0 = instance_vectors(cone, n, SEED)
        b = s0.astype(dt); c = np.zeros(n, dtype=dt)
        bx, by, bb, bc = capi.Buf(x0.astype(dt), mutable=False), capi.Buf(y0.astype(dt), mutable=False), capi.Buf(b), capi.Buf(c)
        capi.check(capi.fn("tb_denseop_apply", dt)(hop.value, 0, 1.0, bx.view(), 1.0, bb.view()))
        capi.check(capi.fn("tb_denseop_apply", dt)(hop.value, 1, -1.0, by.view(), 0.0, bc.view()))
        for bf in (bx, by, bb, bc):
            bf.release()
        capi.check(L.tb_denseop_destroy(hop.value))

        def new_session():
            return host.Session.dense(dt, abuf.view(), m_loc, n, c, b, blocks, fused_op=True, fused_cone=True, row_offset=row_off, m_total=m)

    stream = torch.cuda.ExternalStream(capi.stream_ptr(), device=torch.device("cuda", local_rank))

    def barrier():
        capi.check(L.tb_device_sync())
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    # ---- device-resident timing: K iterations between two events on the library's stream
    s = new_session()
    assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0, device_precond=not qp_stock) == "None"
    if os.environ.get("BENCH_DEBUG"):
        for _ in range(min(warmup, 5)):
            s.step(1)
            print("dbg iter %d tau %.4e res %.4e %.4e %.4e" % (s.last.i, s.last.val_tau, s.last.c0, s.last.c1, s.last.c2), file=sys.stderr)
        print("dbg norms", s.norms(), file=sys.stderr)
        warmup = max(0, warmup - 5)
    s.step(warmup)
    barrier()
    clocks = Clocks(local_rank) if rank == 0 else None
    vl0, vo0 = C.c_uint64(), C.c_uint64()
    capi.check(L.tb_vprog_stats(C.byref(vl0), C.byref(vo0)))
    sp0 = [C.c_uint64() for _ in range(3)]
    capi.check(L.tb_spec_stats(*[C.byref(v) for v in sp0]))
    l0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hw_s, hw_n = C.c_double(), C.c_uint64()
    capi.check(L.tb_host_wait_stats(C.byref(hw_s), C.byref(hw_n)))        # reset
    e0.record(stream)
    th0 = time.perf_counter()
    s.step(steps)
    capi.check(L.tb_flush())             # nothing recorded or parked may be left behind the closing event
    th1 = time.perf_counter()
    e1.record(stream)
    capi.check(L.tb_host_wait_stats(C.byref(hw_s), C.byref(hw_n)))
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = capi.launch_count() - l0
    vl1, vo1 = C.c_uint64(), C.c_uint64()
    capi.check(L.tb_vprog_stats(C.byref(vl1), C.byref(vo1)))
    sp1 = [C.c_uint64() for _ in range(3)]
    capi.check(L.tb_spec_stats(*[C.byref(v) for v in sp1]))
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

    # ---- end to end: one whole Solver::solve through the public API with host buffers (work, c, b - and for the QP
    # front-end the matrices - in host memory; scalars cross the boundary every iteration; the solution is read back)
    barrier()
    s = new_session()
    barrier()                            # session construction differs per rank: start the end-to-end clock together
    t0 = time.perf_counter()
    st = s.begin(max_iter=steps, eps_acc=0.0, eps_inf=0.0, device_precond=not qp_stock)
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
    h2d = (worklen * esize + matrix_h2d) / steps + 3 * esize          # work (+ wrapped matrices) upload amortised + tau/kappa/unit scalars per iteration
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
        served = (sp1[1].value - sp0[1].value) > 0
        wants = (["stream_kernel<float, 1, 1>", "stream_kernel<float, 2, 2>"] if served else ["stream_kernel<float, 1, 1>"]) if pairs_per_iter else ["stream_kernel<float, "]
        vals = [v["dram_bytes_per_launch"] for k, v in tj.items() if k.startswith(args.workload + "|") and any(w in k for w in wants)]
        if vals:
            traffic = sum(vals) / len(vals)
    if nl.value:
        # ALGORITHMIC bytes per launch: one read of this rank's dense block per op / trans_op served.  With
        # the lazy pairing a launch serves an op AND a trans_op from ONE read of A, so the algorithmic figure is twice the
        # bytes actually streamed and `frac` may exceed 1; `streamed_*` is the un-doubled DRAM-side figure.  (QP: the
        # launches timed are the transform_ge ones on G and A_eq; the packed P^(1/2) goes through spmv_kernel.)
        avg_ms = kms.value / nl.value
        ge_elems = (qm * qn + qp_ * qn) if qp_stock else dense_elems if qp_fused else m_loc * n
        alg_per_launch = 6.0 * prof_iters * ge_elems * esize / nl.value
        ach = alg_per_launch / (avg_ms * 1e-3) / 1e9
        streamed = (kbytes.value / nl.value) / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "stream_kernel (TMA bulk-copy matvec)", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full capture)" if traffic else None, "peak_source": peak_src, "launches_timed": int(nl.value), "avg_launch_ms": avg_ms,
                "algorithmic_bytes_per_launch": alg_per_launch, "matvecs_per_launch": 6.0 * prof_iters * (2 if qp_stock else 1) / nl.value,
                "streamed_bytes_per_launch": kbytes.value / nl.value, "streamed_gbs": streamed, "streamed_frac": streamed / peak,
                "note": ("op/trans_op pairs share one read of A (lazy pairing behind tb_denseop_apply) and, with speculative pairing, the "
                         "criteria_conv pair rides on the preceding pass: achieved/frac use the un-fused algorithmic bytes "
                         "(6 reads of A per iteration, SURVEY 8d) and exceed 1 by construction; streamed_frac is bytes actually read / peak") if pairs_per_iter else None}
    abytes_iter = 6.0 * dense_elems * esize
    line = {"metric": "solver iterations/sec", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "Solver::solve (begin + %d iterations + end) through the host layer, work/c/b%s in host memory" % (steps, " and the matrices" if qp_stock else ""),
                    "breakdown_s": {"begin": t1 - t0, "iterate": t2 - t1, "end_and_readback": t0 + t_e2e_local - t2}},
            "gpu_launches": int(launches), "roofline": roof,
            "hbm_frac_whole_iteration": abytes_iter * value / (world * peak * 1e9),
            "algorithmic_bytes_per_iteration": abytes_iter, "pair_fusion": bool(args.pair_fusion), "pairs_fused_per_iteration": pairs_per_iter,
            "speculative_pairing": {"enabled": bool(args.speculation), "passes_with_speculation_per_iteration": (sp1[0].value - sp0[0].value) / steps,
                                    "pairs_served_without_reading_A_per_iteration": (sp1[1].value - sp0[1].value) / steps,
                                    "dropped": sp1[2].value - sp0[2].value,
                                    "note": "the criteria_conv pair's products are computed during the preceding pass over A (its inputs are already final): "
                                            "2 reads of A per iteration instead of 3, bit-identical results"},
            "vector_programs": {"enabled": bool(args.vprog), "launches_per_iteration": (vl1.value - vl0.value) / steps,
                                "micro_ops_per_iteration": (vo1.value - vo0.value) / steps,
                                "note": "small vector commands recorded into one cluster launch per batch (csrc/vprog.cu); each program counts as one of gpu_launches"},
            "host": {"loop_s": th1 - th0, "waiting_for_device_s": hw_s.value, "host_visible_scalars_per_iteration": hw_n.value / steps,
                     "note": "host time of the timed loop and the part of it spent spinning on device results: the rest is issuing launches"},
            "clocks": clk, "last_residuals": [last.c0, last.c1, last.c2], "status_e2e": st}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_leg(spec, 6, 2, args.cpu_sample_blocks)
        print(json.dumps(line))
    if abuf is not None:
        abuf.release()
    if world > 1:
        capi.check(L.tb_dist_finalize())
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
