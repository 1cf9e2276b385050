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
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line), one sampler per rank on its own GPU,
    started before warm-up and stopped after the end-to-end leg; `stop(t0, t1)` keeps the samples stamped inside the timed
    windows [t0, t1] (falling back to every sample taken under load when the windows are shorter than the sampling period)."""

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t0=None, t1=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), [v.lower().startswith("active") for v in f[5:9]]))
            except ValueError:
                continue
        os.unlink(self.path)
        inside = [r for r in rows if t0 is not None and t0 - 0.05 <= r[0] <= t1 + 0.05]
        use = inside if len(inside) >= 3 else rows
        if use:
            reasons = set()
            for r in use:
                for name, on in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3]):
                    if on:
                        reasons.add(name)
            out["sm_mhz"] = float(np.median([r[1] for r in use])); out["sm_max_mhz"] = float(max(r[2] for r in use)); out["reasons"] = sorted(reasons)
            out["samples"] = len(use); out["samples_inside_timed_windows"] = len(inside)
            out["span"] = "timed windows" if use is inside else "warm-up .. end-to-end leg (GPU under load throughout)"
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
def _oracle_modules():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import totsu_oracle as O
    import cpu_workloads as W
    return O, W


def _blas_threads(threads=None):
    """Give the BLAS pool all host cores (torchrun exports OMP_NUM_THREADS=1 to every rank) and report what it really uses."""
    cores = threads or os.cpu_count() or 1
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
    return cores


CONE_KIND = {"soc": "soc", "rpos": "rpos", "psd": "psd"}


def oracle_blocks(cone, rows=None):
    """bench cone spec -> [(kind, len)] for oracle/cpu_workloads.oracle_cone (optionally only the first `rows` rows)."""
    if cone[0] == "soc":
        nb = cone[1] if rows is None else rows // cone[2]
        return [("soc", cone[2])] * nb
    if cone[0] == "rpos":
        return [("rpos", cone[1] if rows is None else rows)]
    return [("psd", cone_rows(cone))]


def cpu_reference_leg(spec, steps, warmup, sample_blocks=None, threads=None, dtype="f32", budget_s=None):
    """The reference's CPU path for this workload: the oracle's port of the matching front-end + F64LAPACK (f64, OpenBLAS
    instead of MKL), on inputs rounded like the device's `dtype`.

    SOCP: `value` is the reference's OWN route - ProbSOCP, one skinny dgemv + one dot per cone block per op
    (socp.rs:83-124) - and `stacked` is the same rows behind ONE MatOp (one dgemv per op, what the fused device route
    computes): the apples-to-apples CPU number.  The whole workload is run when (warmup + steps) iterations of it fit the
    time budget (TB_REF_BUDGET_S, default 200 s) and host memory; otherwise the largest power-of-two fraction of the cone
    blocks that does, scaled linearly in rows (the iteration is dgemv-bound) and labelled as such.  LP: ProbLP is already
    one dgemv per op.  SDP and QP: always the full workload (a PSD block cannot be row-sampled; the QP fits)."""
    O, W = _oracle_modules()
    from totsu_b200 import synth
    cores = _blas_threads(threads)
    as_f32 = dtype == "f32"
    rdt = np.float32 if as_f32 else np.float64
    if budget_s is None:
        budget_s = float(os.environ.get("TB_REF_BUDGET_S", "200"))
    if spec["kind"] == "qp":
        n, m, p = spec["n"], spec["m"], spec["p"]
        psqrt, q, g, h, a, b = qp_instance(n, m, p, rdt)
        f8 = lambda v: np.asarray(v, dtype=np.float64)
        prob = O.ProbQP(O.MatBuild(O.MatType.SymPack(n), f8(psqrt)), O.MatBuild(O.MatType.General(n, 1), f8(q)),
                        O.MatBuild(O.MatType.General(m, n), f8(g)), O.MatBuild(O.MatType.General(m, 1), f8(h)),
                        O.MatBuild(O.MatType.General(p, n), f8(a)), O.MatBuild(O.MatType.General(p, 1), f8(b)), 1e-12, p_is_sqrt=True)
        del g, a, psqrt
        per_iter, timed = W.time_iterations(O, prob, steps, warmup)
        return {"value": 1.0 / per_iter, "unit": "iterations/s", "cores": cores, "kind": "port", "full_size": True, "iterations_timed": timed,
                "s_per_iteration": per_iter, "stacked": None,
                "sample": "the full workload (f64; dspmv on the packed P^(1/2) + dgemv on G and A_eq like ProbQP, OpenBLAS via numpy/scipy instead of MKL)"}
    cone, n = spec["cone"], spec["n"]
    m = cone_rows(cone)
    scale = rdt(1.0 / math.sqrt(n))
    x0, s0, y0 = instance_vectors(cone, n, SEED)
    if cone[0] == "psd":
        k = cone[1]
        a = W.fill_f64(m, n, 0, SEED, scale, as_f32)
        b, c = W.rhs_from_rows(a, x0, s0, y0, rdt)
        prob = W.DenseProblem(O, a, b, c, oracle_blocks(cone))      # ProbSDP's operator tuple with p = 0 (sdp.rs:75-97: one MatOp symmat_f, sk x n)
        per_iter, timed = W.time_iterations(O, prob, steps, warmup)
        return {"value": 1.0 / per_iter, "unit": "iterations/s", "cores": cores, "kind": "port", "full_size": True, "iterations_timed": timed,
                "s_per_iteration": per_iter, "stacked": None,
                "sample": "the full workload (f64; one dgemv per op like ProbSDP's symmat_f, ConePSD::proj = LAPACK dsyevr + dsyr loop "
                          "on %d x %d, OpenBLAS via numpy/scipy instead of MKL)" % (k, k)}
    # ---- row-shardable workloads: SOCP (blocks of bdim rows) and LP (units of 64 rows)
    is_soc = cone[0] == "soc"
    bdim = cone[2] if is_soc else 64
    nblk = m // bdim

    def build(nb):
        if is_soc:
            return W.socp_blocks_problem(O, n, nb, bdim, SEED, scale, as_f32, x0, s0, y0, rdt)
        a = W.fill_f64(nb * bdim, n, 0, SEED, scale, as_f32)
        b, c = W.rhs_from_rows(a, x0, s0[:nb * bdim], y0[:nb * bdim], rdt)
        mb = O.MatBuild(O.MatType.General(nb * bdim, n))
        mb.array = a.reshape(-1, order="F")
        # ProbLP's shape (lp.rs:222-338): one MatOp G (rows x n) + ConeRPos(rows), no equalities
        return O.ProbLP(O.MatBuild(O.MatType.General(n, 1), c), mb, O.MatBuild(O.MatType.General(nb * bdim, 1), b),
                        O.MatBuild(O.MatType.General(0, n)), O.MatBuild(O.MatType.General(0, 1)))

    probe = None
    if sample_blocks is None:
        # probe on a slice that is well outside the last-level cache, then take the largest power-of-two fraction that fits
        pb = max(1, min(nblk, (512 << 20) // (bdim * n * 8)))
        t_pb, _ = W.time_iterations(O, build(pb), 2, 1)
        est_full = t_pb * nblk / pb
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 32 << 30
        sample_blocks = nblk
        while sample_blocks > 1 and ((warmup + steps + 2) * est_full * sample_blocks / nblk > budget_s
                                     or sample_blocks * bdim * n * 8 * 2.5 > 0.6 * avail):
            sample_blocks //= 2
        probe = {"blocks": pb, "s_per_iteration": t_pb, "estimated_full_s_per_iteration": est_full, "budget_s": budget_s}
    sample_blocks = max(1, min(nblk, sample_blocks))
    frac = sample_blocks / nblk
    ms = sample_blocks * bdim
    prob = build(sample_blocks)
    per_iter, timed = W.time_iterations(O, prob, steps, warmup)
    del prob
    full = sample_blocks == nblk
    rows_txt = ("all %d cone blocks = the whole %d x %d A, no extrapolation" % (nblk, m, n)) if (full and is_soc) else \
               ("all %d rows of A, no extrapolation" % m) if full else \
               ("first %d of %d %s (%d x %d rows of the same A), scaled x%g linearly in rows" % (sample_blocks, nblk, "SOC blocks" if is_soc else "64-row units", ms, n, frac))
    out = {"value": (1.0 / per_iter) * frac, "unit": "iterations/s", "cores": cores, "kind": "port", "full_size": full, "iterations_timed": timed,
           "s_per_iteration": per_iter, "row_fraction": frac, "probe": probe, "stacked": None,
           "sample": "%s; f64, %s, OpenBLAS via numpy instead of MKL; %d iterations timed after %d warm-up, %.4f s/iteration measured"
                     % (rows_txt, "per-block dgemv + dot like ProbSOCP (socp.rs:83-124)" if is_soc else "one dgemv per op like ProbLP", timed, warmup, per_iter)}
    if is_soc:
        # the same rows behind ONE MatOp: one dgemv per op / trans_op over the stacked A (what the device's fused route computes)
        a = W.fill_f64(ms, n, 0, SEED, scale, as_f32)
        b, c = W.rhs_from_rows(a, x0, s0[:ms], y0[:ms], rdt)
        st_steps = max(2, min(steps, 6))
        t_st, st_timed = W.time_iterations(O, W.DenseProblem(O, a, b, c, oracle_blocks(cone, ms)), st_steps, 1)
        del a
        out["stacked"] = {"value": (1.0 / t_st) * frac, "unit": "iterations/s", "s_per_iteration": t_st, "iterations_timed": st_timed,
                          "what": "same rows, same cone, ONE stacked A behind a single MatOp (one dgemv per op / trans_op) instead of ProbSOCP's per-block "
                                  "MatOps: the formulation the device's fused DenseOp + ProductCone route computes"}
    return out


# ------------------------------------------------------------------------------------------------------------
PARITY_TOL = {"f32": {1: 5e-6, 10: 5e-5, 100: 1e-4}, "f64": {1: 1e-12, 10: 1e-11, 100: 1e-9}}       # relative l_inf of x_hat, y_hat
PARITY_TOL_PSD = {"f32": {1: 5e-5, 10: 5e-4, 100: 1e-3}, "f64": {1: 1e-9, 10: 1e-8, 100: 1e-7}}     # C4: <= 1e-3 at K = 100


def rel_linf(got, want):
    want = np.asarray(want, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-300)) if want.size else 0.0


def parity_leg(args, spec, new_session, oracle_problem, dev_precond, rank, world, barrier):
    """SURVEY.md 8d "parity check run with the measurement": the instance that was just timed, iterated K in {1, 10[, 100]}
    times on the device(s) and by the f64 oracle (oracle/totsu_oracle.py, one dgemv per op over the same A, generated
    bit-identically on the host by oracle/native.c); x_hat / y_hat compared in relative l_inf, the residual triple
    (solver.rs:391) to 2 significant digits.  The oracle is the checker here, never the thing measured."""
    ks = [k for k in (1, 10, 100) if k <= max(1, args.parity_k)]
    s = new_session()
    assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0, device_precond=dev_precond) == "None"
    dev, done = {}, 0
    for k in ks:
        s.step(k - done)
        done = k
        xh, yh = s.xy()
        dev[k] = (xh, yh, (s.last.c0, s.last.c1, s.last.c2))
    s.close()
    out = None
    if rank == 0:
        t0 = time.perf_counter()
        try:
            O, W = _oracle_modules()
            cores = _blas_threads()
            prob = oracle_problem(O, W)
            snaps, trace = W.iterates(O, prob, ks)
        except MemoryError as e:
            prob = None
            out = {"skipped": str(e)}
        if prob is not None:
            psd = (not spec["kind"] == "qp") and spec["cone"][0] == "psd"
            tol = (PARITY_TOL_PSD if psd else PARITY_TOL)[args.dtype]
            rows, ok = [], True
            for k in ks:
                ex, ey = rel_linf(dev[k][0], snaps[k][0]), rel_linf(dev[k][1], snaps[k][1])
                want = trace[k - 1][1:]
                rt = (5e-3 if args.dtype == "f32" else 1e-8) * (20 if psd else 1)
                res_ok = all((not np.isfinite(w)) or abs(g - w) <= rt * max(abs(w), 1e-3) for g, w in zip(dev[k][2], want))
                good = ex <= tol[k] and ey <= tol[k] and res_ok
                ok = ok and good
                rows.append({"K": k, "rel_linf_x_hat": ex, "rel_linf_y_hat": ey, "tolerance": tol[k],
                             "residuals_device": list(dev[k][2]), "residuals_oracle": [float(v) for v in want], "residuals_agree_to_2_digits": res_ok, "pass": good})
            out = {"pass": ok, "checks": rows, "n_gpus": world, "oracle": "oracle/totsu_oracle.py, f64, one dgemv per op over the same A (inputs rounded to %s), "
                   "OpenBLAS on %d host threads; same b, c as the device" % (args.dtype, cores), "oracle_seconds": time.perf_counter() - t0}
    barrier()
    return out


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS.keys()))
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--repeats", type=int, default=5, help="timed windows of --steps iterations each; the median window is reported")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timeline", default=None, help="write the device timeline of 3 extra iterations to this path (diagnostics; see scripts/timeline_summary.py)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity leg (x_hat / y_hat vs the f64 oracle on the timed workload)")
    ap.add_argument("--parity-k", type=int, default=10, help="largest iteration count of the parity leg: K in {1, 10} (+ 100 when >= 100)")
    ap.add_argument("--cpu-sample-blocks", type=int, default=None)
    ap.add_argument("--pair-fusion", type=int, default=1, help="serve op/trans_op pairs with one read of A when the backend can")
    ap.add_argument("--speculation", type=int, default=1, help="compute the next pair's products in the current read of A when its inputs are already final (csrc/gemv.cu)")
    ap.add_argument("--vprog", type=int, default=1, help="run the small vector commands between streaming launches as one launch per batch (csrc/vprog.cu)")
    ap.add_argument("--pdl", type=int, default=1, help="programmatic dependent launch of the small dependent kernels (csrc/common.cuh launch_pdl)")
    ap.add_argument("--scalar-prefetch", type=int, default=1, help="reductions that followed a host-visible scalar last time ride on its round trip (csrc/prefetch.cu)")
    ap.add_argument("--shim-protocol", type=int, default=0,
                    help="1: drive the backend with the Rust binding's call protocol (tb_view_of_host per operand, tb_buf_retain / "
                         "tb_buf_release per split child; totsu_b200/host/linalg.hpp) instead of carried (handle, offset, length) views")
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
        # same --steps / --warmup as the repo arm; the whole workload whenever (warmup + steps) iterations of it fit the budget
        ref = cpu_reference_leg(spec, steps, warmup, args.cpu_sample_blocks, dtype=args.dtype)
        line = {"impl": "reference", "metric": "solver iterations/sec", "value": ref["value"], "unit": "iterations/s", "n_gpus": 0,
                "steps": steps, "warmup": warmup, "ms_per_step": 1e3 / ref["value"], "higher_is_better": True, "scaling": "strong",
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
    capi.check(L.tb_set_scalar_prefetch(1 if args.scalar_prefetch else 0))
    capi.check(L.tb_set_pdl(1 if args.pdl else 0))
    host.set_shim_protocol(bool(args.shim_protocol))
    config["host_layer"] = ("C++ mirror of the unmodified Solver issuing the Rust binding's call protocol (tb_view_of_host per operand, retain / release per split child)"
                            if args.shim_protocol else "C++ mirror of the unmodified Solver carrying (handle, offset, length) views")
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

        # The device arm is handed P itself: `ProbQP::new` runs `MatBuild::set_sqrt` (qp.rs:386 -> matbuild/mod.rs:220-241), whose
        # map_eig closure the backend serves with the GEMM-only square root (tb_sqrt_psd, csrc/eig.cu) - 0.19 s at n = 8192, once per
        # problem construction and outside Solver::solve.  The f64 oracle keeps the closed form of this diagonal P (dsyevr at
        # n = 8192 takes minutes on the host and yields the same matrix).
        p_packed = (qdata[0].astype(np.float64) ** 2).astype(dt)
        construct_s = []

        def new_session():
            t0_ = time.perf_counter()
            s_ = host.Session.qp(dt, p_packed, qdata[1], qdata[2], qdata[3], qdata[4], qdata[5], 1e-12, p_is_sqrt=False, col_major=True)
            construct_s.append(time.perf_counter() - t0_)
            config["problem_construction_s"] = float(np.median(construct_s))
            config["p_sqrt"] = "MatBuild::set_sqrt on the device (tb_sqrt_psd, GEMM-only Newton-Schulz), inside ProbQP::new"
            return s_

        def oracle_problem(O, W):
            f8 = lambda v: np.asarray(v, dtype=np.float64)
            return O.ProbQP(O.MatBuild(O.MatType.SymPack(qn), f8(qdata[0])), O.MatBuild(O.MatType.General(qn, 1), f8(qdata[1])),
                            O.MatBuild(O.MatType.General(qm, qn), f8(qdata[2])), O.MatBuild(O.MatType.General(qm, 1), f8(qdata[3])),
                            O.MatBuild(O.MatType.General(qp_, qn), f8(qdata[4])), O.MatBuild(O.MatType.General(qp_, 1), f8(qdata[5])), 1e-12, p_is_sqrt=True)
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

        def oracle_problem(O, W):
            st64 = np.asfortranarray(qp_stacked(qn, qm, qp_, dt)[0], dtype=np.float64)
            return W.DenseProblem(O, st64, b, c, [("rotsoc", qn + 2), ("rpos", qm), ("zero", qp_ + qpad)])
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

        def oracle_problem(O, W):
            # the whole A on the host in f64 (C3: 8.6 GB), bit-identical to what the ranks generated in HBM; b and c are the
            # device-computed ones (b = A x0 + s0, c = -A^T y0 through the backend above)
            try:
                import psutil
                avail = psutil.virtual_memory().available
            except Exception:
                avail = 32 << 30
            if m * n * 8 * 2.2 > 0.8 * avail:
                raise MemoryError("the f64 oracle needs %.1f GB for A (+ as much again in calc_precond); the host has %.1f GB available" % (m * n * 8 / 1e9, avail / 1e9))
            return W.DenseProblem(O, W.fill_f64(m, n, 0, SEED, scale, args.dtype == "f32"), b, c, oracle_blocks(cone))

    stream = torch.cuda.ExternalStream(capi.stream_ptr(), device=torch.device("cuda", local_rank))
    clocks = Clocks(local_rank)          # EVERY rank samples its own GPU, started before warm-up: nothing is forked inside a timed window

    def barrier():
        capi.check(L.tb_device_sync())
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    def counters():
        vl, vo = C.c_uint64(), C.c_uint64()
        capi.check(L.tb_vprog_stats(C.byref(vl), C.byref(vo)))
        sp = [C.c_uint64() for _ in range(3)]
        capi.check(L.tb_spec_stats(*[C.byref(v) for v in sp]))
        pf = [C.c_uint64() for _ in range(3)]
        capi.check(L.tb_scalar_prefetch_stats(*[C.byref(v) for v in pf]))
        return np.array([capi.launch_count(), vl.value, vo.value, sp[0].value, sp[1].value, sp[2].value, pf[0].value, pf[1].value, pf[2].value], dtype=np.float64)

    dev_precond = not qp_stock           # the fused route's calc_precond loops run as kernels (tb_recip_clamp); outside every timed window
    # ---- device-resident timing: R windows of EXACTLY K iterations, each bracketed by barrier + synchronize on both sides
    # and timed with CUDA events on the library's stream; value = K / median window (max over ranks per window)
    s = new_session()
    assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0, device_precond=dev_precond) == "None"
    if os.environ.get("BENCH_DEBUG"):
        for _ in range(min(warmup, 5)):
            s.step(1)
            print("dbg iter %d tau %.4e res %.4e %.4e %.4e" % (s.last.i, s.last.val_tau, s.last.c0, s.last.c1, s.last.c2), file=sys.stderr)
        print("dbg norms", s.norms(), file=sys.stderr)
    s.step(warmup)
    repeats = max(1, args.repeats)
    win_ms, win_host, win_wait, win_scalars, win_cnt = [], [], [], [], []
    hw_s, hw_n = C.c_double(), C.c_uint64()
    t_win0 = time.time()
    for _ in range(repeats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = counters()
        capi.check(L.tb_host_wait_stats(C.byref(hw_s), C.byref(hw_n)))        # reset
        barrier()
        e0.record(stream)
        th0 = time.perf_counter()
        s.step(steps)
        capi.check(L.tb_flush())             # nothing recorded or parked may be left behind the closing event
        th1 = time.perf_counter()
        e1.record(stream)
        barrier()
        capi.check(L.tb_host_wait_stats(C.byref(hw_s), C.byref(hw_n)))
        win_ms.append(e0.elapsed_time(e1)); win_host.append(th1 - th0); win_wait.append(hw_s.value); win_scalars.append(hw_n.value)
        win_cnt.append(counters() - c0)
    t_win1 = time.time()
    last = s.last
    # ---- the same region again with per-launch events around the streaming matvec (roofline numerator)
    capi.check(L.tb_prof_enable(1))
    prof_iters = min(steps, 20)
    pf0 = capi.pairs_fused()
    s.step(prof_iters)
    pairs_per_iter = (capi.pairs_fused() - pf0) / prof_iters
    v_l, v_ms, v_b = (C.c_uint64 * 9)(), (C.c_double * 9)(), (C.c_double * 9)()
    capi.check(L.tb_prof_read_variants(v_l, v_ms, v_b))
    capi.check(L.tb_prof_enable(0))
    if args.timeline:
        # real device timeline of 3 iterations (csrc/context.cu tb_timeline_*): one line per launch, completion time on the
        # device and issue time on the host; written by every rank next to each other (PATH.rank<r>)
        capi.check(L.tb_timeline_begin(8192))
        s.step(3)
        capi.check(L.tb_flush())
        need = C.c_size_t()
        capi.check(L.tb_timeline_dump(None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value + 16)
        capi.check(L.tb_timeline_dump(buf, len(buf), C.byref(need)))
        with open(args.timeline + (".rank%d" % rank if world > 1 else ""), "w") as f:
            f.write(buf.value.decode())
    s.close()

    # ---- end to end: one whole Solver::solve through the public API with host buffers (work, c, b - and for the QP
    # front-end the matrices - in host memory; scalars cross the boundary every iteration; the solution is read back).
    # Headline: calc_precond exactly as the unmodified solver does it (host loops over get_mut(), solver.rs:501-506);
    # `with_device_precond` = the same solve with those two loops as kernels.
    def e2e_solve(device_precond):
        barrier()
        ss = new_session()
        barrier()                            # session construction differs per rank: start the end-to-end clock together
        t0 = time.perf_counter()
        st_ = ss.begin(max_iter=steps, eps_acc=0.0, eps_inf=0.0, device_precond=device_precond)
        assert st_ == "None"
        t1 = time.perf_counter()
        st_, _ = ss.run()
        t2 = time.perf_counter()
        ss.end()
        ss.solution()
        capi.check(L.tb_device_sync())
        t3 = time.perf_counter()
        ss.close()
        return t3 - t0, {"begin": t1 - t0, "iterate": t2 - t1, "end_and_readback": t3 - t2}, st_

    t_e2e, e2e_break, st = e2e_solve(False)
    t_e2e_dp = e2e_solve(True)[0] if dev_precond else t_e2e
    clk = clocks.stop(t_win0, t_win1)
    worklen = 4 * (n + 2 * m + 1) + 2 * (n + m + 1)
    h2d = (worklen * esize + matrix_h2d) / steps + 3 * esize          # work (+ wrapped matrices) upload amortised + tau/kappa/unit scalars per iteration
    d2h = (n + m) * esize / steps + 6 * esize          # solution readback amortised + tau, kappa, g_x, g_y, |p|, |d|

    # ---- max over ranks (per window), then the median window
    win = np.array(win_ms, dtype=np.float64)
    tail = np.array([t_e2e, t_e2e_dp, clk["sm_mhz"] or 0.0], dtype=np.float64)
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor(np.concatenate([win, tail[:2], [-tail[2]]]), dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tt = tt.cpu().numpy()
        win, t_e2e, t_e2e_dp = tt[:repeats], float(tt[repeats]), float(tt[repeats + 1])
        clk["sm_mhz_min_over_ranks"] = float(-tt[repeats + 2])
    med = int(np.argsort(win)[len(win) // 2])
    ms_total = float(win[med])
    value = steps / (ms_total * 1e-3)
    e2e_value = steps / t_e2e
    cnt = win_cnt[med]
    peak, peak_src = measured_peaks()
    roof = None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    served = cnt[4] > 0
    if os.path.exists(tpath) and args.dtype == "f32" and world == 1:
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this workload
        tj = json.load(open(tpath))
        wants = (["stream_kernel<float, 1, 1>", "stream_kernel<float, 2, 2>"] if served else ["stream_kernel<float, 1, 1>"]) if pairs_per_iter else ["stream_kernel<float, "]
        vals = [v["dram_bytes_per_launch"] for k, v in tj.items() if k.startswith(args.workload + "|") and any(w in k for w in wants)]
        if vals:
            traffic = sum(vals) / len(vals)
    nl = sum(v_l)
    if nl:
        # ALGORITHMIC bytes per launch: one read of this rank's dense block per op / trans_op served.  With
        # the lazy pairing a launch serves an op AND a trans_op from ONE read of A, so the algorithmic figure is twice the
        # bytes actually streamed and `frac` may exceed 1; `streamed_*` is the un-doubled DRAM-side figure.  (QP: the
        # launches timed are the transform_ge ones on G and A_eq; the packed P^(1/2) goes through spmv_kernel.)
        kms, kbytes = sum(v_ms), sum(v_b)
        avg_ms = kms / nl
        ge_elems = (qm * qn + qp_ * qn) if qp_stock else dense_elems if qp_fused else m_loc * n
        alg_per_launch = 6.0 * prof_iters * ge_elems * esize / nl
        ach = alg_per_launch / (avg_ms * 1e-3) / 1e9
        streamed = (kbytes / nl) / (avg_ms * 1e-3) / 1e9
        names = {3: "stream_kernel<1,0> (op)", 1: "stream_kernel<0,1> (trans_op)", 4: "stream_kernel<1,1> (op + trans_op pair)",
                 8: "stream_kernel<2,2> (pair + the speculated criteria_conv pair)"}
        variants = {}
        for vi in range(9):
            if v_l[vi]:
                vms = v_ms[vi] / v_l[vi]
                vgb = (v_b[vi] / v_l[vi]) / (vms * 1e-3) / 1e9
                variants[names.get(vi, "stream_kernel<%d,%d>" % (vi // 3, vi % 3))] = {
                    "launches_timed": int(v_l[vi]), "avg_launch_ms": vms, "products_served_per_launch": vi // 3 + vi % 3,
                    "streamed_gbs": vgb, "streamed_frac": vgb / peak}
        roof = {"bound": "hbm", "kernel": "stream_kernel (TMA bulk-copy matvec)", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full capture)" if traffic else None, "peak_source": peak_src, "launches_timed": int(nl), "avg_launch_ms": avg_ms,
                "algorithmic_bytes_per_launch": alg_per_launch, "matvecs_per_launch": 6.0 * prof_iters * (2 if qp_stock else 1) / nl,
                "streamed_bytes_per_launch": kbytes / nl, "streamed_gbs": streamed, "streamed_frac": streamed / peak,
                "per_variant": variants,
                "note": ("op/trans_op pairs share one read of A (lazy pairing behind tb_denseop_apply) and, with speculative pairing, the "
                         "criteria_conv pair rides on the preceding pass: achieved/frac use the un-fused algorithmic bytes "
                         "(6 reads of A per iteration, SURVEY 8d) and exceed 1 by construction; streamed_frac is bytes actually read / peak") if pairs_per_iter else None}
    abytes_iter = 6.0 * dense_elems * esize
    line = {"metric": "solver iterations/sec", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": config,
            "windows": {"repeats": repeats, "ms_per_window_max_over_ranks": [float(v) for v in win], "reported": "median",
                        "value_min": steps / (float(win.max()) * 1e-3), "value_max": steps / (float(win.min()) * 1e-3)},
            "e2e": {"value": e2e_value, "unit": "iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "Solver::solve (begin + %d iterations + end) through the host layer, work/c/b%s in host memory; calc_precond's host loops as in solver.rs:501-506"
                            % (steps, " and the matrices" if qp_stock else ""),
                    "breakdown_s": e2e_break, "with_device_precond": steps / t_e2e_dp},
            "gpu_launches": int(cnt[0]), "roofline": roof,
            "hbm_frac_whole_iteration": abytes_iter * value / (world * peak * 1e9),
            "algorithmic_bytes_per_iteration": abytes_iter, "pair_fusion": bool(args.pair_fusion), "pairs_fused_per_iteration": pairs_per_iter,
            "speculative_pairing": {"enabled": bool(args.speculation), "passes_with_speculation_per_iteration": cnt[3] / steps,
                                    "pairs_served_without_reading_A_per_iteration": cnt[4] / steps,
                                    "dropped": int(cnt[5]),
                                    "note": "the criteria_conv pair's products are computed during the preceding pass over A (its inputs are already final): "
                                            "2 reads of A per iteration instead of 3, bit-identical results"},
            "vector_programs": {"enabled": bool(args.vprog), "launches_per_iteration": cnt[1] / steps,
                                "micro_ops_per_iteration": cnt[2] / steps,
                                "note": "small vector commands recorded into one cluster launch per batch (csrc/vprog.cu); each program counts as one of gpu_launches"},
            "scalar_prefetch": {"enabled": bool(args.scalar_prefetch), "prefetch_batches_per_iteration": cnt[6] / steps, "scalars_served_without_a_round_trip_per_iteration": cnt[7] / steps,
                                "dropped": int(cnt[8]),
                                "note": "g_x, g_y and |d| of criteria_conv (solver.rs:599-608) are computed behind the kappa / |p| round trips - as extra micro-ops of the "
                                        "program that posts kappa / |p| (no launch of their own; a separate small kernel when a vector is too long for the cluster "
                                        "executor) - and served from the mapped host box: 6 host round trips per iteration become 3"},
            "programmatic_dependent_launch": bool(args.pdl),
            "host": {"loop_s": win_host[med], "waiting_for_device_s": win_wait[med], "host_visible_scalars_per_iteration": win_scalars[med] / steps,
                     "note": "host time of the median window's loop and the part of it spent spinning on device results: the rest is issuing launches"},
            "clocks": clk, "last_residuals": [last.c0, last.c1, last.c2], "status_e2e": st}
    # ---- parity of the workload just timed against the f64 oracle, every rank runs the device side
    if not args.no_parity:
        line["parity"] = parity_leg(args, spec, new_session, oracle_problem, dev_precond, rank, world, barrier)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_leg(spec, 6, 2, args.cpu_sample_blocks if args.cpu_sample_blocks else (64 if spec.get("cone", ("",))[0] == "soc" else None),
                                                     dtype=args.dtype, budget_s=25.0)
        print(json.dumps(line))
    if abuf is not None:
        abuf.release()
    if world > 1:
        capi.check(L.tb_dist_finalize())
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
