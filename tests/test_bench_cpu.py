"""CPU checks of the benchmark's own constructions (no GPU): the stacked dense form of ProbQP used by bench.py's fused C2
route is the operator / right-hand side the oracle's ProbQP builds (qp.rs:98-140, 196-215, 20-45), the packed svec order,
and the reference arm's JSON contract on a small workload."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import bench  # noqa: E402
import totsu_oracle as O  # noqa: E402


def test_qp_stacked_equals_oracle_probqp_operator():
    qn, qm, qp_ = 40, 28, 5
    dt = np.float64
    qdata = bench.qp_instance(qn, qm, qp_, dt)
    stacked, b, c, pad = bench.qp_stacked(qn, qm, qp_, dt, qdata)
    psqrt, q, g, h, a_eq, b_eq = qdata
    prob = O.ProbQP(O.MatBuild(O.MatType.SymPack(qn), psqrt.copy()), O.MatBuild(O.MatType.General(qn, 1), q.copy()),
                    O.MatBuild(O.MatType.General(qm, qn), g.copy()), O.MatBuild(O.MatType.General(qm, 1), h.copy()),
                    O.MatBuild(O.MatType.General(qp_, qn), a_eq.copy()), O.MatBuild(O.MatType.General(qp_, 1), b_eq.copy()), 1e-12, p_is_sqrt=True)
    op_c, op_a, op_b, cone, _ = prob.problem()
    m0, n = op_a.size()
    assert stacked.shape == (m0 + pad, n) and (m0 + pad) % 4 == 0
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(n), rng.standard_normal(m0)
    out = np.zeros(m0); op_a.op(1.0, x, 0.0, out)
    assert np.allclose(stacked[:m0] @ x, out, rtol=1e-13, atol=1e-13)
    out_t = np.zeros(n); op_a.trans_op(1.0, y, 0.0, out_t)
    assert np.allclose(stacked[:m0].T @ y, out_t, rtol=1e-13, atol=1e-13)
    assert not stacked[m0:].any() and not b[m0:].any()          # padded rows: 0 . x = 0 in the ConeZero block
    one = np.ones(1)
    bo = np.zeros(m0); op_b.op(1.0, one, 0.0, bo)
    co = np.zeros(n); op_c.op(1.0, one, 0.0, co)
    assert np.allclose(b[:m0], bo) and np.allclose(c, co)


def test_svec_matches_reference_packing():
    """cone_psd.rs:18 / matbuild: upper triangle by columns, off-diagonals scaled by sqrt(2)."""
    rng = np.random.default_rng(1)
    k = 7
    g = rng.standard_normal((k, k)); s = (g + g.T) / 2
    want = np.array([s[r, c] * (1.0 if r == c else np.sqrt(2.0)) for c in range(k) for r in range(c + 1)])
    assert np.allclose(bench.svec(s), want)


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "socp_small_128x64_A8192x4096",
                          "--steps", "3", "--warmup", "3", "--cpu-sample-blocks", "8"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "iterations/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["workload"].startswith("socp_small")


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    """The driver launches `bench.py --impl reference` like the GPU arm (torchrun for N > 1): rank 0 alone times the CPU
    path (with the BLAS pool restored: torchrun exports OMP_NUM_THREADS=1) and prints the line, the other ranks exit 0."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29431",
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "socp_small_128x64_A8192x4096",
           "--steps", "3", "--warmup", "3", "--cpu-sample-blocks", "4"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0
    assert d["cpu_baseline"]["cores"] >= 1 and (os.cpu_count() == 1 or d["cpu_baseline"]["cores"] > 1)
