#!/usr/bin/env python
"""Diagnostic for the tcgen05 symmetric GEMM (csrc/psd_tc.cu): one case per invocation so a hang costs one timeout.
    python scripts/tc_check.py <k> <splitk> [engine]   -> prints max relative error vs numpy f64 and exact-symmetry check"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from totsu_b200 import capi  # noqa: E402


def main():
    k = int(sys.argv[1]); splitk = int(sys.argv[2]); engine = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    capi.init(0)
    L = capi.lib()
    rng = np.random.default_rng(k + splitk)

    def sym():
        g = rng.standard_normal((k, k))
        return ((g + g.T) / 2).astype(np.float32)
    a, b, d = sym(), sym(), sym()
    ab, bb, db = (capi.Buf(np.asfortranarray(m).reshape(-1, order="F").copy(), mutable=False) for m in (a, b, d))
    c = np.zeros(k * k, dtype=np.float32)
    cb = capi.Buf(c)
    alpha, beta, gamma = 0.75, -0.5, 1.25
    capi.check(L.tb_symm_gemm_f32(k, alpha, ab.view(), bb.view(), beta, db.view(), gamma, cb.view(), engine, splitk))
    capi.check(L.tb_device_sync())
    cb.release()
    got = c.reshape(k, k, order="F").astype(np.float64)
    a64, b64, d64 = a.astype(np.float64), b.astype(np.float64), d.astype(np.float64)
    full = alpha * (a64 @ b64) + beta * d64 + gamma * np.eye(k)
    want = np.triu(full) + np.triu(full, 1).T           # upper triangle mirrored
    scale = np.abs(want).max()
    err = np.abs(got - want).max() / scale
    # reference error of a plain fp32 product, for context
    f32 = alpha * (a @ b).astype(np.float64) + beta * d64 + gamma * np.eye(k)
    err32 = np.abs(np.triu(f32) - np.triu(full)).max() / scale
    print("k=%d splitk=%d engine=%d: rel err %.3e (numpy fp32 matmul: %.3e), exactly symmetric: %s, swap=%s"
          % (k, splitk, engine, err, err32, bool(np.array_equal(got, got.T)), os.environ.get("TB_TC_DESC_SWAP", "0")))
    if err > 1e-3:
        bad = np.abs(got - want) / scale
        print("  first bad entries:", np.argwhere(bad > 1e-3)[:8].tolist(), "got", got[:2, :4].tolist(), "want", want[:2, :4].tolist())
    sys.exit(0 if err < 5e-6 else 1)


if __name__ == "__main__":
    main()
