#!/usr/bin/env python
"""Measured iterate errors of the PSD test instances (tests/test_solver_gpu.py SYN['sdp_like'], SYN['sdp_tc']) against the f64
oracle at K = 1, 10, 100, both routes, both precisions: the numbers the tolerances in test_iterates_match_oracle are set from."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import helpers as H  # noqa: E402
from helpers import capi, ZERO, RPOS, PSD  # noqa: E402
from totsu_b200 import host  # noqa: E402

SYN = {"sdp_like": ([(PSD, 36), (RPOS, 10), (ZERO, 3)], 20, 4), "sdp_tc": ([(PSD, 2080)], 40, 5), "sdp_k128": ([(PSD, 8256)], 64, 6)}


def main():
    capi.init(0)
    out = {}
    for name, (blocks, n, seed) in SYN.items():
        m = sum(l for _, l in blocks)
        for dt in (np.float64, np.float32):
            a, b, c = H.make_instance(m, n, blocks, seed=seed, dtype=dt)
            ks = [1, 10, 100]
            snaps, trace = H.oracle_iterates(a, b, c, blocks, ks)
            abuf, av = H.device_matrix(a)
            for fused in (False, True):
                s = host.Session.dense(dt, av, m, n, c, b, blocks, fused_op=fused, fused_cone=fused)
                assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0, device_precond=fused) == "None"
                done = 0
                for k in ks:
                    s.step(k - done); done = k
                    xh, yh = s.xy()
                    res = [abs(g - w) / max(abs(w), 1e-3) for g, w in zip((s.last.c0, s.last.c1, s.last.c2), trace[k - 1][1:]) if np.isfinite(w)]
                    out["%s|%s|%s|K=%d" % (name, np.dtype(dt).name, "fused" if fused else "stock", k)] = [H.rel_linf(xh, snaps[k][0]), H.rel_linf(yh, snaps[k][1]), max(res) if res else 0.0]
                s.close()
            abuf.release()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
