#!/usr/bin/env python
"""TB_PF_DEBUG=1 python scripts/pf_debug.py: prints the scalar-prefetch decisions of iterations 12-13 of a streaming-size SOCP."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from helpers import capi  # noqa: E402
from totsu_b200 import host  # noqa: E402

capi.init(0)
blocks, n = [(H.SOC, 64)] * 32 + [(H.RPOS, 512)], 1024
m = sum(l for _, l in blocks)
a, b, c = H.make_instance(m, n, blocks, seed=13, dtype=np.float32)
abuf, av = H.device_matrix(a)
s = host.Session.dense(np.float32, av, m, n, c, b, blocks, fused_op=True, fused_cone=True)
assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0, device_precond=True) == "None"
for it in range(14):
    print("[pf] ---- iteration %d" % it, file=sys.stderr, flush=True)
    s.step(1)
s.close()
abuf.release()
