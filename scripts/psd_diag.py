import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import helpers as H
from totsu_b200 import capi
import totsu_oracle as O
capi.init(0); L = capi.lib()
def proj(x, dt):
    k = int((np.sqrt(8 * x.size + 1) - 1) / 2 + 0.5)
    y = x.astype(dt).copy()
    xb, wb = capi.Buf(y), capi.Buf(np.zeros(2 * k * k + k, dtype=dt))
    capi.check(capi.fn("tb_proj_psd", dt)(xb.view(), 1e-12, wb.view()))
    xb.release(); wb.release()
    return y
for k in (7, 16, 33, 64, 128, 129):
    rng = np.random.default_rng(k)
    g = rng.standard_normal((k, k))
    x = H.svec((g + g.T) / 2)
    want = x.copy()
    O.ConePSD(np.zeros(O.ConePSD.query_worklen(x.size)), 1e-12).proj(False, want)
    for path in (0, 2):
        capi.check(L.tb_set_psd_path(path))
        for dt in (np.float32, np.float64):
            p1 = proj(x, dt); p2 = proj(p1, dt)
            print("k=%3d path=%d %s: err %.2e idem %.2e (max|want| %.2f)" % (k, path, np.dtype(dt).name, np.abs(p1 - want).max(), np.abs(p2.astype(np.float64) - p1).max(), np.abs(want).max()))
capi.check(L.tb_set_psd_path(0))
