"""Debug: dump the sign-iteration state for small k."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import O, capi, svec
capi.init(0)
np.set_printoptions(linewidth=200, precision=5)
for dt in (np.float32, np.float64):
    for k in (2, 3, 10):
        for path in (0, 1):
            capi.check(capi.lib().tb_set_psd_path(path))
            rng = np.random.default_rng(k)
            g = rng.standard_normal((k, k))
            x = svec((g + g.T) / 2).astype(dt)
            want = x.astype(np.float64).copy()
            O.ConePSD(np.zeros(O.ConePSD.query_worklen(x.size)), 1e-12).proj(False, want)
            w = np.zeros(2 * k * k + k, dtype=dt)
            xb, wb = capi.Buf(x), capi.Buf(w)
            capi.check(capi.fn("tb_proj_psd", dt)(xb.view(), 1e-12, wb.view()))
            xb.release(); wb.release()
            err = np.abs(x - want).max()
            print(dt.__name__, "k", k, "path", path, "err", err)
            if err > 1e-4 and k <= 3:
                print(" got ", x); print(" want", want)
                print(" X   ", w[:k * k]); print(" S   ", w[k * k + k:])
                X = w[:k * k].reshape(k, k).astype(np.float64)
                ew, ev = np.linalg.eigh(X)
                print(" sign", ((ev * np.sign(ew)) @ ev.T).ravel())
capi.check(capi.lib().tb_set_psd_path(0))
