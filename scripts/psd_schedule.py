#!/usr/bin/env python
"""Error and time of the f32 ConePSD projection per sign-iteration schedule (TB_PSD_STEPS="n1,n2" must be set by the caller:
the library reads it once).  Prints one JSON line: max error vs numpy eigh (relative to max|X|) for random and hard spectra
at k = 96 / 512, and ms per projection at k = 512."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from totsu_b200 import capi
capi.init(0); L = capi.lib(); dt = np.float32
stream = torch.cuda.ExternalStream(capi.stream_ptr())

def svec(x):
    k = x.shape[0]
    out = np.empty(k * (k + 1) // 2)
    p = 0
    for c in range(k):
        out[p:p + c + 1] = x[:c + 1, c] * np.sqrt(2.0); out[p + c] = x[c, c]; p += c + 1
    return out

def smat(v, k):
    x = np.zeros((k, k)); p = 0
    for c in range(k):
        x[:c + 1, c] = v[p:p + c + 1] / np.sqrt(2.0); x[c, c] = v[p + c]; p += c + 1
    return np.triu(x) + np.triu(x, 1).T

def spectrum(k, lam, seed):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((k, k)))
    x = (q * lam) @ q.T
    return (x + x.T) / 2

def proj_err(xm):
    k = xm.shape[0]
    x32 = svec(xm).astype(dt)
    xm64 = smat(x32.astype(np.float64), k)
    w, v = np.linalg.eigh(xm64)
    want = (v * np.maximum(w, 0)) @ v.T
    xb, wb = capi.Buf(x32.copy()), capi.Buf(np.zeros(2 * k * k + k, dtype=dt))
    capi.check(capi.fn("tb_proj_psd", dt)(xb.view(), 1e-12, wb.view()))
    got = smat(xb.download().astype(np.float64), k)
    xb.release(); wb.release()
    return float(np.abs(got - want).max() / max(np.abs(xm64).max(), 1e-30))

out = {"schedule": os.environ.get("TB_PSD_STEPS", "10,8")}
rng = np.random.default_rng(3)
for k in (96, 512):
    g = rng.standard_normal((k, k))
    cases = {
        "random": (g + g.T) / 2,
        "logspaced": spectrum(k, np.concatenate([np.logspace(-12, 0, k // 2), -np.logspace(-12, 0, k // 2)]), 4),
        "lowrank": spectrum(k, np.concatenate([np.ones(5), np.zeros(k - 5)]), 4),
        "near_boundary": spectrum(k, np.concatenate([rng.uniform(0.5, 1, k // 2), rng.standard_normal(k // 2) * 1e-6]), 4),
        "clustered_small": spectrum(k, np.concatenate([np.ones(k // 2), rng.uniform(-1e-3, 1e-3, k // 2)]), 5),
        "neg_def": spectrum(k, -np.abs(rng.standard_normal(k)) - 0.1, 6),
    }
    for name, xm in cases.items():
        out["err_k%d_%s" % (k, name)] = proj_err(xm)
k = 512
g = rng.standard_normal((k, k))
x32 = svec((g + g.T) / 2).astype(dt)
xb, wb = capi.Buf(x32.copy()), capi.Buf(np.zeros(2 * k * k + k, dtype=dt))
f = capi.fn("tb_proj_psd", dt)
for _ in range(3):
    capi.check(f(xb.view(), 1e-12, wb.view()))
capi.check(L.tb_device_sync())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(20):
    capi.check(f(xb.view(), 1e-12, wb.view()))
capi.check(L.tb_flush()); e1.record(stream); capi.check(L.tb_device_sync()); torch.cuda.synchronize()
out["ms_per_projection_k512"] = e0.elapsed_time(e1) / 20
print(json.dumps(out))
