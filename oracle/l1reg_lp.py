"""Config C1 (TEST INFRASTRUCTURE ONLY): the instance of the reference's `l1reg_lp` example, restated.

Follows /root/reference/examples/l1reg_lp/src/main.rs:
  * :50      `Xoshiro256StarStar::seed_from_u64(0)`  - crate rand_xoshiro 0.6.0 (un-vendored; pinned in
             examples/Cargo.lock).  Published algorithm: `seed_from_u64` runs SplitMix64 from the seed and takes four
             consecutive outputs as the state s[0..4]; `next_u64` is xoshiro256** (Blackman & Vigna):
             result = rotl(s1 * 5, 7) * 9, then the xoshiro state update with t = s1 << 17 and rotl(s3, 45).
  * :53      `rng.gen::<f64>()` - crate rand 0.8.5, `Standard` for f64: (next_u64() >> 11) * 2^-53, i.e. [0, 1) with 53 bits.
  * :52-53   x = 2 x l matrix filled column-major by `MatBuild::by_fn` (totsu/src/matbuild/mod.rs:68-78: for c { for r }).
  * :54-59   y_smp = cos(5 x[0,smp]) * cos(7 x[1,smp]).
  * :65-106  LP data: n = 3l+1 (z, alpha, beta, bias), m = 4l, p = 0, lambda = 0.2, gaussian kernel sigma^2 = 1/8 (:16-29).
  * :111-116 solved with eps_acc = 1e-3 through ProbLP.

Pinned against the reference's own output of this program: examples/l1reg_lp/plot.svg (committed by the reference) draws
the 20 sample points (x0, y, x1) as circles - radius 5 where |alpha_i| > 0.001, else 2 (main.rs:190-200) - and the fitted
surface as 40 polylines (main.rs:152-188); tests/test_c1_l1reg_lp.py checks this module + the oracle's solve against the
pixel coordinates extracted into tests/golden/l1reg_lp_plot.json.  The generator is additionally pinned to the xoshiro256**
known-answer vector of its authors' reference implementation (state 1,2,3,4 -> 11520, 0, 1509978240, ...).
"""
from __future__ import annotations

import math

import numpy as np

M64 = (1 << 64) - 1


def _rotl(x: int, k: int) -> int:
    return ((x << k) | (x >> (64 - k))) & M64


class SplitMix64:
    """rand_xoshiro 0.6.0 `SplitMix64` (seed_from_u64: state = seed)."""

    def __init__(self, seed: int):
        self.x = seed & M64

    def next_u64(self) -> int:
        self.x = (self.x + 0x9E3779B97F4A7C15) & M64
        z = self.x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        return z ^ (z >> 31)


class Xoshiro256StarStar:
    def __init__(self, s):
        self.s = [v & M64 for v in s]

    @classmethod
    def seed_from_u64(cls, seed: int) -> "Xoshiro256StarStar":
        sm = SplitMix64(seed)
        return cls([sm.next_u64() for _ in range(4)])      # from_rng fills the 32 seed bytes little-endian, 8 at a time

    def next_u64(self) -> int:
        s = self.s
        result = (_rotl((s[1] * 5) & M64, 7) * 9) & M64
        t = (s[1] << 17) & M64
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = _rotl(s[3], 45)
        return result

    def gen_f64(self) -> float:
        return (self.next_u64() >> 11) * (1.0 / (1 << 53))


def kernel(x: np.ndarray, ci: int, xj: np.ndarray, cj: int) -> float:       # main.rs:16-29
    sigma_sq = 1.0 / 8.0
    norm_sq = 0.0
    for r in range(x.shape[0]):
        d = x[r, ci] - xj[r, cj]
        norm_sq += d * d
    return math.exp(-norm_sq / sigma_sq)


def instance(l: int = 20, seed: int = 0, lam: float = 0.2):
    """Returns (x [2, l], y [l], vec_c [n], mat_g [m, n], vec_h [m]) exactly as main.rs:50-106 builds them."""
    rng = Xoshiro256StarStar.seed_from_u64(seed)
    x = np.zeros((2, l))
    for c in range(l):                 # by_fn: column-major order of calls (matbuild/mod.rs:73-77)
        for r in range(2):
            x[r, c] = rng.gen_f64()
    y = np.array([math.cos(5.0 * x[0, s]) * math.cos(7.0 * x[1, s]) for s in range(l)])
    n, m = l * 3 + 1, l * 4
    vec_c = np.zeros(n)
    for i in range(l):
        vec_c[i] = 1.0
        vec_c[l * 2 + i] = lam
    g = np.zeros((m, n))
    for i in range(l):
        g[i, i] = -1.0
        g[l + i, i] = -1.0
        g[l * 2 + i, l + i] = 1.0
        g[l * 3 + i, l + i] = -1.0
        g[l * 2 + i, l * 2 + i] = -1.0
        g[l * 3 + i, l * 2 + i] = -1.0
        g[i, l * 3] = 1.0
        g[l + i, l * 3] = -1.0
    for r in range(l):
        for c in range(l):
            k = kernel(x, r, x, c)
            g[r, l + c] = k
            g[l + r, l + c] = -k
    h = np.zeros(m)
    for i in range(l):
        h[i] = y[i]
        h[l + i] = -y[i]
    return x, y, vec_c, g, h


def wx(x: np.ndarray, alpha: np.ndarray, xi: np.ndarray) -> float:           # main.rs:32-42
    return float(sum(alpha[i] * kernel(x, i, xi, 0) for i in range(alpha.size)))
