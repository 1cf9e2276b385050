"""Config C1 (BASELINE.json configs[0]): the reference's `l1reg_lp` example (examples/l1reg_lp/src/main.rs:45-123,
n = 61, m = 80, p = 0, eps_acc = 1e-3, Xoshiro256** seed 0).

CPU (`-m "not gpu"`): the restated instance (oracle/l1reg_lp.py) and the oracle's ProbLP solve are pinned against the
reference's OWN output of this program, examples/l1reg_lp/plot.svg (fixture tests/golden/l1reg_lp_plot.json, extracted by
tests/golden/make_fixtures.py): the 20 sample points, which of the alpha_i are non-zero (circle radius 5 vs 2,
main.rs:190-200) and the fitted surface (40 polylines, main.rs:152-188), all at the plot's 1-pixel resolution.
GPU (`-m gpu`): the same LP through the ProbLP front-end on the device (stock MatOp + ConeRPos route, f64 and f32) -
final answer vs the oracle <= 1e-3, the tolerance the reference's own tests use."""
import json
import os

import numpy as np
import pytest

import l1reg_lp as C1
import totsu_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "l1reg_lp_plot.json")
L = 20


def _oracle_solve(eps_acc=1e-3):
    x, y, c, g, h = C1.instance()
    n, m = c.size, h.size
    prob = O.ProbLP(O.MatBuild(O.MatType.General(n, 1), c.copy()), O.MatBuild(O.MatType.General(m, n), g.reshape(-1, order="F").copy()),
                    O.MatBuild(O.MatType.General(m, 1), h.copy()), O.MatBuild(O.MatType.General(0, n)), O.MatBuild(O.MatType.General(0, 1)))
    s = O.Solver()
    s.par.eps_acc = eps_acc
    sol = s.solve(prob.problem())
    return (x, y, c, g, h), np.array(sol[0]), np.array(sol[1])


def _projection(x, y, circles):
    """plotters' 3-D projection (main.rs:138-143: yaw 0.5, pitch 0.3, scale 0.8, no perspective) is affine in
    (x0, y, x1): fit pixel = P . [x0, y, x1, 1] on the 20 drawn sample points; returns P (2 x 4) and the worst residual."""
    pts = np.column_stack([x[0], y, x[1], np.ones(x.shape[1])])
    pm, worst = [], 0.0
    for k in range(2):
        sol = np.linalg.lstsq(pts, circles[:, k].astype(np.float64), rcond=None)[0]
        worst = max(worst, float(np.abs(pts @ sol - circles[:, k]).max()))
        pm.append(sol)
    return np.array(pm), worst


def test_xoshiro256starstar_known_answer():
    """Reference vector of the xoshiro256** authors' implementation (the one rand_xoshiro's own test holds)."""
    r = C1.Xoshiro256StarStar([1, 2, 3, 4])
    assert [r.next_u64() for _ in range(6)] == [11520, 0, 1509978240, 1215971899390074240, 1216172134540287360, 607988272756665600]


def test_instance_and_solution_match_the_reference_plot():
    gold = json.load(open(GOLDEN))
    circles = np.array(gold["circles_cx_cy_r"])
    (x, y, c, g, h), sol, _ = _oracle_solve()
    assert c.size == 61 and g.shape == (80, 61)                      # n = 61, m = 80
    # (1) sample points: 20 points x 2 pixel coordinates, 8 free parameters; integer rounding allows <= ~0.75 px
    pm, worst = _projection(x, y, circles)
    assert worst < 0.8, worst
    # ... and a wrong generator cannot fit: the next 40 draws of the same stream are nowhere near
    rng = C1.Xoshiro256StarStar.seed_from_u64(1)
    xb = np.array([[rng.gen_f64() for _ in range(2)] for _ in range(L)]).T
    yb = np.cos(5 * xb[0]) * np.cos(7 * xb[1])
    assert _projection(xb, yb, circles)[1] > 10.0
    # (2) support of alpha: radius 5 <=> |alpha_i| > 0.001 (main.rs:193)
    alpha, bias = sol[L:2 * L], sol[3 * L]
    assert [5 if abs(a) > 0.001 else 2 for a in alpha] == circles[:, 2].tolist()
    # (3) the fitted surface on the 20 x 20 grid, drawn as polylines (main.rs:165-188)
    grid = [f / (L - 1) for f in range(L)]
    surf = np.array([[C1.wx(x, alpha, np.array([[a], [b]])) + bias for b in grid] for a in grid])      # surf[i0, i1]
    lines = np.array(gold["surface_polylines"], dtype=np.float64)           # [40, 20, 2]
    worst = 0.0
    for i0 in range(L):          # lines of constant x0: vertices over x1
        pts = np.column_stack([np.full(L, grid[i0]), surf[i0, :], grid, np.ones(L)])
        worst = max(worst, float(np.abs(pts @ pm.T - lines[i0]).max()))
    for i1 in range(L):          # lines of constant x1: vertices over x0
        pts = np.column_stack([grid, surf[:, i1], np.full(L, grid[i1]), np.ones(L)])
        worst = max(worst, float(np.abs(pts @ pm.T - lines[L + i1]).max()))
    # 1 px = 1/73 in y (the fitted scale): the reference's solution and the oracle's (both stopped at eps_acc = 1e-3) agree
    # to the plot's resolution
    assert worst < 2.0, worst


@pytest.mark.gpu
@pytest.mark.parametrize("dtname", ["f64", "f32"])
def test_c1_on_device_matches_oracle(dtname):
    from totsu_b200 import capi, host
    capi.init(0)
    dt = np.float64 if dtname == "f64" else np.float32
    (x, y, c, g, h), want_x, want_y = _oracle_solve(eps_acc=1e-3)           # the example's own run (main.rs:112)
    s = host.Session.lp(dt, c, g, h)
    st, got_x, got_y = s.solve(max_iter=200_000, eps_acc=1e-3)
    s.close()
    assert st == "None"
    # final answer vs the oracle <= 1e-3 ("the tolerance the reference's own tests use"); in f64 the device
    # follows the oracle's path iteration for iteration and stops at the same one
    tol = 1e-9 if dt == np.float64 else 1e-3
    assert np.abs(got_x.astype(np.float64) - want_x).max() <= tol, np.abs(got_x - want_x).max()
    assert np.abs(got_y.astype(np.float64) - want_y).max() <= tol, np.abs(got_y - want_y).max()
    assert abs(float(c @ got_x.astype(np.float64)) - float(c @ want_x)) <= 1e-3
    # feasibility of the device answer: G x <= h up to the accuracy asked for
    assert (g @ got_x.astype(np.float64) - h).max() <= 5e-3
    # support of alpha as drawn by the reference (radius 5 <=> |alpha| > 0.001)
    gold = json.load(open(GOLDEN))
    r = [row[2] for row in gold["circles_cx_cy_r"]]
    a = got_x[L:2 * L]
    assert all((abs(ai) > 0.001) == (ri == 5) for ai, ri in zip(a, r) if abs(abs(ai) - 0.001) > 5e-4)
