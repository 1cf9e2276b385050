#!/usr/bin/env python
"""Regenerates tests/golden/cortex_m_lp_trace.txt from the reference checkout (only available in the build container:
/root/reference does not exist on the GPU box, which is why the excerpt is committed).

The reference is Rust and cannot be built in this image, so the only outputs "of the reference itself" are the ones it
ships: the QEMU log of examples/nostd_cortex-m (a full residual trace of the solver loop, FloatGeneric<f64>).  The known
answers of totsu/tests/*.rs, totsu_core/tests/solver.rs and the unit vectors of matop.rs / cone_psd.rs / matbuild are small
enough to be written out inside tests/test_oracle_golden.py next to the file:line they come from."""
import os
import sys

SRC = "/root/reference/examples/nostd_cortex-m/log_qemu.txt"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cortex_m_lp_trace.txt")
HEADER = ("# Known-answer trace of the reference (examples/nostd_cortex-m/log_qemu.txt:7-25) for the LP at\n"
          "# examples/nostd_cortex-m/src/main.rs:64-89 (FloatGeneric<f64>, max_iter=100000, log_period=10).\n")


SVG = "/root/reference/examples/l1reg_lp/plot.svg"
SVG_DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "l1reg_lp_plot.json")


def l1reg_plot():
    """examples/l1reg_lp/plot.svg is the reference's own output of config C1 (main.rs:122-203): 20 circles = the sample
    points (x0, y, x1) in projected pixel coordinates, radius 5 where |alpha_i| > 0.001 else 2 (main.rs:190-200), and 40
    green polylines of 20 vertices = the fitted surface wx + bias on a 20 x 20 grid (main.rs:165-188: first the 20 lines of
    constant x0, then the 20 lines of constant x1)."""
    import json
    import re
    svg = open(SVG).read()
    circles = [[int(a), int(b), int(r)] for a, b, r in re.findall(r'<circle cx="(-?\d+)" cy="(-?\d+)" r="(\d+)"', svg)]
    lines = []
    for attrs, pts in re.findall(r'<polyline ([^>]*?)points="([^"]*)"', svg):
        if "#007F00" not in attrs:
            continue
        lines.append([[int(v) for v in p.split(",")] for p in pts.split()])
    assert len(circles) == 20 and len(lines) == 40 and all(len(l) == 20 for l in lines)
    json.dump({"source": "examples/l1reg_lp/plot.svg (reference output of examples/l1reg_lp/src/main.rs)",
               "circles_cx_cy_r": circles, "surface_polylines": lines}, open(SVG_DST, "w"))
    print("wrote", SVG_DST)


def main():
    if os.path.exists(SVG):
        l1reg_plot()
    if not os.path.exists(SRC):
        sys.exit("reference checkout not found at /root/reference: nothing to regenerate")
    lines = open(SRC).read().splitlines()
    out = []
    for ln in lines[6:25]:                        # lines 7-25, 1-based
        if ln.startswith("[DEBUG] "):
            out.append(ln[len("[DEBUG] "):])      # the residual lines printed by solver.rs:391
        elif ln.startswith("solve ->"):
            out.append("# " + ln)                 # the solution the example prints
    body = "\n".join(out) + "\n"
    new = HEADER + body
    old = open(DST).read() if os.path.exists(DST) else None
    if old is not None and old != new:
        print("fixture differs from the reference log: rewriting", file=sys.stderr)
    open(DST, "w").write(new)
    print("wrote", DST)


if __name__ == "__main__":
    main()
