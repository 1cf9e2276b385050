#!/usr/bin/env python
"""Lay out one solver iteration from a device timeline (bench.py --timeline PATH; csrc/context.cu tb_timeline_*).

    python scripts/timeline_summary.py gpurun_out/tl_c3.txt [out.md]

Every line of the dump is one kernel launch: the site that launched it (file:line of the library, or the number of micro-ops
of a vector program), the time its kernel COMPLETED on the device and the time the host issued it, both in microseconds since
the first launch.  `dt` = completion minus the previous completion = the kernel's own time plus whatever gap preceded it;
`host lead` = completion of the previous kernel minus the host's issue of this one: negative means the device sat idle
waiting for the host (a host round trip or a slow issue), positive means the launch was already queued."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_src = {}


def label(site):
    """file:line -> the kernel named at that launch site (looked up in the source), or the site itself."""
    m = re.match(r"(\w+\.cu):(\d+)$", site)
    if not m:
        return site
    f, ln = m.group(1), int(m.group(2))
    if f not in _src:
        try:
            _src[f] = open(os.path.join(ROOT, "totsu_b200", "csrc", f)).read().splitlines()
        except OSError:
            _src[f] = []
    lines = _src[f]
    for k in range(ln - 1, max(ln - 8, -1), -1):
        if k < len(lines):
            mm = re.search(r"launch_pdl\(\s*([\w:]+(?:<[^>]*>)?)", lines[k]) or re.search(r"([\w:]+(?:<[^<>]*>)?)\s*<<<", lines[k]) \
                or re.search(r"(launch_stream<[^>]*>)", lines[k]) or re.search(r"TB_NCCL\(g_nccl\.(\w+)", lines[k])
            if mm:
                return "%s (%s)" % (mm.group(1), site)
    return site


def main():
    path = sys.argv[1]
    rows = []
    for ln in open(path):
        f = ln.split()
        if len(f) == 4:
            rows.append((f[1], float(f[2]), float(f[3])))
    if not rows:
        sys.exit("empty timeline")
    # iterations are delimited by the first site repeating with the same period: take the middle third of the launches
    n = len(rows)
    per = n // 3
    lo, hi = per, 2 * per
    out = ["| # | launch site | completes at us | dt us (kernel + gap before it) | host issued at us | host lead us |", "|---:|---|---:|---:|---:|---:|"]
    t0 = rows[lo - 1][1]
    agg = {}
    for i in range(lo, hi):
        site, dev, host = rows[i]
        dt = dev - rows[i - 1][1]
        lead = rows[i - 1][1] - host
        lab = label(site)
        out.append("| %d | `%s` | %.1f | %.1f | %.1f | %.1f |" % (i - lo, lab, dev - t0, dt, host - t0, lead))
        a = agg.setdefault(lab.split(" (")[0], [0, 0.0, 0.0])
        a[0] += 1; a[1] += dt
        if lead < 0:
            a[2] += -lead
    total = rows[hi - 1][1] - t0
    head = ["# Device timeline of one iteration: %s" % os.path.basename(path), "",
            "%d launches, %.1f us (launches %d..%d of %d recorded over 3 iterations).  dt = completion - previous completion." % (hi - lo, total, lo, hi - 1, n), "",
            "| kernel | launches | sum of dt us | share | of which device idle waiting for the host us |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        head.append("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k, v[0], v[1], 100 * v[1] / total, v[2]))
    text = "\n".join(head + [""] + out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
