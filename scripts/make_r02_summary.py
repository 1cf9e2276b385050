#!/usr/bin/env python
"""Build profiles/r02_summary.md and profiles/r02_scaling_summary.md from the bench JSON lines kept under profiles/."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(name):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return None
    for ln in open(path).read().splitlines()[::-1]:
        if ln.strip().startswith("{"):
            return json.loads(ln)
    return None


def row(label, name):
    d = load(name)
    if d is None:
        return None
    par = d.get("parity") or {}
    pk = ", ".join("K=%d: x %.1e y %.1e" % (c["K"], c["rel_linf_x_hat"], c["rel_linf_y_hat"]) for c in par.get("checks", []))
    roof = d.get("roofline") or {}
    return "| %s | %d | %.4f | %.1f | %.1f | %s | %s | `%s` |" % (
        label, d["steps"], d["ms_per_step"], d["value"], d["e2e"]["value"],
        ("%.3f" % roof["streamed_frac"]) if roof.get("streamed_frac") else "-", ("pass (%s)" % pk) if par.get("pass") else ("-" if not par else "FAIL"), name)


def main():
    out = ["# Round 2 - single-GPU bench lines (one B200, f32; `python bench.py ...`, median of 5 timed windows, CUDA events)", "",
           "| workload | steps | ms / iteration | it/s device-timed | it/s end to end (host buffers) | streaming kernel: bytes read / measured HBM peak | parity vs the f64 oracle | file |",
           "|---|---:|---:|---:|---:|---:|---|---|"]
    for label, name in (("C3 SOCP 1024 x SOC(64), A 65536 x 16384 (default line, driver's flags)", "r02_bench_c3_n1_final.json"),
                        ("C3, 100 steps", "r02_bench_c3_n1_100.json"),
                        ("C3, binding's call protocol (`--shim-protocol 1`; mid-round, before the launch-count work)", "r02_bench_c3_n1_shim_protocol.json"),
                        ("C2 QP n = 8192, m = 8192, p = 1024, fused route", "r02_bench_c2_fused_n1.json"),
                        ("C2, stock `ProbQP` route (`set_sqrt` on the device)", "r02_bench_c2_stock_n1.json"),
                        ("C4 SDP PSD(512), A 131328 x 1024", "r02_bench_c4_n1.json"),
                        ("C3 in f64 on the device", "r02_bench_c3_n1_f64.json")):
        r = row(label, name)
        if r:
            out.append(r)
    ref = load("r02_bench_ref_c3_full.json")
    if ref:
        cb = ref.get("cpu_baseline", {})
        out += ["", "Reference arm (`bench.py --impl reference`, the oracle port of the `totsu_f64lapack` path, whole C3 problem in f64, %d host threads, "
                "OpenBLAS): **%.3f it/s** (%.2f s / iteration, stock per-block `ProbSOCP` route); one stacked `dgemv` per op instead: %.2f it/s "
                "(`r02_bench_ref_c3_full.json`)." % (cb.get("cores", 0), ref["value"], ref["ms_per_step"] / 1e3, (cb.get("stacked") or {}).get("value", float("nan")))]
    open(os.path.join(P, "r02_summary.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))

    sc = ["# Round 2 - C3 strong scaling on one 8-GPU B200 box (NVLink 5 / NVSwitch), f32", "",
          "`torchrun --nproc-per-node N bench.py --gpus N --steps 100 --warmup 10` (and the driver's `--steps 20 --warmup 5`); value = median of 5 windows, max over ranks.", "",
          "## End of round 2 (one-launch cone pair, reductions on the cluster, prefetch riding in the trigger's program, ...: 15 launches per iteration at 1 GPU, 19 at N > 1)", "",
          "| N | run | ms / iteration | it/s | efficiency vs N = 1 on the same box | it/s end to end | streaming kernel ms / launch | file |", "|---:|---|---:|---:|---:|---:|---:|---|"]

    def block(files, base_name):
        base = load(base_name)
        b1 = base["value"] if base else None
        if base:
            sc.append("| 1 | 100 steps | %.4f | %.1f | 1.00 | %.1f | %.4f | `%s` |" % (base["ms_per_step"], base["value"], base["e2e"]["value"], base["roofline"]["avg_launch_ms"], base_name))
        for n, label, name in files:
            d = load(name)
            if d is None:
                continue
            sc.append("| %d | %s | %.4f | %.1f | %s | %.1f | %.4f | `%s` |" % (n, label, d["ms_per_step"], d["value"], ("%.2f" % (d["value"] / (n * b1))) if b1 else "-",
                                                                          d["e2e"]["value"], (d.get("roofline") or {}).get("avg_launch_ms", float("nan")), name))
    block([(4, "100 steps", "r02_bench_c3_n4_final.json"), (4, "driver's flags (20 steps, parity + CPU legs)", "r02_bench_c3_n4_driver_final.json"),
           (8, "100 steps", "r02_bench_c3_n8_final.json"), (8, "driver's flags (20 steps, parity + CPU legs)", "r02_bench_c3_n8_driver_final.json")], "r02_bench_c3_n1_same_box_as_n8_final.json")
    sc += ["", "2-GPU box:", "", "| N | run | ms / iteration | it/s | efficiency vs N = 1 on the same box | it/s end to end | streaming kernel ms / launch | file |", "|---:|---|---:|---:|---:|---:|---:|---|"]
    block([(2, "100 steps", "r02_bench_c3_n2_final.json"), (2, "driver's flags (20 steps, parity + CPU legs)", "r02_bench_c3_n2_driver_final.json")], "r02_bench_c3_n1_same_box_as_n2_final.json")
    sc += ["", "`r02_timeline_c3_n8_rank0.md` lays one 8-GPU iteration out launch by launch (379 us with the timeline's own event records, 310 us without): the two streaming passes",
           "are 2 x 0.090 ms; the two peer exchanges (push 24 us + wait 8 us each) are the largest remaining item, then the programs between them; the device waits for the host",
           "three times (~6 us each).", "",
           "## Mid-round (before the launch-count work; same code paths A/B'd on one box)", "",
           "| N | path | ms / iteration | it/s | efficiency vs N = 1 on the same box | it/s end to end | streaming kernel ms / launch | file |", "|---:|---|---:|---:|---:|---:|---:|---|"]
    base = load("r02_bench_c3_n1_same_box_as_n8.json")
    b1 = base["value"] if base else None
    if base:
        sc.append("| 1 | - | %.4f | %.1f | 1.00 | %.1f | %.4f | `r02_bench_c3_n1_same_box_as_n8.json` |" % (base["ms_per_step"], base["value"], base["e2e"]["value"], base["roofline"]["avg_launch_ms"]))
    for n in (2, 4, 8):
        for tag, label in (("_driver", "peer stores, driver's flags (20 steps)"), ("", "peer stores, speculated pair in the carrying exchange (default)"),
                           ("_nospecx", "peer stores, speculated pair exchanged on arrival (`TB_P2P_SPEC=0`)"), ("_nccl", "ncclAllGather / ncclAllReduce (`TB_P2P=0`)")):
            name = "r02_bench_c3_n%d%s.json" % (n, tag)
            d = load(name)
            if d is None:
                continue
            sc.append("| %d | %s | %.4f | %.1f | %s | %.1f | %.4f | `%s` |" % (n, label, d["ms_per_step"], d["value"], ("%.2f" % (d["value"] / (n * b1))) if b1 else "-",
                                                                          d["e2e"]["value"], (d.get("roofline") or {}).get("avg_launch_ms", float("nan")), name))
    sc += ["", "Round 1 (driver, `SCALE_r01.json`): 654.8 / 613.5 / 1637.1 / 2285.7 it/s at 1 / 2 / 4 / 8.",
           "The mid-round end-to-end figures of the 2-GPU box predate the device block pool (stalls of `begin` / `end`, see `r02_ab_pdl_prefetch.md`)."]
    open(os.path.join(P, "r02_scaling_summary.md"), "w").write("\n".join(sc) + "\n")
    print("\n".join(sc))


if __name__ == "__main__":
    main()
