#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches_c3_bench
    python scripts/summarize_ncu.py full gpurun_out/prof_stream.ncu-rep profiles/r01_stream_kernel_full [workload]

`launches`: copies the per-launch csv and writes <out>_summary.md (per-kernel count, total device time, share).
`full`: dumps the metrics the roofline needs (duration, DRAM bytes, DRAM %, occupancy, registers) per captured
launch into <out>.csv / <out>.md and records DRAM traffic per launch in profiles/traffic.json for bench.py.
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__cycles_active.avg", "launch__shared_mem_per_block_dynamic"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def launches(src, out):
    rows = list(csv.reader(open(src, errors="replace")))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    shutil.copyfile(src, out + ".csv")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) < len(rows[h]):
            continue
        try:
            ns = float(r[-1].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[4].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    with open(out + "_summary.md", "w") as f:
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.2f%% |\n" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
    print(open(out + "_summary.md").read())


def full(rep, out, workload=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [w for w in WANT if w in idx]
    recs = []
    for r in rows[2:]:
        recs.append({c: (r[idx[c]], units[idx[c]]) for c in cols})
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for rec in recs:
            w.writerow([rec[c][0] for c in cols])
    traffic = {}
    with open(out + ".md", "w") as f:
        for rec in recs:
            name = rec["Kernel Name"][0].split("(")[0]
            rd = float(rec["dram__bytes_read.sum"][0]) * UNIT[rec["dram__bytes_read.sum"][1]]
            wr = float(rec["dram__bytes_write.sum"][0]) * UNIT[rec["dram__bytes_write.sum"][1]]
            traffic.setdefault(name, []).append(rd + wr)
            f.write("### `%s`\n\n" % name)
            for c in cols[1:]:
                f.write("- %s = %s %s\n" % (c, rec[c][0], rec[c][1]))
            f.write("- DRAM traffic (read + write) = %.6e bytes\n\n" % (rd + wr))
    print(open(out + ".md").read())
    tj = os.path.join(os.path.dirname(out), "traffic.json")
    d = json.load(open(tj)) if os.path.exists(tj) else {}
    for k, v in traffic.items():
        d["%s|%s" % (workload or "default", k)] = {"dram_bytes_per_launch": sum(v) / len(v), "captures": len(v), "source": os.path.basename(out) + ".csv"}
    json.dump(d, open(tj, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:])
