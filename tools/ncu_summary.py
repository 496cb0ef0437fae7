"""Summarises an ncu launch list (csv) and a `--set full` capture into profiles/<tag>_*.{md,json,csv}.

    python tools/ncu_summary.py <tag> <launches.csv> <report.ncu-rep | raw.csv> [workload=nstar] [n=1000000] [note ...]

The capture is either the report itself or its `ncu -i report --page raw --csv` export (a report of the full library
embeds its 60 MB of cubins and does not fit the 64 MiB that travel back from the GPU box, so the export is made there).
<tag>_traffic.json is keyed by bench.py's workload and kernel names and carries the md5 of the device-code sources (csrc/ minus the
host-only translation units, `_lib.kernel_source_hash`) the captured library was built from: bench.py reports `roofline.traffic`
from it only when its own sources hash to the same.
"""
import hashlib
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tag, launches, rep = sys.argv[1:4]
    rest = sys.argv[4:]
    workload = next((a.split("=", 1)[1] for a in rest if a.startswith("workload=")), "nstar")
    n_per_gpu = int(next((a.split("=", 1)[1] for a in rest if a.startswith("n=")), "1000000"))
    notes = [a for a in rest if not a.startswith(("workload=", "n="))]
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    if os.path.abspath(launches) != os.path.abspath(os.path.join(out_dir, f"{tag}_launches.csv")):
        shutil.copy(launches, os.path.join(out_dir, f"{tag}_launches.csv"))
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, rows = r, rows[i + 1:]
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows:
        try:
            agg[r[ki].split("(")[0].strip()].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    sweep = [k for k in agg if "project" not in k and "bspline" not in k and "prep" not in k]
    tsweep = sum(sum(agg[k]) for k in sweep)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h = rr[0]

    def col(r, n):
        return r[h.index(n)] if n in h else "nan"
    L = [f"# {tag}: ncu summary (B200, sm_100a)", "",
         "Commands: `tools/evidence.sh` (each ncu pass only after the same `python bench.py --workload " + workload +
         " --steps 3 --warmup 3 --no-cpu-baseline --no-extras` had exited 0 without ncu):", "",
         "    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file launches.csv <bench command>",
         "    ncu --set full --clock-control none -k regex:'<the sweep's kernels>' -s <skip> -c 12 -o full <bench command>; ncu -i full.ncu-rep --page raw --csv",
         ""] + notes + ["",
         "## Launch list: share of device time (per-launch times are cold-cache and serialised: compare shares)", "",
         "| kernel | launches | avg us | share of all | share of the sweep |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        sh = f"{sum(v) / tsweep:.3f}" if k in sweep else "create-time"
        L.append(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / tot:.3f} | {sh} |")
    L += ["", "## `--set full` capture (first launch of each kernel)", "",
          "| kernel | duration us | dram read MB | dram write MB | dram % of peak | fp64 pipe % | dmma % | issue active % | achieved occupancy % | regs | warp instructions |",
          "|---|---|---|---|---|---|---|---|---|---|---|"]
    seen, traffic = set(), {}
    for r in rr[2:]:
        k = col(r, "Kernel Name").split("(")[0].strip()
        if k in seen:
            continue
        seen.add(k)
        rd, wr = float(col(r, "dram__bytes_read.sum")), float(col(r, "dram__bytes_write.sum"))
        unit = rr[1][h.index("dram__bytes_read.sum")]
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1e6)
        short = k.replace("void ", "").replace("bf::", "").split("<")[0]
        short = {"stats_kernel_tma": "stats_kernels", "ragged_stats_kernel": "ragged_stats_kernel"}.get(short, short)
        traffic[short] = traffic.get(short, 0.0) + (rd + wr) * scale
        tu = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(rr[1][h.index("gpu__time_duration.sum")], 1.0)
        L.append(f"| `{k}` | {float(col(r, 'gpu__time_duration.sum')) * tu:.1f} | {rd * scale / 1e6:.1f} | {wr * scale / 1e6:.1f} | "
                 f"{float(col(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | "
                 f"{float(col(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active')):.1f} | "
                 f"{float(col(r, 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active')):.1f} | "
                 f"{float(col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f} | "
                 f"{float(col(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} | "
                 f"{col(r, 'launch__registers_per_thread')} | {float(col(r, 'smsp__inst_executed.sum')) / 1e6:.1f} M |")
    open(os.path.join(out_dir, f"{tag}_ncu_summary.md"), "w").write("\n".join(L) + "\n")
    sys.path.insert(0, ROOT)
    from bayesfmmm_b200 import _lib
    path = os.path.join(out_dir, f"{tag}_traffic.json")
    doc = json.load(open(path)) if os.path.exists(path) else {}
    doc["unit"] = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, first launch of each kernel)"
    doc[workload] = {"n_per_gpu": n_per_gpu, "source_md5": _lib.kernel_source_hash(), "kernels": traffic}
    json.dump(doc, open(path, "w"), indent=1)
    print("\n".join(L[-8:]))


if __name__ == "__main__":
    main()
