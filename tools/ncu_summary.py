"""Summarises an ncu launch list (csv) and a `--set full` report into profiles/<tag>_*.{md,json,csv}.

    python tools/ncu_summary.py <tag> <launches.csv> <report.ncu-rep> [note ...]
"""
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
    notes = sys.argv[4:]
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
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
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h = rr[0]

    def col(r, n):
        return r[h.index(n)] if n in h else "nan"
    L = [f"# {tag}: ncu summary (B200, sm_100a)", "",
         "Commands (each after the same command line exited 0 without ncu):", "",
         "    ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline",
         "    ncu --set full --clock-control none --import-source on -k regex:'z_kernel|chi_kernel|ssr_kernel|stats_kernel' -s 20 -c 8 -o prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline",
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
        short = k.replace("void ", "").split("<")[0]
        traffic[short] = (rd + wr) * scale
        L.append(f"| `{k}` | {float(col(r, 'gpu__time_duration.sum')):.1f} | {rd * scale / 1e6:.1f} | {wr * scale / 1e6:.1f} | "
                 f"{float(col(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | "
                 f"{float(col(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active')):.1f} | "
                 f"{float(col(r, 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active')):.1f} | "
                 f"{float(col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f} | "
                 f"{float(col(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} | "
                 f"{col(r, 'launch__registers_per_thread')} | {float(col(r, 'smsp__inst_executed.sum')) / 1e6:.1f} M |")
    open(os.path.join(out_dir, f"{tag}_ncu_summary.md"), "w").write("\n".join(L) + "\n")
    json.dump({"unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)",
               "n_per_gpu": 1000000, "kernels": traffic}, open(os.path.join(out_dir, f"{tag}_traffic.json"), "w"), indent=1)
    print("\n".join(L[-8:]))


if __name__ == "__main__":
    main()
