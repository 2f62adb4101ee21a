#!/usr/bin/env python
"""Turn one GPU-box session (tools/gpu_round.sh) into the tracked summaries under profiles/:
    python tools/ncu_summary.py gpurun_out/<tag> profiles/<name>
writes <name>_launches.csv (per-kernel aggregate of the ncu launch list), <name>_ncu_full.csv (selected metrics of
the `ncu --set full` capture, read with `ncu -i ... --page raw --csv`), <name>_bench.json, <name>_sweep.jsonl."""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault((r[4].split("(")[0].replace("void ", ""), r[8], r[7]), []).append(float(r[-1]) / 1e3)
    tot = sum(sum(v) for k, v in agg.items() if "synth" not in k[0])
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "block", "launches", "mean_us", "min_us", "max_us", "sum_us", "share_of_chain"])
        for (k, g, b), v in agg.items():
            w.writerow([k, g, b, len(v), f"{sum(v)/len(v):.1f}", f"{min(v):.1f}", f"{max(v):.1f}", f"{sum(v):.1f}",
                        "" if "synth" in k else f"{sum(v)/tot:.3f}"])


def full(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{m} [{units[idx[m]]}]" for m in METRICS if m in idx])
        for r in rows[2:]:
            w.writerow([r[idx["Kernel Name"]].split("(")[0].replace("void ", "")] + [r[idx[m]] for m in METRICS if m in idx])


def traffic(rep, dst, tag):
    """profiles/traffic_latest.json: measured DRAM bytes per FRAME of every point kernel (dram__bytes_read.sum +
    dram__bytes_write.sum of one `ncu --set full` launch / frames of that launch); bench.py scales it to its chunk."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
        frames = None
        if "Grid Size" in idx:
            g = [int(x) for x in r[idx["Grid Size"]].strip("()").split(",")]
            # point kernels and k_outline: grid = (blocks or plateaus, frames); the per-frame kernels: grid = (frames)
            frames = g[0] if name in ("k_peaks", "k_frame_logic", "k_finalize") else g[1]
        b = sum(float(r[idx[m]]) * scale[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        if frames:
            per[name] = {"bytes_per_frame": b / frames, "frames_per_launch": frames, "us": float(r[idx["gpu__time_duration.sum"]]) *
                         {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1.0)}
    with open(dst, "w") as f:
        json.dump({"source": f"profiles/{tag}_ncu_full.csv (ncu --set full --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum per launch / frames per launch)",
                   "per_kernel": per}, f, indent=1)


def main():
    src, name = sys.argv[1], sys.argv[2]
    os.makedirs(os.path.dirname(name) or ".", exist_ok=True)
    if os.path.exists(f"{src}/launches.csv"):
        launches(f"{src}/launches.csv", f"{name}_launches.csv")
    if os.path.exists(f"{src}/prof.ncu-rep"):
        full(f"{src}/prof.ncu-rep", f"{name}_ncu_full.csv")
        traffic(f"{src}/prof.ncu-rep", os.path.join(os.path.dirname(name) or ".", "traffic_latest.json"), os.path.basename(name))
    for fn in ("bench.json", "sweep.jsonl"):
        if os.path.exists(f"{src}/{fn}") and os.path.getsize(f"{src}/{fn}"):
            shutil.copy(f"{src}/{fn}", f"{name}_{fn}")


if __name__ == "__main__":
    main()
