"""Summarise gpurun_out ncu artefacts into small tracked text files under profiles/.

    python profiles/summarize.py <tag> [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof.ncu-rep]

Writes profiles/<tag>_launches.txt (per-kernel launch count, mean duration and SHARE of the step;
ncu times are cold-cache and serialised, so only the shares are comparable with the CUDA-event
numbers bench.py prints) and profiles/<tag>_kernels.txt (key metrics of the --set full capture).
"""
import argparse
import collections
import csv
import os
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct",
]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) > vi:
            agg.setdefault(r[ki].split("(")[0][:70], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# source: {path} (ncu --metrics gpu__time_duration.sum --clock-control none)\n")
        f.write(f"# {'kernel':70s} {'n':>4s} {'mean_ms':>10s} {'max_ms':>10s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:72s} {len(v):4d} {sum(v) / len(v) / 1e6:10.3f} {max(v) / 1e6:10.3f} {sum(v) / tot * 100:6.1f}%\n")
    print(open(out).read())


def kernels(rep, out):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(raw.splitlines()))
    H, U = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# source: {rep} (ncu --set full --clock-control none --import-source on)\n")
        for r in rows[2:]:
            f.write(f"\n== {r[H.index('Kernel Name')][:100]}\n")
            for m in METRICS:
                if m in H:
                    i = H.index(m)
                    f.write(f"   {m:80s} {r[i]:>16s} {U[i]}\n")
    print(open(out).read())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    a = ap.parse_args()
    here = os.path.dirname(os.path.abspath(__file__))
    if a.launches:
        launches(a.launches, os.path.join(here, f"{a.tag}_launches.txt"))
    if a.rep:
        kernels(a.rep, os.path.join(here, f"{a.tag}_kernels.txt"))
