#!/usr/bin/env python
"""Turn gpurun_out/<tag>_{launches.csv,*.ncu-rep} into the small tracked summaries under profiles/.

  python scripts/summarize_ncu.py <tag>        # e.g. r01c

Writes profiles/<tag>_launches.md (per-kernel share of the step, from the gpu__time_duration.sum launch list) and
profiles/<tag>_<name>.csv (one row per captured launch with the roofline-relevant raw metrics of the --set full pass).
"""
import collections
import csv
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = (
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
    "sm__cycles_elapsed.max.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
)


def launches(tag):
    path = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    if not os.path.exists(path):
        return
    hdr, agg = None, collections.OrderedDict()
    for r in csv.reader(open(path, errors="replace")):
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        try:
            ns = float(d["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        k = (d["Kernel Name"].split("(")[0], d["Grid Size"], d["Block Size"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    byname = collections.OrderedDict()
    for (n, g, b), (c, t) in agg.items():
        e = byname.setdefault(n, [0, 0.0, set()])
        e[0] += c
        e[1] += t
        e[2].add(g)
    out = os.path.join(ROOT, "profiles", f"{tag}_launches.md")
    with open(out, "w") as f:
        f.write(f"# {tag}: launch list (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n")
        f.write("Per-launch times under ncu are serialised and cold-cache: compare SHARES with bench.py's "
                "`kernel_ms_per_step`, not absolutes.\n\n| kernel | launches | total ms | share | grids |\n|---|---:|---:|---:|---|\n")
        for n, (c, t, g) in sorted(byname.items(), key=lambda x: -x[1][1]):
            gs = ", ".join(sorted(g)[:4]) + (" …" if len(g) > 4 else "")
            f.write(f"| `{n}` | {c} | {t / 1e6:.3f} | {100 * t / tot:.1f}% | {gs} |\n")
        f.write(f"\ntotal {tot / 1e6:.3f} ms over {sum(v[0] for v in byname.values())} launches\n")
    print("wrote", out)


def report(path, tag):
    name = os.path.basename(path)[len(tag) + 1:-len(".ncu-rep")]
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    if len(rows) < 3:
        print("no rows in", path)
        return
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if h in KEEP]
    out = os.path.join(ROOT, "profiles", f"{tag}_{name}.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
        for i in cols:
            vals = [r[i] for r in rows[2:]]
            if hdr[i] == "Kernel Name":
                vals = [v.split("(")[0] for v in vals]
            w.writerow([hdr[i], units[i]] + vals)
    print("wrote", out)


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag)
    for p in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_*.ncu-rep"))):
        report(p, tag)
