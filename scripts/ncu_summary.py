"""Summarise ncu artefacts brought back in gpurun_out/ into small text files under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/NAME.md
    python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/NAME.md
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "lts__t_sector_hit_rate.pct",
]


def launches(src, dst):
    with open(src) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    scale = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", "")) * scale.get(row.get("Metric Unit", "ns"), 1.0)
        agg[row["Kernel Name"][:90]][0] += 1
        agg[row["Kernel Name"][:90]][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as out:
        out.write(f"# ncu launch list summary of {src}\n\n(gpu__time_duration.sum, --clock-control none; cold-cache, "
                  "serialised: compare SHARES)\n\n| total ms | launches | share | kernel |\n|---:|---:|---:|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
            out.write(f"| {v[1] / 1e6:.3f} | {v[0]} | {100 * v[1] / tot:.1f}% | `{k}` |\n")
        out.write(f"\ntotal {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches\n")


def full(src, dst):
    """src: a .ncu-rep, or the CSV `ncu -i X.ncu-rep --page raw --csv` wrote on the GPU box (scripts/r2_ncu.sh keeps
    only that for captures too large to bring back)."""
    if src.endswith(".csv"):
        with open(src) as fh:
            raw = fh.read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as out:
        out.write(f"# ncu --set full summary of {src}\n\n")
        for row in rows[2:]:
            out.write(f"## `{row[hdr.index('Kernel Name')][:100]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    out.write(f"| {k} | {row[i]} | {units[i]} |\n")
            out.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
