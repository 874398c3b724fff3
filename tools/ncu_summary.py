"""Summarise ncu outputs into the small CSV / text files kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/rNN_launches.txt [--skip-warmup-steps W]
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/rNN_ncu_full_summary.csv

`launches`: per-kernel launch count, total and share of the captured list (gpu__time_duration.sum pass).
`full`: one row per captured launch with the counters the roofline discussion uses.
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "smsp__inst_executed.sum",
    "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct",
]


def launches(src: str, dst: str) -> None:
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0]
        a = agg.setdefault(name, [0, 0.0, r[7], r[8]])
        a[0] += 1
        a[1] += float(r[14])
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {src}: {len(rows)} launches, {total / 1e6:.3f} ms of kernel time (cold-cache, serialised by ncu)\n")
        f.write(f"{'kernel':60s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}  block grid(last)\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name[:60]:60s} {a[0]:5d} {a[1] / 1e3:10.1f} {a[1] / 1e3 / a[0]:9.2f} {100 * a[1] / total:6.1f}%  {a[2]} {a[3]}\n")
    print(open(dst).read())


def full(src: str, dst: str) -> None:
    out = subprocess.run(
        ["ncu", "-i", src, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True
    ).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr = rd[0]
    keep = [hdr.index("Kernel Name")] + [hdr.index(m) for m in METRICS if m in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        for r in rd:
            row = [r[i] for i in keep]
            row[0] = row[0][:70]
            w.writerow(row)
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
