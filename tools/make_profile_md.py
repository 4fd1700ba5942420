"""Writes profiles/<name>.md from an ncu launch list (csv) and optional ncu --set full reports.

    python tools/make_profile_md.py profiles/r1_final_launch_list.md gpurun_out/launches.csv "<command>" [rep.ncu-rep ...]
"""
import collections
import csv
import gzip
import os
import shutil
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from launch_summary import load  # noqa: E402

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def rep_rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("unnamed>::", ""),
             "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for w in WANT:
            if w in hdr:
                d[w] = r[hdr.index(w)] + " " + units[hdr.index(w)]
        res.append(d)
    return res


def main():
    out_md, launches, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    reps = sys.argv[4:]
    rows = load(launches)
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for _, name, _, _, v in rows:
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    with open(out_md, "w") as f:
        f.write(f"# ncu launch list — per-launch `gpu__time_duration.sum` (cold cache, serialised: compare SHARES)\n\n")
        f.write(f"Command: `{cmd}`\n\n{len(rows)} launches, {tot / 1e3:.2f} ms of kernel time.\n\n")
        f.write("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {t / 1e3:.2f} | {t / c:.1f} | {100 * t / tot:.1f}% |\n")
        for rp in reps:
            f.write(f"\n## `ncu --set full` capture `{os.path.basename(rp)}`\n\n")
            for d in rep_rows(rp):
                f.write(f"* `{d['kernel']}` grid {d['grid']} block {d['block']}\n")
                for w in WANT:
                    if w in d:
                        f.write(f"  * {w}: {d[w]}\n")
    gz = out_md.replace(".md", ".csv.gz")
    with open(launches, "rb") as fi, gzip.open(gz, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    print("wrote", out_md, gz)


if __name__ == "__main__":
    main()
