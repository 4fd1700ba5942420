"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for x in csv.DictReader(lines):
        if x.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        name = re.sub(r"\((int|bool|unsigned int)\)", "", x["Kernel Name"])      # template-argument casts, e.g. <128, 1, (int)-1, 1>
        name = re.sub(r"\(.*", "", name).replace("unnamed>::", "").replace("void ", "")
        rows.append((int(x["ID"]), name, x["Grid Size"], x["Block Size"], v))
    return rows


def main():
    rows = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for _, name, _, _, v in rows:
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"{len(rows)} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e3:9.2f} ms {100 * t / tot:5.1f}%  n={c:5d}  avg={t / c:8.1f} us  {k}")


if __name__ == "__main__":
    main()
