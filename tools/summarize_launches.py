"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares.
    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.txt"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else (v * 1e6 if u == "s" else v))
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {T / 1e3:.2f} ms of kernel time (ncu: cold-cache, serialised — compare shares)")
    print(f"{'kernel':42s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{k[:42]:42s} {cnt[k]:8d} {v / 1e3:10.2f} {v / T * 100:6.1f}% {v / cnt[k]:9.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
