"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the last
complete training step (between the last two adamw launches).  Usage: launch_summary.py launches.csv"""
import collections
import csv
import re
import sys


def short(name):
    m = re.search(r"([A-Za-z0-9_]+)(<[^(]*)?\(", name)
    base = m.group(1) if m else name
    targs = m.group(2) if m and m.group(2) else ""
    targs = re.sub(r"\(anonymous namespace\)::|afb::|__nv_bfloat16", lambda k: "bf16" if "bfloat" in k.group(0) else "", targs)
    return (base + targs)[:70]


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    idx = [i for i, n in enumerate(names) if "adamw" in n]
    step = rows[idx[-2] + 2: idx[-1] + 2] if len(idx) >= 2 else rows
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in step:
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        n = short(r["Kernel Name"])
        tot[n] += v
        cnt[n] += 1
    T = sum(tot.values())
    print(f"# last full step: {len(step)} launches, {T / 1e3:.2f} ms (ncu-serialised, cold cache)")
    print(f"{'total_us':>10} {'share':>6} {'n':>4} {'avg_us':>9}  kernel")
    for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{v:10.1f} {100 * v / T:5.1f}% {cnt[n]:4d} {v / cnt[n]:9.1f}  {n}")


if __name__ == "__main__":
    main(sys.argv[1])
