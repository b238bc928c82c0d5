"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list: per-kernel
time and DRAM bytes of the last complete training step (between the last two adamw launches).
Usage: launch_summary3.py launches.csv"""
import collections
import csv
import sys

from launch_summary import short

UNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6,
        "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()          # launch id -> {name, us, rd, wr}
    for r in rows:
        e = per.setdefault(r["ID"], {"name": r["Kernel Name"], "us": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        if r["Metric Name"].startswith("gpu__time"):
            e["us"] = v
        elif "read" in r["Metric Name"]:
            e["rd"] = v
        else:
            e["wr"] = v
    launches = list(per.values())
    idx = [i for i, e in enumerate(launches) if "adamw" in e["name"]]
    step = launches[idx[-2] + 2: idx[-1] + 2] if len(idx) >= 2 else launches
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
    for e in step:
        a = agg[short(e["name"])]
        a[0] += e["us"]; a[1] += e["rd"]; a[2] += e["wr"]; a[3] += 1
    T = sum(a[0] for a in agg.values())
    MB = sum(a[1] + a[2] for a in agg.values())
    print(f"# last full step: {len(step)} launches, {T / 1e3:.2f} ms (ncu-serialised, cold cache), DRAM traffic {MB / 1e3:.2f} GB "
          f"(= {MB / T * 1e3:.0f} GB/s averaged over the kernel time; ncu runs every kernel alone with cold caches, so part of each output is still in L2 when the kernel ends: written MB are under-counted)")
    print(f"{'total_us':>10} {'share':>6} {'n':>4} {'avg_us':>8} {'rd_MB':>8} {'wr_MB':>8} {'GB/s':>7}  kernel   (MB per launch)")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{a[0]:10.1f} {100 * a[0] / T:5.1f}% {a[3]:4d} {a[0] / a[3]:8.1f} {a[1] / a[3]:8.1f} {a[2] / a[3]:8.1f} {(a[1] + a[2]) / a[0] * 1e3:7.0f}  {n}")


if __name__ == "__main__":
    main(sys.argv[1])
