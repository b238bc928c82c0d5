"""Summarise an `ncu --page source --csv --print-source sass` dump: instruction mix and top stall sites.
Usage: ncu -i X.ncu-rep --page source --csv --print-source sass --kernel-id :::K > dump.csv; python tools/ncu_sass_top.py dump.csv [top] [section]"""
import csv
import collections
import sys

allrows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # n-th kernel section of the dump
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
rows = allrows[starts[which]:starts[which + 1]]
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
print(f"[{which + 1}/{len(starts) - 1}]", rows[0][1][:110])
tot_inst = sum(int(r[ci["Instructions Executed"]]) for r in body)
tot_samp = sum(int(r[ci["# Samples"]]) for r in body)
print(f"instructions executed {tot_inst}   stall samples {tot_samp}   sass lines {len(body)}")
mix = collections.Counter()
for r in body:
    op = r[ci["Source"]].split()
    op = [o for o in op if not o.startswith("@")]
    mix[op[0].split(".")[0] if op else "?"] += int(r[ci["Instructions Executed"]])
print("mix:", ", ".join(f"{k} {100 * v / tot_inst:.1f}%" for k, v in mix.most_common(22)))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.Counter()
for r in body:
    for h in stall_cols:
        agg[h] += int(r[ci[h]] or 0)
print("stalls:", ", ".join(f"{k[6:]} {100 * v / max(tot_samp, 1):.1f}%" for k, v in agg.most_common(10)))
print("top stall sites:")
for r in sorted(body, key=lambda r: -int(r[ci["# Samples"]]))[:top]:
    why = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"  {100 * int(r[ci['# Samples']]) / max(tot_samp, 1):5.1f}%  exec {int(r[ci['Instructions Executed']]):>9}  {r[ci['Source']].strip()[:70]:70s} {why}")
