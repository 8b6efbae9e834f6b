"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
Usage: python scripts/launch_list.py gpurun_out/launches.csv "<command that was profiled>" > profiles/xyz.txt"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
cmd = sys.argv[2] if len(sys.argv) > 2 else "?"
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].isdigit()]
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows:
    name = r[4].split("(")[0][:150]
    tot[name] += float(r[-1]) / 1e6
    cnt[name] += 1
total = sum(tot.values())
print(f"# ncu launch list of `{cmd}` (gpu__time_duration.sum, --clock-control none)")
print("# per-launch times are serialised/cold-cache: compare SHARES. Our kernels: nb200::nuts_kernel<...>.")
print(f"# total device time {total:.1f} ms over {len(rows)} launches\n")
for name, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{100 * ms / total:6.2f}%  {ms:10.2f} ms  {cnt[name]:4d} launches  {name}")
