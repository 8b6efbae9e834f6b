"""One-page text summary of an ncu report (headline metrics, stall mix, hottest source lines).
Usage: python scripts/ncu_summary.py report.ncu-rep > profiles/xyz.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, zip(vals, units)))
def g(k):
    v = d.get(k)
    return f"{v[0]} {v[1]}" if v else "n/a"
print(f"report: {rep}")
print(f"kernel: {d.get('Kernel Name', ('?',''))[0]}")
print(f"grid/block: {g('Grid Size')} / {g('Block Size')}")
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for k in keys:
    print(f"  {k:70s} {g(k)}")
print("stall mix (warps stalled per issue):")
for k in hdr:
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        v = float(d[k][0] or 0)
        if v >= 0.05:
            print(f"  {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
try:
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Line No')
    agg = []
    for r in rows[hi + 1:]:
        if len(r) < 8: continue
        if r[0] != '' and r[2] == '-':
            try: agg.append((int(r[4] or 0), int(r[7] or 0), r[0], r[1].strip()[:100]))
            except ValueError: pass
    ts = sum(a[0] for a in agg); ti = sum(a[1] for a in agg)
    print(f"hottest source lines (of {ts} stall samples, {ti} warp instructions):")
    for s, i, ln, text in sorted(agg, reverse=True)[:20]:
        print(f"  {100*s/ts:5.1f}% samples {100*i/ti:5.1f}% inst  L{ln}: {text}")
except StopIteration:
    print("(no source correlation in this report)")
