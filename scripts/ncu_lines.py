"""Aggregate an ncu report's warp-stall samples and executed instructions by CUDA source line.
Usage: python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Line No')
agg = []
for r in rows[hi + 1:]:
    if len(r) < 8: continue
    if r[0] != '' and r[2] == '-':
        try: agg.append((int(r[4] or 0), int(r[7] or 0), r[0], r[1].strip()[:115]))
        except ValueError: pass
ts = sum(a[0] for a in agg); ti = sum(a[1] for a in agg)
print("total samples", ts, "total warp-instructions", ti)
for s, i, ln, src in sorted(agg, reverse=True)[:top]:
    print(f"{100*s/ts:5.1f}% samp {100*i/ti:5.1f}% inst  L{ln}: {src}")
