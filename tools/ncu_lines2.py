#!/usr/bin/env python
"""Per-source-line table of an ncu report imported with --import-source on: warp instructions, shared-memory wavefronts and
wavefronts per executed shared-memory instruction, for every line that touches shared memory or is hot.
usage: tools/ncu_lines2.py report.ncu-rep [min_inst_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] in ('File Path', 'Function Name') or hdr is None: continue
    if r[0] != '' and r[0].isdigit():
        d = dict(zip(hdr, r))
        def num(k):
            try: return float(d.get(k, '0').replace(',', ''))
            except ValueError: return 0.0
        agg.append((int(r[0]), r[1].strip()[:100], num('Instructions Executed'), num('L1 Wavefronts Shared'), num('L1 Wavefronts Shared Ideal'),
                    num('Warp Stall Sampling (All Samples)'), num('Avg. Threads Executed')))
tot_i = sum(a[2] for a in agg) or 1; tot_w = sum(a[3] for a in agg) or 1; tot_s = sum(a[5] for a in agg) or 1
print(f"total warp-inst {tot_i:.4e}  total smem wavefronts {tot_w:.4e}  samples {tot_s:.0f}")
print(f"{'line':>5s} {'inst%':>6s} {'smp%':>6s} {'smemWF%':>7s} {'WF':>10s} {'ideal':>10s} {'thr':>5s}  source")
for a in sorted(agg):
    if 100*a[2]/tot_i >= minpct or a[3] > 0:
        print(f"{a[0]:5d} {100*a[2]/tot_i:6.2f} {100*a[5]/tot_s:6.2f} {100*a[3]/tot_w:7.2f} {a[3]:10.3g} {a[4]:10.3g} {a[6]:5.1f}  {a[1]}")
