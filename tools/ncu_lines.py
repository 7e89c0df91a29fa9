#!/usr/bin/env python
"""Per-source-line summary (samples, warp instructions) of an ncu report imported with --import-source on.
usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] == 'Function Name' or hdr is None: continue
    if r[0] != '':   # a source line row with aggregated metrics
        d = dict(zip(hdr, r))
        def num(k):
            try: return float(d.get(k, '0').replace(',', ''))
            except ValueError: return 0.0
        agg.append((fname, r[0], r[1].strip()[:90], num('Warp Stall Sampling (All Samples)'), num('Instructions Executed'),
                    num('L1 Wavefronts Shared'), num('stall_long_sb'), num('stall_short_sb'), num('stall_wait'), num('stall_math'), num('stall_barrier')))
tot_s = sum(a[3] for a in agg) or 1; tot_i = sum(a[4] for a in agg) or 1
print(f"total samples {tot_s:.0f}  total warp-inst {tot_i:.3e}")
print(f"{'file:line':28s} {'smp%':>6s} {'inst%':>6s} {'smemWF':>10s} {'lsb':>6s} {'ssb':>6s} {'wait':>6s} {'math':>6s} {'bar':>6s}  source")
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print(f"{a[0]+':'+a[1]:28s} {100*a[3]/tot_s:6.2f} {100*a[4]/tot_i:6.2f} {a[5]:10.3g} {a[6]:6.0f} {a[7]:6.0f} {a[8]:6.0f} {a[9]:6.0f} {a[10]:6.0f}  {a[2]}")
