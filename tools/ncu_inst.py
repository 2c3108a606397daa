#!/usr/bin/env python
"""Per-source-line share of executed warp instructions of a kernel in an .ncu-rep. usage: ncu_inst.py rep [top]"""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; lines = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None: continue
    try: ln = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr[2:], r[2:]))
    try: lines[(cur, ln)] = (float(d['Instructions Executed']), r[1][:130])
    except ValueError: pass
tot = sum(v[0] for v in lines.values()); acc = 0
print(f"total warp instructions {tot:.0f}")
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    acc += v[0]
    print(f"{k[0]}:{k[1]:<5d} {100 * v[0] / tot:5.1f}% cum {100 * acc / tot:5.1f}%  {v[1].strip()}")
