#!/usr/bin/env python
"""Per-source-line hot spots of one kernel in an .ncu-rep (compiled with -lineinfo, captured with --import-source on).
usage: ncu_lines.py report.ncu-rep [top_n] [kernel-index]"""
import csv
import io
import subprocess
import sys


def main(path, top=25):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file = None; hdr = None; lines = {}
    kern = None; seen_kernels = []
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]; continue
        if r[0] == 'Function Name':
            kern = r[1]
            if kern not in seen_kernels: seen_kernels.append(kern)
            continue
        if r[0] == 'Line No':
            hdr = r; continue
        if hdr is None or r[0] in ('', 'Kernel Name'):
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr[2:], r[2:]))
        def num(k):
            try: return float(d.get(k, '0').replace(',', ''))
            except ValueError: return 0.0
        key = (cur_file, ln)
        e = lines.setdefault(key, [r[1][:110], 0, 0, 0, 0, 0, 0])
        e[1] += num('# Samples'); e[2] += num('Instructions Executed'); e[3] += num('stall_long_sb'); e[4] += num('stall_short_sb') + num('stall_mio')
        e[5] += num('L1 Wavefronts Shared'); e[6] += num('stall_barrier')
    tot_s = sum(v[1] for v in lines.values()) or 1; tot_i = sum(v[2] for v in lines.values()) or 1
    print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
    print(f"{'file:line':22s} {'smp%':>6s} {'inst%':>6s} {'longsb':>7s} {'sh/mio':>7s} {'barr':>6s} {'smemWF':>10s}  source")
    for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{f + ':' + str(ln):22s} {100 * v[1] / tot_s:6.1f} {100 * v[2] / tot_i:6.1f} {v[3]:7.0f} {v[4]:7.0f} {v[6]:6.0f} {v[5]:10.0f}  {v[0].strip()}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
