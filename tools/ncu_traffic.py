#!/usr/bin/env python
"""DRAM traffic of every captured launch of one kernel in an .ncu-rep -> profiles/roofline_traffic.json
usage: ncu_traffic.py report.ncu-rep out.json [launches_per_step]"""
import csv, io, json, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
def col(name):
    i = hdr.index(name); u = units[i]
    mul = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1, 'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1}.get(u, 1)
    return [float(r[i].replace(',', '')) * mul for r in rows[2:]]
rd, wr, dur = col('dram__bytes_read.sum'), col('dram__bytes_write.sum'), col('gpu__time_duration.sum')
n = len(rd)
res = {"kernel": rows[2][hdr.index('Kernel Name')][:80], "launches": n, "dram_bytes_total": sum(rd) + sum(wr),
       "dram_bytes_per_launch": (sum(rd) + sum(wr)) / n, "dram_bytes_read_total": sum(rd), "dram_bytes_write_total": sum(wr),
       "ncu_duration_total_s": sum(dur), "largest_launch": {"dram_bytes": max(a + b for a, b in zip(rd, wr)), "duration_s": max(dur)},
       "source": sys.argv[1].split('/')[-1], "note": "ncu --set full replays each launch cold-cache and serialised; bytes are per launch, summed over one step's launches"}
json.dump(res, open(sys.argv[2], 'w'), indent=1)
print(json.dumps(res, indent=1))
