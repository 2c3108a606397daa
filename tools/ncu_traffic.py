#!/usr/bin/env python
"""DRAM traffic of every captured launch of one kernel in an .ncu-rep -> profiles/roofline_traffic.json
usage: ncu_traffic.py report.ncu-rep out.json [candidates_per_step bytes_per_candidate]
The file records the SHA-256 of the kernel source it was captured from (basal_b200/csrc/align.cu); bench.py reports
roofline.traffic only while that still matches."""
import csv, hashlib, io, json, os, subprocess, sys
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
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
res["kernel_source_sha256"] = hashlib.sha256(open(os.path.join(root, "basal_b200", "csrc", "align.cu"), "rb").read()).hexdigest()
if len(sys.argv) > 4:
    cands, bpc = float(sys.argv[3]), float(sys.argv[4])
    res.update({"candidates_per_step": cands, "algorithmic_bytes_per_launch": cands * bpc / n, "dram_bytes_per_candidate": (sum(rd) + sum(wr)) / cands})
json.dump(res, open(sys.argv[2], 'w'), indent=1)
print(json.dumps(res, indent=1))
