#!/usr/bin/env python
"""Golden hashes for the native BAM writer: SHA-256 of the DECOMPRESSED BAM stream that the reference's own samtools
(0.1.18, vendored under /root/reference/samtools) produces with `samtools view -bS` for every tests/golden/*/expected.sam.

Runs only in the build container (needs /root/reference); the samtools binary is compiled from the sources where they
lie into a temporary directory (nothing is copied into the repository). Output: tests/golden/bam_sha256.json.
"""
import glob
import gzip
import hashlib
import json
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = "/root/reference/samtools"
MAIN = ("bgzf kstring bam_aux bam bam_import sam bam_index bam_pileup bam_lpileup bam_md razf faidx bedidx knetfile bam_sort "
        "sam_header bam_reheader kprobaln bam_cat bam_tview bam_plcmd sam_view bam_rmdup bam_rmdupse bam_mate bam_stat bam_color "
        "bamtk kaln bam2bcf bam2bcf_indel errmod sample cut_target phase bam2depth").split()
BCF = "bcf vcf bcfutils prob1 em kfunc kmin index fet mut bcf2qcall".split()
FLAGS = ["-w", "-O2", "-D_FILE_OFFSET_BITS=64", "-D_LARGEFILE64_SOURCE", "-D_USE_KNETFILE", "-D_CURSES_LIB=0", f"-I{SRC}", f"-I{SRC}/bcftools"]


def build(tmp):
    objs = []
    for f in MAIN:
        o = os.path.join(tmp, f + ".o"); subprocess.check_call(["gcc", "-c", *FLAGS, f"{SRC}/{f}.c", "-o", o]); objs.append(o)
    for f in BCF:
        o = os.path.join(tmp, "bcf_" + f + ".o"); subprocess.check_call(["gcc", "-c", *FLAGS, f"{SRC}/bcftools/{f}.c", "-o", o]); objs.append(o)
    exe = os.path.join(tmp, "samtools")
    subprocess.check_call(["gcc", "-o", exe, *objs, "-lm", "-lz", "-lpthread"])
    return exe


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        exe = build(tmp)
        for sam in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*", "expected.sam"))):
            bam = os.path.join(tmp, "x.bam")
            with open(bam, "wb") as fh:
                subprocess.check_call([exe, "view", "-bS", sam], stdout=fh, stderr=subprocess.DEVNULL)
            raw = gzip.open(bam, "rb").read()
            out[os.path.basename(os.path.dirname(sam))] = {"sha256": hashlib.sha256(raw).hexdigest(), "bytes": len(raw)}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "bam_sha256.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
