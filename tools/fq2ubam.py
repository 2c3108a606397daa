#!/usr/bin/env python
"""FASTQ (one file, or two mate files -> interleaved) -> unaligned SAM text on stdout (flag 4 / 77 / 141), for
basal_b200/bin/sam2bam: test input for the CLI's BAM read path (reads.cpp:85-108).  usage: fq2ubam.py a.fq [b.fq] > u.sam"""
import sys


def records(path):
    with open(path) as fh:
        while True:
            h = fh.readline()
            if not h:
                return
            s = fh.readline().strip(); fh.readline(); q = fh.readline().strip()
            yield h[1:].split()[0], s, q


def main():
    a = records(sys.argv[1]); b = records(sys.argv[2]) if len(sys.argv) > 2 else None
    out = ["@HD\tVN:1.0\tSO:unsorted"]
    for ra in a:
        if b is None:
            out.append(f"{ra[0]}\t4\t*\t0\t0\t*\t*\t0\t0\t{ra[1]}\t{ra[2]}")
        else:
            rb = next(b)
            out.append(f"{ra[0]}\t77\t*\t0\t0\t*\t*\t0\t0\t{ra[1]}\t{ra[2]}")
            out.append(f"{rb[0]}\t141\t*\t0\t0\t*\t*\t0\t0\t{rb[1]}\t{rb[2]}")
    sys.stdout.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
