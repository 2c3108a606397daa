#!/usr/bin/env python
"""Generates the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE BINARY.

    make -f oracle/Makefile.ref          # builds oracle/_ref/basal from /root/reference
    python tests/golden/make_golden.py   # rewrites tests/golden/<case>/{ref.fa,reads*.fq,expected.sam} + cases.json

The reference has no tests or golden vectors of its own, so these files pin the oracle
(and through it the CUDA path) to the reference's observable behaviour: the SAM text it prints with a
fixed -S.  The @PG line (which echoes argv) is dropped.  Inputs are tiny and hand-shaped to hit the edge
cases SURVEY.md §8c lists: N runs, lowercase and IUPAC letters, short islands, several sequences, reads at
chromosome ends, planted repeat copies (multi-hits, -w, -r), indels on both strands, all -M families,
-n 0/1/2, -u, -R, PE insert limits.
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "basal")
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_ref(rng):
    """3 sequences: repeats, an N run, lowercase and IUPAC letters, a short island, a tiny last sequence."""
    seqs = []
    for n in (6000, 4100, 333):
        seqs.append(rng.integers(0, 4, n).astype(np.uint8))
    unit = rng.integers(0, 4, 220).astype(np.uint8)
    for c, p in ((0, 400), (0, 2404), (1, 1208), (1, 3000)):           # 4 copies at positions = 0 mod 4
        seqs[c][p:p + 220] = unit
    rc_unit = (3 - unit)[::-1]
    seqs[0][4800:5020] = rc_unit                                         # and one reverse-complement copy
    near = unit.copy(); near[[30, 97, 160]] = (near[[30, 97, 160]] + 1) & 3
    seqs[1][500:720] = near                                              # a diverged copy (3 mismatches)
    out = []
    for i, s in enumerate(seqs):
        a = ACGT[s].copy()
        if i == 0:
            a[3000:3100] = ord("N"); a[3105:3117] = ord("C")             # 12-base island between N runs (<16: not indexed)
            a[3117:3140] = ord("n")
            a[1500:1600] |= 32                                           # lowercase stretch
            a[[700, 701, 1711]] = [ord("R"), ord("Y"), ord("K")]         # IUPAC letters code as 0
        if i == 1:
            a[:7] = ord("N")                                             # leading Ns
            a[-5:] = ord("N")
        out.append((f"seq{i + 1}", a))
    return out


def write_fa(path, seqs):
    with open(path, "w") as fh:
        for name, a in seqs:
            fh.write(f">{name} some description\n")
            s = a.tobytes().decode()
            for i in range(0, len(s), 70):
                fh.write(s[i:i + 70] + "\n")


def comp(a):
    t = np.zeros(256, np.uint8); t[:] = ord("N")
    for x, y in zip(b"ACGTacgt", b"TGCAtgca"):
        t[x] = y
    return t[a][::-1]


def simulate(rng, seqs, rule, n, L, paired, conv, indel):
    frm = rule[0].encode()[0]
    tos = [c.encode()[0] for c in rule[2:] if c != "-"]
    has_del = "-" in rule[2:]
    r1, r2 = [], []
    for i in range(n):
        c = int(rng.choice(len(seqs), p=[0.55, 0.4, 0.05]))
        a = np.char.upper(seqs[c][1].view("S1")).view(np.uint8) if False else seqs[c][1]
        a = np.frombuffer(a.tobytes().upper(), dtype=np.uint8)
        ins = int(np.clip(rng.normal(260, 40), L, min(500, len(a)))) if paired else L
        if len(a) < ins + 2:
            ins = L
        p = int(rng.integers(0, len(a) - ins + 1))
        if i % 11 == 0:
            p = 0 if i % 2 else len(a) - ins                              # fragments flush with the ends
        if i % 5 == 0 and c < 2:                                          # land in the repeat family
            starts = {0: (400, 2404, 4800), 1: (1208, 3000, 500)}[c]
            p = starts[int(rng.integers(0, 3))] + int(rng.integers(0, 220 - L + 1)) if not paired else p
        frag = a[p:p + ins].copy()
        if rng.random() < 0.5:
            frag = comp(frag)

        def mutate(x):
            x = x.copy()
            if tos:
                m = (x == frm) & (rng.random(len(x)) < conv)
                x[m] = rng.choice(tos, size=int(m.sum()))
            if has_del:
                keep = ~((x == frm) & (rng.random(len(x)) < 0.03))
                x = x[keep]
            e = rng.random(len(x)) < 0.01
            x[e] = rng.choice(list(b"ACGT"), size=int(e.sum()))
            if rng.random() < 0.05:
                x[rng.integers(0, len(x), size=int(rng.integers(1, 8)))] = ord("N")
            if indel and rng.random() < 0.25 and len(x) > 60:
                q = int(rng.integers(15, len(x) - 20)); k = int(rng.integers(1, 4))
                x = np.concatenate([x[:q], x[q + k:]]) if rng.random() < 0.5 else np.concatenate([x[:q], rng.choice(list(b"ACGT"), size=k).astype(np.uint8), x[q:]])
            return x
        ext = 12
        m1_src = a[p:p + min(L + ext, len(a) - p)] if False else None
        # mate 1 = first L bases of the (possibly reverse-complemented) fragment, padded from the reference if needed
        big = frag
        x1 = mutate(big)[:L]
        if len(x1) < L:
            x1 = np.concatenate([x1, rng.choice(list(b"ACGT"), size=L - len(x1)).astype(np.uint8)])
        r1.append(x1)
        if paired:
            x2 = comp(mutate(big))[:L] if False else comp(mutate(big)[-L:])
            if len(x2) < L:
                x2 = np.concatenate([x2, rng.choice(list(b"ACGT"), size=L - len(x2)).astype(np.uint8)])
            r2.append(x2)
    return r1, r2


def write_fq(path, reads, suffix):
    with open(path, "w") as fh:
        for i, r in enumerate(reads):
            s = r.tobytes().decode()
            q = "".join(chr(33 + 30 + (j * 7 + i) % 11) for j in range(len(s)))
            fh.write(f"@read{i}{suffix} extra\n{s}\n+\n{q}\n")


CASES = [
    # name, rule, paired, L, n, conv, indel, args
    ("ct_se", "C:T", False, 100, 400, 0.9, False, ["-S", "7"]),
    ("ct_se_all", "C:T", False, 100, 400, 0.9, True, ["-S", "11", "-g", "3", "-n", "1", "-u", "-R"]),
    ("ct_se_w2_r2", "C:T", False, 60, 300, 0.9, False, ["-S", "7", "-w", "2", "-r", "2", "-u"]),
    ("ct_se_r0", "C:T", False, 60, 300, 0.9, False, ["-S", "5", "-r", "0", "-u", "-n", "2"]),
    ("ct_se_s12", "C:T", False, 75, 300, 0.9, False, ["-S", "7", "-s", "12", "-I", "2", "-v", "5", "-f", "1", "-u"]),
    ("ag_pe", "A:G", True, 100, 300, 0.9, False, ["-S", "7"]),
    ("ag_pe_lim", "A:G", True, 100, 300, 0.9, False, ["-S", "7", "-m", "200", "-x", "280", "-u", "-R"]),
    ("ct_pe_gap", "C:T", True, 90, 300, 0.9, True, ["-S", "3", "-g", "2", "-n", "1", "-u", "-w", "3"]),
    ("acgt_se", "A:CGT", False, 100, 400, 0.05, False, ["-S", "7", "-w", "100", "-u"]),
    ("tdel_se", "T:-", False, 100, 400, 0.0, True, ["-S", "7", "-g", "3", "-u", "-R"]),
    ("gmulti_se", "G:ACT-", False, 100, 300, 0.05, True, ["-S", "9", "-g", "2", "-r", "0", "-u", "-n", "1"]),
    # seed geometry extremes (SURVEY §8c vi); read lengths keep (L - I + 1) % s != 0 (trap 3)
    ("ct_se_i1_s10", "C:T", False, 63, 300, 0.9, False, ["-S", "7", "-s", "10", "-I", "1", "-u"]),
    ("ct_se_i16", "C:T", False, 100, 300, 0.9, False, ["-S", "7", "-I", "16", "-u"]),
    ("ct_se_k", "C:T", False, 100, 300, 0.9, False, ["-S", "12345", "-k", "0.5", "-u"]),
    # strands and filters on pairs (SURVEY §8c iv, vii)
    ("ag_pe_n2", "A:G", True, 100, 300, 0.9, False, ["-S", "13", "-n", "2", "-u"]),
    ("ag_pe_f0", "A:G", True, 100, 300, 0.9, False, ["-S", "7", "-f", "0", "-u"]),
    # -r 2 on pairs: every pair of the best level, and EVERY hit of an unpaired multi-hit mate (pairs.cpp:232-305)
    ("ag_pe_r2", "A:G", True, 100, 300, 0.9, False, ["-S", "7", "-f", "0", "-u", "-r", "2"]),
    # mixed read lengths (SURVEY trap 3): reads with (L - I + 1) % s == 0 (99, 83, 67, 51 at -s 16 -I 4) inherit the start
    # offset and the stale seed hashes of earlier reads of the same aligner object (align.cpp:476-480, 79-150); the
    # reference ran with -p 1, where that is deterministic
    ("ct_se_mixed", "C:T", False, 100, 500, 0.9, False, ["-S", "7", "-u"], [100, 99, 98, 83, 90, 67, 51, 100, 99]),
    ("ag_pe_mixed", "A:G", True, 100, 400, 0.9, False, ["-S", "7", "-u", "-n", "1"], [100, 99, 97, 83, 99, 67, 96]),
    ("ct_se_mixed_s12", "C:T", False, 100, 400, 0.9, True, ["-S", "5", "-s", "12", "-I", "3", "-g", "1", "-u"], [100, 98, 86, 74, 93, 62, 100]),
]


def main():
    if not os.path.exists(REF_BIN):
        sys.exit("oracle/_ref/basal is missing: make -f oracle/Makefile.ref")
    meta = {}
    only = set(sys.argv[1:])                     # optional: regenerate just the named cases
    if only:
        meta = json.load(open(os.path.join(HERE, "cases.json")))
    for ci, case in enumerate(CASES):
        name, rule, paired, L, n, conv, indel, args = case[:8]
        lengths = case[8] if len(case) > 8 else None
        if only and name not in only:
            continue
        rng = np.random.default_rng(100 + ci)
        d = os.path.join(HERE, name)
        os.makedirs(d, exist_ok=True)
        seqs = make_ref(np.random.default_rng(42))
        write_fa(os.path.join(d, "ref.fa"), seqs)
        r1, r2 = simulate(rng, seqs, rule, n, L, paired, conv, indel)
        if lengths:                              # cut read i to a length from the list (mates independently)
            r1 = [r[:lengths[int(rng.integers(0, len(lengths)))]] for r in r1]
            r2 = [r[:lengths[int(rng.integers(0, len(lengths)))]] for r in r2]
        if paired:
            write_fq(os.path.join(d, "reads_1.fq"), r1, "/1"); write_fq(os.path.join(d, "reads_2.fq"), r2, "/2")
            inp = ["-a", "reads_1.fq", "-b", "reads_2.fq"]
        else:
            write_fq(os.path.join(d, "reads.fq"), r1, "")
            inp = ["-a", "reads.fq"]
        full = inp + ["-d", "ref.fa", "-M", rule] + args
        subprocess.run([REF_BIN] + full + ["-p", "1", "-o", "out.sam"], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with open(os.path.join(d, "out.sam")) as fh, open(os.path.join(d, "expected.sam"), "w") as out:
            for line in fh:
                if not line.startswith("@PG"):
                    out.write(line)
        os.unlink(os.path.join(d, "out.sam"))
        meta[name] = {"args": full, "paired": paired}
        print(name, sum(1 for _ in open(os.path.join(d, "expected.sam"))), "lines")
    with open(os.path.join(HERE, "cases.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
