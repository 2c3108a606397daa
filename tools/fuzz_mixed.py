#!/usr/bin/env python
"""Mixed read lengths (SURVEY trap 3): the CPU oracle against the unmodified reference binary run with -p 1.

Reads whose start-offset range is empty ((len - I + 1) % s == 0) inherit the previous read's start offset and stale
seed hashes in the reference (align.cpp:476-480, 79-150); with -p 1 that is deterministic. Every case trims the
simulated fixed-length reads of a small synthetic config to random lengths that mix both kinds.
Build container only. usage: fuzz_mixed.py [n_cases] [seed]; exit code = number of differences."""
import dataclasses
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import synth  # noqa: E402


def trim_fastq(path, rng, lengths, first_full):
    """Rewrite a FASTQ file with every read cut to a length drawn from `lengths` (the first `first_full` reads stay whole)."""
    out = []
    with open(path) as fh:
        lines = fh.read().split("\n")
    k = 0
    for i in range(0, len(lines) - 3, 4):
        name, seq, plus, qual = lines[i:i + 4]
        if k >= first_full:
            n = min(len(seq), int(rng.choice(lengths)))
            seq, qual = seq[:n], qual[:n]
        out += [name, seq, plus, qual]; k += 1
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    bad = 0
    for it in range(n):
        cid = int(rng.choice([1, 2, 2, 3, 4]))
        cfg = synth.baseline_config(cid, float(rng.choice([0.0004, 0.001])))
        rule = cfg.rule
        s, I = (16, 4) if rng.random() < 0.6 else (int(rng.integers(10, 17)), int(rng.integers(1, 7)))
        L = cfg.read_len
        empty = [x for x in range(max(40, s + I), L + 1) if (x - I + 1) % s == 0]
        other = [x for x in range(max(40, s + I), L + 1) if (x - I + 1) % s != 0]
        lengths = list(rng.choice(empty, size=min(3, len(empty)), replace=False)) + list(rng.choice(other, size=4, replace=False)) + [L]
        flags = ["-S", str(int(rng.integers(1, 99999))), "-s", str(s), "-I", str(I), "-u"]
        if rng.random() < 0.4: flags += ["-g", str(int(rng.integers(1, 4)))]
        if rng.random() < 0.4: flags += ["-n", str(int(rng.integers(0, 3)))]
        if rng.random() < 0.3: flags += ["-w", str(int(rng.choice([2, 5, 50])))]
        with tempfile.TemporaryDirectory() as tmp:
            paths = synth.materialise(dataclasses.replace(cfg, rule=rule), tmp, limit=1200)
            first_full = int(rng.choice([0, 1, 1]))
            for p in (paths["a"], paths["b"]):
                if p:
                    trim_fastq(p, rng, lengths, first_full)
            args = ["-a", os.path.basename(paths["a"])] + (["-b", os.path.basename(paths["b"])] if paths["b"] else []) + ["-d", "ref.fa", "-M", rule] + flags
            try:
                want = helpers.run_cli(helpers.REF_BIN, args + ["-p", "1"], tmp, "ref.sam")
            except Exception as e:
                print(f"[{it}] reference failed: {' '.join(args)} :: {str(e)[-120:]}"); continue
            got = helpers.run_cli(helpers.ORACLE_BIN, args, tmp, "orc.sam")
            if got != want:
                bad += 1
                g, w = got.splitlines(), want.splitlines()
                nd = sum(1 for x, y in zip(g, w) if x != y)
                k = next((i for i, (x, y) in enumerate(zip(g, w)) if x != y), min(len(g), len(w)))
                print(f"[{it}] DIFF cfg{cid} lengths {sorted(set(int(x) for x in lengths))} first_full {first_full} {' '.join(args)}  lines {len(g)} vs {len(w)}, {nd} differ, first at {k}:\n   got  {g[k][:150] if k < len(g) else None}\n   want {w[k][:150] if k < len(w) else None}")
            else:
                print(f"[{it}] ok cfg{cid} {rule} lengths {sorted(set(int(x) for x in lengths))} first_full {first_full} {' '.join(flags)}")
    print("differences:", bad)
    sys.exit(min(bad, 100))


if __name__ == "__main__":
    main()
