#!/usr/bin/env python
"""Random flag combinations: the CPU oracle (oracle/_build/basal_oracle) against the unmodified reference binary
(oracle/_ref/basal) on small synthetic inputs. Build container only. usage: fuzz_oracle.py [n_cases] [seed]
Prints every combination whose SAM differs (ignoring @PG); exit code = number of differences."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import synth  # noqa: E402

RULES = ["C:T", "A:G", "G:A", "T:C", "A:T", "C:G", "A:CGT", "C:AT", "T:-", "G:ACT-", "A:C-"]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    bad = 0
    for it in range(n):
        cid = int(rng.choice([1, 2, 3, 4]))
        cfg = synth.baseline_config(cid, float(rng.choice([0.0004, 0.001, 0.003])))
        rule = str(rng.choice(RULES))
        s = int(rng.integers(10, 17)); I = int(rng.integers(1, 9))
        L = cfg.read_len
        if (L - I + 1) % s == 0:
            I += 1                                              # SURVEY trap 3: stale start offset in the reference
        flags = ["-S", str(int(rng.integers(1, 99999))), "-s", str(s), "-I", str(I)]
        if rng.random() < 0.5: flags += ["-g", str(int(rng.integers(0, 4)))]
        if rng.random() < 0.5: flags += ["-w", str(int(rng.choice([1, 2, 5, 50, 1000])))]
        if rng.random() < 0.5: flags += ["-n", str(int(rng.integers(0, 3)))]
        if rng.random() < 0.4: flags += ["-r", str(int(rng.choice([0, 1] if cfg.paired else [0, 1, 2])))]
        if rng.random() < 0.5: flags += ["-v", str(rng.choice(["0.05", "0.1", "3", "8", "15"]))]
        if rng.random() < 0.5: flags += ["-u"]
        if rng.random() < 0.3: flags += ["-R"]
        if rng.random() < 0.3: flags += ["-f", str(int(rng.integers(0, 4)))]
        if cfg.paired and rng.random() < 0.4: flags += ["-m", str(int(rng.integers(28, 300))), "-x", str(int(rng.integers(300, 700)))]
        if rng.random() < 0.2: flags += ["-k", str(rng.choice(["1e-7", "1e-4", "0.01"]))]
        with tempfile.TemporaryDirectory() as tmp:
            import dataclasses
            cfg2 = dataclasses.replace(cfg, rule=rule)
            paths = synth.materialise(cfg2, tmp, limit=1500)
            args = ["-a", os.path.basename(paths["a"])] + (["-b", os.path.basename(paths["b"])] if paths["b"] else []) + ["-d", "ref.fa", "-M", rule] + flags
            try:
                want = helpers.run_cli(helpers.REF_BIN, args + ["-p", "1"], tmp, "ref.sam")
            except Exception as e:  # the reference refuses or crashes on this combination
                print(f"[{it}] reference failed: {' '.join(args)} :: {str(e)[-120:]}"); continue
            try:
                got = helpers.run_cli(helpers.ORACLE_BIN, args, tmp, "orc.sam")
            except Exception as e:
                print(f"[{it}] ORACLE FAILED: {' '.join(args)} :: {str(e)[-200:]}"); bad += 1; continue
            if got != want:
                bad += 1
                g, w = got.splitlines(), want.splitlines()
                k = next((i for i, (x, y) in enumerate(zip(g, w)) if x != y), min(len(g), len(w)))
                print(f"[{it}] DIFF cfg{cid} {' '.join(args)}  lines {len(g)} vs {len(w)}, first at {k}:\n   got  {g[k][:160] if k < len(g) else None}\n   want {w[k][:160] if k < len(w) else None}")
            else:
                print(f"[{it}] ok cfg{cid} {rule} {' '.join(flags)}")
    print("differences:", bad)
    sys.exit(min(bad, 100))


if __name__ == "__main__":
    main()
