"""Shared test helpers: oracle binding, reference binary runner, SAM comparison, synthetic cases."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from basal_b200 import capi  # noqa: E402
import synth  # noqa: E402

ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
ORACLE_BIN = os.path.join(ROOT, "oracle", "_build", "basal_oracle")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "basal")
GPU_BIN = os.path.join(ROOT, "basal_b200", "bin", "basal")
GOLDEN = os.path.join(ROOT, "tests", "golden")

_oracle_api = None


def oracle_api() -> capi._Api:
    """ctypes table of oracle/_build/liboracle.so (orc_* functions share basal_gpu.h's structs)."""
    global _oracle_api
    if _oracle_api is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
        _oracle_api = capi._Api(C.CDLL(ORACLE_SO), "orc_", False)
    return _oracle_api


def oracle_context(params) -> capi.Context:
    return capi.Context(params, api=oracle_api())


def have_ref() -> bool:
    return os.path.exists(REF_BIN)


def run_cli(binary: str, args: list[str], cwd: str, out: str) -> str:
    """Run a basal-compatible binary with -o out; returns the SAM text without the @PG line."""
    cmd = [binary] + args + ["-o", out]
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=3600)
    if p.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed ({p.returncode}): {p.stderr.decode()[-2000:]}")
    with open(os.path.join(cwd, out)) as fh:
        return "".join(l for l in fh if not l.startswith("@PG"))


def sam_records(text: str) -> list[str]:
    return [l for l in text.splitlines() if l and not l.startswith("@")]


def small_case(cid: int, scale: float, limit: int | None = None):
    """(cfg, chrs, mate1 matrix, mate2 matrix|None) of a scaled-down BASELINE config, all in memory."""
    cfg = synth.baseline_config(cid, scale)
    chrs = synth.make_reference(cfg)
    sim = synth.ReadSimulator(cfg, chrs)
    m1s, m2s = [], []
    for m1, m2 in sim.chunks(limit=limit):
        m1s.append(m1)
        if m2 is not None:
            m2s.append(m2)
    return cfg, chrs, np.concatenate(m1s), (np.concatenate(m2s) if m2s else None)


def flags_to_params(cfg, extra: dict | None = None):
    """capi.Params for a synth.Config (its flags tuple is in basal CLI syntax)."""
    kw = dict(rule=cfg.rule)
    f = list(cfg.flags)
    i = 0
    while i < len(f):
        k, v = f[i], f[i + 1]
        if k == "-v": kw["v"] = v
        elif k == "-g": kw["g"] = int(v)
        elif k == "-s": kw["s"] = int(v); kw["s_given"] = True
        elif k == "-I": kw["I"] = int(v)
        elif k == "-w": kw["w"] = int(v)
        i += 2
    if extra:
        kw.update(extra)
    return capi.make_params(**kw)


def assert_records_equal(got: np.ndarray, want: np.ndarray, what: str, fields=None):
    fields = fields or [n for n in got.dtype.names if n not in ("all_first",)]
    for f in fields:
        g, w = got[f], want[f]
        if not np.array_equal(g, w):
            bad = np.flatnonzero(g != w)
            i = int(bad[0])
            raise AssertionError(f"{what}: field {f} differs at {len(bad)} of {len(g)} records; first #{i}: got {got[i]} want {want[i]}")
