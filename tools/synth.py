"""Deterministic synthetic references and base-conversion reads.

Test / bench tooling only: nothing here is on the product path.  All data is
generated with numpy ``default_rng`` from fixed seeds (``1000+config`` for the
reference, ``2000+config`` for the reads) so the GPU box and this container
produce identical bytes.

Shapes follow SURVEY.md §8(d):
  reference  iid ACGT, ``nchr`` equal chromosomes, 0.5 % of the length
             overwritten by repeat families (2 kb element x 20..200 copies,
             0..2 % divergence), one 10 kb N run per chromosome.
  reads      uniform start, strand 50/50, fixed length, conversion of the
             from-base on the fragment's own strand, 0.5 % substitutions, 1 % of
             reads with 1..3 N, optional deletions of the from-base (``to`` holds
             '-') and optional 1..3 bp indels.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Iterator

import numpy as np

ACGT = np.frombuffer(b"ACGTN", dtype=np.uint8)
CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    CODE[_c] = _i
    CODE[_c + 32] = _i


@dataclasses.dataclass
class Config:
    """One BASELINE.json config (or a scaled-down copy of it)."""
    cid: int                 # 1..5, selects the rng seeds
    rule: str                # -M value, e.g. "C:T", "A:CGT", "T:-"
    ref_len: int
    n_reads: int             # reads (SE) or pairs (PE)
    read_len: int
    paired: bool
    conv_rate: float
    flags: tuple = ()        # extra basal flags, e.g. ("-w", "100")
    indel_frac: float = 0.0
    nchr: int = 24
    name: str = ""


def baseline_config(cid: int, scale: float = 1.0) -> Config:
    """The five BASELINE.json configs; ``scale`` shrinks reference and read count."""
    mb = 1_000_000
    s = lambda x: max(int(x * scale), 1)
    if cid == 1:
        return Config(1, "C:T", s(50 * mb), s(1_000_000), 100, False, 0.95,
                      ("-v", "0.1", "-g", "0", "-s", "16", "-I", "4"), name="C1 C:T SE100 50Mb")
    if cid == 2:
        return Config(2, "A:G", s(500 * mb), s(10_000_000), 150, True, 0.95, (), name="C2 A:G PE150 500Mb")
    if cid == 3:
        return Config(3, "A:CGT", s(500 * mb), s(20_000_000), 100, False, 0.05, ("-w", "100"),
                      name="C3 A:CGT SE100 500Mb -w 100")
    if cid == 4:
        return Config(4, "T:-", s(500 * mb), s(20_000_000), 100, False, 0.02, ("-g", "3"),
                      indel_frac=0.02, name="C4 T:- SE100 500Mb -g 3")
    if cid == 5:
        return Config(5, "C:T", s(3100 * mb), s(100_000_000), 150, True, 0.95, (), name="C5 C:T PE150 3.1Gb")
    raise ValueError(cid)


# --------------------------------------------------------------------------- reference

def make_reference(cfg: Config) -> list[tuple[str, np.ndarray]]:
    """Return [(name, uint8 codes 0..3 / 4=N)] for every chromosome."""
    rng = np.random.default_rng(1000 + cfg.cid)
    nchr = cfg.nchr
    clen = max(cfg.ref_len // nchr, 200)
    total = clen * nchr
    g = rng.integers(0, 4, size=total, dtype=np.uint8)
    # repeat families: ~0.5 % of the genome
    budget = int(total * 0.005)
    elem = 2000 if clen > 40_000 else max(clen // 20, 50)
    nfam = 50
    placed = 0
    for f in range(nfam):
        if placed >= budget:
            break
        base = rng.integers(0, 4, size=elem, dtype=np.uint8)
        copies = int(rng.integers(20, 201))
        copies = max(2, min(copies, (budget - placed) // elem if elem else 0, total // (4 * elem)))
        div = rng.uniform(0.0, 0.02)
        for _ in range(copies):
            c = int(rng.integers(0, nchr))
            p = int(rng.integers(0, clen - elem))
            e = base.copy()
            m = rng.random(elem) < div
            e[m] = (e[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
            if rng.random() < 0.5:
                e = (3 - e)[::-1]
            g[c * clen + p: c * clen + p + elem] = e
            placed += elem
    chrs = []
    nrun = 10_000 if clen > 200_000 else max(clen // 50, 20)
    for c in range(nchr):
        seq = g[c * clen:(c + 1) * clen].copy()
        p = int(rng.integers(0, clen - nrun))
        seq[p:p + nrun] = 4
        chrs.append((f"chr{c + 1}", seq))
    return chrs


def write_fasta(chrs: list[tuple[str, np.ndarray]], path: str, width: int = 60) -> None:
    with open(path, "wb") as fh:
        for name, codes in chrs:
            fh.write(b">" + name.encode() + b"\n")
            asc = ACGT[codes]
            n = len(asc)
            full = (n // width) * width
            if full:
                block = np.empty((full // width, width + 1), dtype=np.uint8)
                block[:, :width] = asc[:full].reshape(-1, width)
                block[:, width] = 10
                fh.write(block.tobytes())
            if n > full:
                fh.write(asc[full:].tobytes() + b"\n")


def reference_ascii(chrs: list[tuple[str, np.ndarray]]) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Concatenated ASCII bases + offsets + lengths (the bsl_index_build inputs)."""
    lens = np.array([len(c) for _, c in chrs], dtype=np.uint32)
    offs = np.zeros(len(chrs), dtype=np.uint64)
    offs[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    cat = np.concatenate([ACGT[c] for _, c in chrs])
    return cat, offs, lens


# --------------------------------------------------------------------------- reads

def _parse_rule(rule: str) -> tuple[int, list[int], bool]:
    frm = int(CODE[ord(rule[0])])
    tos = [int(CODE[ord(ch)]) for ch in rule[2:] if ch != "-"]
    return frm, tos, "-" in rule[2:]


def _revcomp(codes: np.ndarray) -> np.ndarray:
    out = codes[:, ::-1].copy()
    m = out < 4
    out[m] = 3 - out[m]
    return out


class ReadSimulator:
    """Chunked read generator; ``chunks()`` yields (mate1, mate2|None) uint8 ASCII matrices."""

    def __init__(self, cfg: Config, chrs: list[tuple[str, np.ndarray]]):
        self.cfg = cfg
        self.clen = len(chrs[0][1])
        self.nchr = len(chrs)
        self.genome = np.concatenate([c for _, c in chrs])
        self.frm, self.tos, self.has_del = _parse_rule(cfg.rule)
        self.rng = np.random.default_rng(2000 + cfg.cid)

    def _window(self, pos: np.ndarray, width: int) -> np.ndarray:
        idx = pos[:, None] + np.arange(width, dtype=np.int64)[None, :]
        return self.genome[idx]

    def _convert(self, frag: np.ndarray) -> np.ndarray:
        rng = self.rng
        cfg = self.cfg
        if self.tos:
            m = (frag == self.frm) & (rng.random(frag.shape) < cfg.conv_rate)
            k = int(m.sum())
            if k:
                pick = np.array(self.tos, dtype=np.uint8)[rng.integers(0, len(self.tos), size=k)]
                frag[m] = pick
        return frag

    def _errors_and_ns(self, reads: np.ndarray) -> np.ndarray:
        rng = self.rng
        n, L = reads.shape
        m = (rng.random(reads.shape) < 0.005) & (reads < 4)
        k = int(m.sum())
        if k:
            reads[m] = (reads[m] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3
        who = np.flatnonzero(rng.random(n) < 0.01)
        if len(who):
            cnt = rng.integers(1, 4, size=len(who))
            for t in range(3):
                sel = who[cnt > t]
                reads[sel, rng.integers(0, L, size=len(sel))] = 4
        return reads

    def _mate(self, strand: np.ndarray, left: np.ndarray, L: int) -> np.ndarray:
        """L bases in fragment orientation whose reference window starts at ``left``."""
        cfg = self.cfg
        spare = 24 if (self.has_del or cfg.indel_frac > 0) else 0
        W = L + spare
        # for the reverse strand the window extends to the left of its end
        lo = np.where(strand == 0, left, left - spare)
        lo = np.clip(lo, 0, len(self.genome) - W)
        win = self._window(lo, W)
        rc = strand == 1
        if rc.any():
            win[rc] = _revcomp(win[rc])
        win = self._convert(win)
        if spare:
            rng = self.rng
            n = win.shape[0]
            keep = np.ones(win.shape, dtype=bool)
            if self.has_del:
                keep &= ~((win == self.frm) & (rng.random(win.shape) < cfg.conv_rate))
            src = np.argsort(~keep, axis=1, kind="stable")[:, :L + 8]
            win = np.take_along_axis(win, src, axis=1)
            if cfg.indel_frac > 0:
                who = np.flatnonzero(rng.random(n) < cfg.indel_frac)
                if len(who):
                    k = rng.integers(1, 4, size=len(who))
                    p = rng.integers(10, L - 10, size=len(who))
                    ins = rng.random(len(who)) < 0.5
                    j = np.arange(L, dtype=np.int64)[None, :]
                    sub = win[who]
                    # deletion of k read bases at p: j -> j (j<p) | j+k
                    didx = np.where(j < p[:, None], j, j + k[:, None])
                    dele = np.take_along_axis(sub, didx, axis=1)
                    # insertion of k random bases at p
                    iidx = np.where(j < p[:, None], j, np.maximum(j - k[:, None], 0))
                    inse = np.take_along_axis(sub, iidx, axis=1)
                    rnd = rng.integers(0, 4, size=inse.shape, dtype=np.uint8)
                    inside = (j >= p[:, None]) & (j < (p + k)[:, None])
                    inse[inside] = rnd[inside]
                    new = np.where(ins[:, None], inse, dele)
                    out = win[:, :L].copy()
                    out[who] = new
                    win = out
            win = win[:, :L]
        return np.ascontiguousarray(win)

    def chunks(self, chunk: int = 500_000, limit: int | None = None) -> Iterator[tuple[np.ndarray, np.ndarray | None]]:
        cfg = self.cfg
        rng = self.rng
        L = cfg.read_len
        total = cfg.n_reads if limit is None else min(limit, cfg.n_reads)
        done = 0
        while done < total:
            n = min(chunk, total - done)
            c = rng.integers(0, self.nchr, size=n)
            strand = (rng.random(n) < 0.5).astype(np.int8)
            if cfg.paired:
                ins = np.clip(np.rint(rng.normal(300, 30, size=n)), L, min(600, self.clen - 1)).astype(np.int64)
            else:
                ins = np.full(n, L, dtype=np.int64)
            margin = 32
            p = rng.integers(margin, self.clen - 600 - margin, size=n) if self.clen > 1400 else \
                rng.integers(0, max(self.clen - int(ins.max()), 1), size=n)
            g0 = c.astype(np.int64) * self.clen + p
            # mate 1 = first L bases of the fragment (fragment orientation)
            left1 = np.where(strand == 0, g0, g0 + ins - L)
            m1 = self._errors_and_ns(self._mate(strand, left1, L))
            m2 = None
            if cfg.paired:
                left2 = np.where(strand == 0, g0 + ins - L, g0)
                m2 = self._mate(strand, left2, L)
                m2 = self._errors_and_ns(_revcomp(m2))
            yield ACGT[m1], (ACGT[m2] if m2 is not None else None)
            done += n


def fastq_bytes(reads: np.ndarray, first_index: int, suffix: str = "") -> bytes:
    """FASTQ text of (n, L) uint8 ASCII reads; names r<i><suffix>, qualities 'I' (vectorised per digit count)."""
    n, L = reads.shape
    idx = np.arange(first_index, first_index + n, dtype=np.int64)
    suf = np.frombuffer(suffix.encode(), dtype=np.uint8)
    parts = []
    lo = 0
    while lo < n:
        d = len(str(int(idx[lo])))
        hi = int(np.searchsorted(idx, 10 ** d, side="left"))
        hi = max(hi, lo + 1)
        k = hi - lo
        width = 2 + d + len(suf) + 1 + L + 3 + L + 1
        rec = np.empty((k, width), dtype=np.uint8)
        rec[:, 0] = ord("@"); rec[:, 1] = ord("r")
        sub = idx[lo:hi]
        for j in range(d):
            rec[:, 2 + j] = 48 + (sub // 10 ** (d - 1 - j)) % 10
        p = 2 + d
        if len(suf):
            rec[:, p:p + len(suf)] = suf
            p += len(suf)
        rec[:, p] = 10; p += 1
        rec[:, p:p + L] = reads[lo:hi]; p += L
        rec[:, p] = 10; rec[:, p + 1] = ord("+"); rec[:, p + 2] = 10; p += 3
        rec[:, p:p + L] = ord("I"); p += L
        rec[:, p] = 10
        parts.append(rec.tobytes())
        lo = hi
    return b"".join(parts)


def write_fastq(path: str, reads: np.ndarray, first_index: int, suffix: str = "", append: bool = False) -> None:
    with open(path, "ab" if append else "wb") as fh:
        fh.write(fastq_bytes(reads, first_index, suffix))


def materialise(cfg: Config, outdir: str, limit: int | None = None, chunk: int = 500_000) -> dict:
    """Write ref.fa + reads(_1/_2).fq for ``cfg`` into outdir; returns paths."""
    os.makedirs(outdir, exist_ok=True)
    chrs = make_reference(cfg)
    ref = os.path.join(outdir, "ref.fa")
    write_fasta(chrs, ref)
    sim = ReadSimulator(cfg, chrs)
    a = os.path.join(outdir, "reads_1.fq" if cfg.paired else "reads.fq")
    b = os.path.join(outdir, "reads_2.fq") if cfg.paired else None
    idx = 0
    first = True
    for m1, m2 in sim.chunks(chunk=chunk, limit=limit):
        write_fastq(a, m1, idx, "/1" if cfg.paired else "", append=not first)
        if m2 is not None:
            write_fastq(b, m2, idx, "/2", append=not first)
        idx += len(m1)
        first = False
    return {"ref": ref, "a": a, "b": b, "n": idx}


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--scale", type=float, default=0.01)
    ap.add_argument("--limit", type=int, default=None)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    print(materialise(baseline_config(a.config, a.scale), a.out, a.limit))
