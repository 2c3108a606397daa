"""CPU checks of the two facts the one-bit screen (`screen_bits`) rests on, for every single-conversion rule.

Codes follow Param::SetAlign (param.cpp:216-233): from-base = 1, the single convert-to base = 3, the remaining two
bases get 0 and 2 in ACGT order; CountMismatch (align.h:126-128) lets reference code 1 match read codes 1 and 3 and
compares everything else exactly.
"""
import itertools

import numpy as np
import pytest

NT = "ACGT"
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}
RULES = [(f, t) for f in NT for t in NT if f != t]


def codes(frm: str, to: str) -> dict:
    c = {frm: 1, to: 3}
    rest = iter((0, 2))
    for b in NT:
        if b not in c:
            c[b] = next(rest)
    return c


def matches(ref_code: int, read_code: int) -> bool:
    """CountMismatch for one base: ((q & XC(r)) ^ r) == 0 with XC(01) = 01, XC(other) = 11."""
    xc = 1 if ref_code == 1 else 3
    return ((read_code & xc) ^ ref_code) == 0


@pytest.mark.parametrize("frm,to", RULES)
def test_low_bit_is_a_lower_bound(frm, to):
    """A match under the full rule implies equal low bits, so 'low bits differ' never over-counts mismatches."""
    c = codes(frm, to)
    for r, q in itertools.product(NT, NT):
        if matches(c[r], c[q]):
            assert (c[r] & 1) == (c[q] & 1), (frm, to, r, q)


@pytest.mark.parametrize("frm,to", RULES)
def test_complement_flips_the_low_bit_uniformly(frm, to):
    """low bit of code(complement(x)) = low bit of code(x) ^ flip with one flip for all four bases (DevIndex::flip):
    that is what lets reverse-strand windows be screened on the forward one-bit plane."""
    c = codes(frm, to)
    flips = {(c[b] ^ c[COMP[b]]) & 1 for b in NT}
    assert len(flips) == 1


@pytest.mark.parametrize("frm,to", [("C", "T"), ("A", "G"), ("G", "A"), ("A", "T")])
def test_mirrored_reverse_strand_count_is_a_lower_bound(frm, to):
    """Reverse-strand candidate: the read against the rc plane (full rule) vs. the REVERSED read's low bits against the
    forward plane's low bits ^ flip, over any sub-range of the read (what one 32-byte sector covers)."""
    rng = np.random.default_rng(11)
    c = codes(frm, to)
    flip = (c["A"] ^ c[COMP["A"]]) & 1
    for _ in range(200):
        L = int(rng.integers(30, 160))
        fwd = rng.choice(list(NT), size=L)                        # forward reference window, position p
        rc_codes = np.array([c[COMP[b]] for b in fwd[::-1]])      # rc plane codes at rc position k = L-1-p
        read = rng.choice(list(NT), size=L)
        # plant a near match so that both small and large counts occur
        if rng.random() < 0.5:
            inv = {v: k for k, v in c.items()}
            read = np.array([inv[int(x)] for x in rc_codes])
            for i in rng.choice(L, size=int(rng.integers(0, 6)), replace=False):
                read[i] = rng.choice(list(NT))
        q = np.array([c[b] for b in read])
        full = np.array([not matches(int(r), int(x)) for r, x in zip(rc_codes, q)])
        # mirrored: reversed read position k' = L-1-k faces forward position p = k'
        low_fwd = np.array([c[b] & 1 for b in fwd]) ^ flip
        low_read_rev = (q & 1)[::-1]
        low = low_fwd != low_read_rev                              # indexed by forward position p = L-1-k
        assert np.all(low <= full[::-1])
        a, b = sorted(rng.integers(0, L, size=2))
        assert low[a:b].sum() <= full[::-1][a:b].sum() <= full.sum()


# ------------------------------------------------------------------ gap_possible (reduce_round) is a necessary condition of GapAlign

def gap_search_finds(mm0, mms, L, thr, h, s, G):
    """The index search of GapAlign (align.cpp:348-410) over mismatch indicator arrays: mm0[k] at shift 0, mms[d][k] at
    shift d (d = -1, +1, -2, ...). Returns True when it would call AddHit."""
    if thr < 2:
        return False
    P0 = [k for k in range(L) if mm0[k]][: thr - 1]
    ret0 = P0[thr - 2] if len(P0) == thr - 1 else L
    P0 += [L] * (thr - 1 - len(P0))
    if ret0 < h + s:
        return False
    for tt in range(1, 2 * G + 1):
        t = (tt + 1) // 2
        sh = -t if tt % 2 else t
        sh1 = min(sh, 0)
        if thr < 1 + t:
            break
        PR = [k for k in range(L) if mms[sh][L - 1 - k]][: thr - 1]
        PR += [L] * (thr - 1 - len(PR))
        rl = L - t - 1
        for i in range(thr - t):
            gp = P0[i]
            if gp < 6 or gp >= rl:
                continue
            for j in range(thr - t - i):
                m2 = PR[j]
                if m2 < 6 or m2 >= rl:
                    continue
                if gp + m2 - sh1 < L:
                    continue
                return True
    return False


def gap_possible(mm0, mms, L, thr, G):
    """basal_b200/csrc/align.cu: gap_possible — mismatches of the first L/2 bases at shift 0, or of the last
    L - G - L/2 + 1 bases at one of the gap shifts, at most thr - 2."""
    if thr < 2:
        return False
    M = L // 2
    start = G + M - 1
    if sum(mm0[:M]) <= thr - 2:
        return True
    for tt in range(1, 2 * G + 1):
        t = (tt + 1) // 2
        sh = -t if tt % 2 else t
        if thr < 1 + t:
            break
        if sum(mms[sh][start:L]) <= thr - 2:
            return True
    return False


@pytest.mark.parametrize("G", [1, 2, 3])
def test_gap_possible_is_necessary(G):
    rng = np.random.default_rng(100 + G)
    found = 0
    for it in range(4000):
        L = int(rng.integers(40, 151))
        thr = int(rng.integers(0, 16))
        s = 16
        h = int(rng.integers(0, max(1, L - s)))
        # a read that really has one gap at a random place (so that the search succeeds often), plus noise, or pure noise
        dens = float(rng.choice([0.02, 0.05, 0.1, 0.3, 0.7]))
        shifts = [d for t in range(1, G + 1) for d in (-t, t)]
        if rng.random() < 0.6:
            gp = int(rng.integers(0, L)); d_true = int(rng.choice(shifts))
            mm0 = [(k >= gp and rng.random() < 0.7) or rng.random() < dens for k in range(L)]
            mms = {d: [((k < gp or d != d_true) and rng.random() < 0.7) or rng.random() < dens for k in range(L)] for d in shifts}
        else:
            mm0 = list(rng.random(L) < dens)
            mms = {d: list(rng.random(L) < dens) for d in shifts}
        for k in range(h, min(L, h + s)):
            mm0[k] = False                                        # the seed matched
        if gap_search_finds(mm0, mms, L, thr, h, s, G):
            found += 1
            assert gap_possible(mm0, mms, L, thr, G), (L, thr, h, G)
    assert found > 50                                             # the property was exercised
