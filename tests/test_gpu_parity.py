"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on seeded inputs.

Bit-exact bar (integer / byte / index work): every field of every result record must be equal.
"""
import numpy as np
import pytest

import helpers
from basal_b200 import capi

pytestmark = pytest.mark.gpu


def _build_both(cfg, chrs, params):
    cat, offs, lens = helpers.synth.reference_ascii(chrs)
    gpu = capi.Context(params, device=0)
    gpu.index_build(cat, offs, lens)
    orc = helpers.oracle_context(params)
    orc.index_build(cat, offs, lens)
    return gpu, orc


@pytest.mark.parametrize("s,I", [(16, 4), (12, 2), (10, 1), (13, 5)])
def test_index_matches_oracle(s, I):
    cfg, chrs, m1, _ = helpers.small_case(1, 0.004, limit=10)
    params = capi.make_params(rule="C:T", s=s, I=I, s_given=True)
    gpu, orc = _build_both(cfg, chrs, params)
    gi, oi = gpu.index_info(), orc.index_info()
    for f in ("n_seq", "n_kmers", "sum_length", "n_words", "n_entries", "max_kmer_num"):
        assert getattr(gi, f) == getattr(oi, f), f
    g = gpu.index_download(); o = orc.index_download()
    for name, a, b in zip(("bucket_start", "n_fwd", "loc", "fwd_plane", "rc_plane"), g, o):
        assert np.array_equal(a, b), name
    gpu.close(); orc.close()


def test_index_with_iupac_lowercase_and_short_islands():
    rng = np.random.default_rng(5)
    seqs = []
    for n in (5000, 333, 64, 40, 17):
        a = rng.integers(0, 4, n)
        s = np.frombuffer(b"ACGT", np.uint8)[a].copy()
        s[rng.random(n) < 0.02] = ord("N")
        s[rng.random(n) < 0.01] = ord("R")
        s[rng.random(n) < 0.05] |= 32           # lowercase
        s[n // 2: n // 2 + 9] = ord("n")
        seqs.append(s)
    lens = np.array([len(x) for x in seqs], np.uint32)
    offs = np.zeros(len(seqs), np.uint64); offs[1:] = np.cumsum(lens[:-1])
    cat = np.concatenate(seqs)
    for rule in ("C:T", "A:CGT", "T:-"):
        params = capi.make_params(rule=rule, s=11, I=3, s_given=True)
        gpu = capi.Context(params); gpu.index_build(cat, offs, lens)
        orc = helpers.oracle_context(params); orc.index_build(cat, offs, lens)
        for name, a, b in zip(("bucket_start", "n_fwd", "loc", "fwd_plane", "rc_plane"), gpu.index_download(), orc.index_download()):
            assert np.array_equal(a, b), (rule, name)
        assert gpu.index_info().max_kmer_num == orc.index_info().max_kmer_num
        gpu.close(); orc.close()


SE_CASES = [
    (1, 0.01, {}),                          # C:T SE100
    (3, 0.002, {}),                         # A:CGT -w 100
    (4, 0.002, {}),                         # T:- -g 3
    (1, 0.01, {"g": 2, "n": 1}),            # gapped single conversion, all four strands
    (3, 0.002, {"w": 3}),                   # -w feedback
    (4, 0.002, {"n": 2, "w": 2}),
    (4, 0.002, {"g": 0, "n": 1}),           # '-'-only rule without gaps: one-bit screen + CountMismatch_new exact count
    (1, 0.01, {"g": 3, "w": 3}),            # single conversion with gaps: one-bit screen with the prefix bound of GapAlign's first test
]


@pytest.mark.parametrize("cid,scale,extra", SE_CASES)
def test_se_matches_oracle(cid, scale, extra):
    cfg, chrs, m1, _ = helpers.small_case(cid, scale, limit=20000)
    params = helpers.flags_to_params(cfg, extra)
    gpu, orc = _build_both(cfg, chrs, params)
    batch = capi.ReadBatch.from_matrix(m1, readset=0, first_index=0)
    got = gpu.align_se(batch); want = orc.align_se(batch)
    helpers.assert_records_equal(got, want, f"SE config {cid} {extra}")
    gs, os_ = gpu.stats(), orc.stats()
    assert gs.seed_lookups == os_.seed_lookups and gs.candidates == os_.candidates
    gpu.close(); orc.close()


@pytest.mark.parametrize("cid,rule,extra", [(2, "A:G", {}), (5, "C:T", {}), (2, "A:G", {"g": 1, "w": 4}), (2, "T:-", {"g": 3, "n": 1})])
def test_pe_matches_oracle(cid, rule, extra):
    cfg, chrs, m1, m2 = helpers.small_case(cid, 0.001 if cid == 2 else 0.0002, limit=10000)
    kw = dict(extra); kw["rule"] = rule
    params = helpers.flags_to_params(cfg, kw)
    gpu, orc = _build_both(cfg, chrs, params)
    a = capi.ReadBatch.from_matrix(m1, readset=1); b = capi.ReadBatch.from_matrix(m2, readset=2)
    ga, gb, gp = gpu.align_pe(a, b); wa, wb, wp = orc.align_pe(a, b)
    helpers.assert_records_equal(gp, wp, "pair records", fields=["n_pairs", "insert", "chain", "na", "nb"])
    helpers.assert_records_equal(ga, wa, "mate 1 records")
    helpers.assert_records_equal(gb, wb, "mate 2 records")
    gpu.close(); orc.close()


@pytest.mark.parametrize("cid,rule,extra", [(1, "C:T", {"n": 1}), (2, "A:G", {}), (1, "A:T", {"n": 1}), (2, "G:A", {"n": 1})])
def test_large_sequences_match_oracle(cid, rule, extra):
    """Three sequences of ~330 kb: windows whose 65 536-coordinate block lies wholly inside one sequence take the
    strand-table path of the one-bit screen, the ones around the boundaries take the anchor search (both strands)."""
    import dataclasses
    cfg = dataclasses.replace(helpers.synth.baseline_config(cid, 0.02 if cid == 1 else 0.002), nchr=3)
    chrs = helpers.synth.make_reference(cfg)
    sim = helpers.synth.ReadSimulator(cfg, chrs)
    m1, m2 = next(sim.chunks(limit=12000))
    kw = dict(extra); kw["rule"] = rule
    params = helpers.flags_to_params(cfg, kw)
    gpu, orc = _build_both(cfg, chrs, params)
    if m2 is None:
        batch = capi.ReadBatch.from_matrix(m1, readset=0, first_index=0)
        helpers.assert_records_equal(gpu.align_se(batch), orc.align_se(batch), f"SE {rule}")
    else:
        a = capi.ReadBatch.from_matrix(m1, readset=1); b = capi.ReadBatch.from_matrix(m2, readset=2)
        ga, gb, gp = gpu.align_pe(a, b); wa, wb, wp = orc.align_pe(a, b)
        helpers.assert_records_equal(gp, wp, "pair records", fields=["n_pairs", "insert", "chain", "na", "nb"])
        helpers.assert_records_equal(ga, wa, "mate 1 records"); helpers.assert_records_equal(gb, wb, "mate 2 records")
    gs, os_ = gpu.stats(), orc.stats()
    assert gs.seed_lookups == os_.seed_lookups and gs.candidates == os_.candidates
    gpu.close(); orc.close()


@pytest.mark.parametrize("env", [{"BSL_SUB_BATCH": "3000"}, {"BSL_CAND_CAP": "60000"}, {"BSL_HIT_CAP": "2"},
                                 {"BSL_SUB_BATCH": "1500", "BSL_CAND_CAP": "30000", "BSL_HIT_CAP": "3"}])
@pytest.mark.parametrize("pe", [False, True])
def test_capacity_paths_match_oracle(env, pe, monkeypatch):
    """Sub-ranges, a flat candidate space that runs out (reads re-run on the large-capacity path) and tiny hit
    lists must not change a single record."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    if pe:
        cfg, chrs, m1, m2 = helpers.small_case(2, 0.001, limit=8000)
        params = helpers.flags_to_params(cfg, {"g": 1, "n": 1})
        gpu, orc = _build_both(cfg, chrs, params)
        a = capi.ReadBatch.from_matrix(m1, readset=1); b = capi.ReadBatch.from_matrix(m2, readset=2)
        ga, gb, gp = gpu.align_pe(a, b); wa, wb, wp = orc.align_pe(a, b)
        helpers.assert_records_equal(gp, wp, "pair records", fields=["n_pairs", "insert", "chain", "na", "nb"])
        helpers.assert_records_equal(ga, wa, "mate 1 records"); helpers.assert_records_equal(gb, wb, "mate 2 records")
    else:
        cfg, chrs, m1, _ = helpers.small_case(3, 0.002, limit=12000)
        params = helpers.flags_to_params(cfg, {"r": 2})
        gpu, orc = _build_both(cfg, chrs, params)
        batch = capi.ReadBatch.from_matrix(m1, readset=0, first_index=0)
        got, gall = gpu.align_se(batch, all_cap=400000); want, wall = orc.align_se(batch, all_cap=400000)
        helpers.assert_records_equal(got, want, "SE records")
        assert len(gall) == int(got["n_hits"][got["status"] == capi.BSL_ST_MULTI].sum())      # the product lists repeat hits only; the oracle also lists unique ones
        for i in np.flatnonzero(got["status"] == capi.BSL_ST_MULTI)[:2000]:
            g = gall[got["all_first"][i]: got["all_first"][i] + got["n_hits"][i]]; w = wall[want["all_first"][i]: want["all_first"][i] + want["n_hits"][i]]
            assert np.array_equal(g["loc"], w["loc"]) and np.array_equal(g["chr"], w["chr"])
        gs, os_ = gpu.stats(), orc.stats()
        assert gs.seed_lookups == os_.seed_lookups and gs.candidates == os_.candidates
    gpu.close(); orc.close()


# ------------------------------------------------------------------ mixed read lengths: carried aligner state (SURVEY trap 3)

def _mixed(matrix, rng, lengths):
    """Cut every read of a fixed-length matrix to a length drawn from `lengths`; returns the reads as strings."""
    return [bytes(row[: int(rng.choice(lengths))]).decode() for row in matrix]


@pytest.mark.parametrize("pe,env", [(False, {}), (True, {}), (False, {"BSL_SUB_BATCH": "1500"}), (True, {"BSL_SUB_BATCH": "1200"})])
def test_mixed_lengths_match_oracle(pe, env, monkeypatch):
    """Reads whose start-offset range is empty ((L - I + 1) % s == 0: 99, 83, 67, 51 / 147, 131 at the defaults) inherit the
    start offset and the stale seed hashes of earlier reads of the same aligner object (align.cpp:476-480, 79-150). The
    oracle keeps that state like a -p 1 run of the reference (tools/fuzz_mixed.py pins it against the binary); the CUDA
    path must give the same records AND the same look-up / candidate counts, also when the call is cut into sub-ranges."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(11)
    if pe:
        cfg, chrs, m1, m2 = helpers.small_case(2, 0.001, limit=6000)
        lengths = [150, 147, 149, 131, 148, 115, 150, 99, 140]
    else:
        cfg, chrs, m1, m2 = helpers.small_case(1, 0.01, limit=8000)
        lengths = [100, 99, 98, 83, 97, 67, 96, 51, 100]
    params = helpers.flags_to_params(cfg, {"n": 1} if pe else {})
    gpu, orc = _build_both(cfg, chrs, params)
    a = capi.ReadBatch.from_strings(_mixed(m1, rng, lengths), readset=1 if pe else 0)
    if pe:
        b = capi.ReadBatch.from_strings(_mixed(m2, rng, lengths), readset=2)
        ga, gb, gp = gpu.align_pe(a, b); wa, wb, wp = orc.align_pe(a, b)
        helpers.assert_records_equal(gp, wp, "pair records", fields=["n_pairs", "insert", "chain", "na", "nb"])
        helpers.assert_records_equal(ga, wa, "mate 1 records"); helpers.assert_records_equal(gb, wb, "mate 2 records")
    else:
        helpers.assert_records_equal(gpu.align_se(a), orc.align_se(a), "SE records")
    gs, os_ = gpu.stats(), orc.stats()
    assert gs.seed_lookups == os_.seed_lookups and gs.candidates == os_.candidates
    gpu.close(); orc.close()


def test_context_reads_restore_the_carried_state():
    """bsl_batch::n_context: mapping reads [k, n) with the right earlier reads in front as context gives the records the
    whole batch gives for them (how the CLI keeps -p 1 semantics across its batches)."""
    rng = np.random.default_rng(12)
    cfg, chrs, m1, _ = helpers.small_case(1, 0.01, limit=3000)
    reads = _mixed(m1, rng, [100, 99, 98, 83, 97, 67, 96, 100])
    params = helpers.flags_to_params(cfg, {})
    gpu, orc = _build_both(cfg, chrs, params)
    whole = gpu.align_se(capi.ReadBatch.from_strings(reads))
    k = 1700
    ctx = reads[:k]                                   # (a superset of) the reads that matter; all of them as context
    part = gpu.align_se(capi.ReadBatch.from_strings(ctx + reads[k:], first_index=0, n_context=k))
    helpers.assert_records_equal(part[k:], whole[k:], "records after a context prefix")
    want = orc.align_se(capi.ReadBatch.from_strings(ctx + reads[k:], first_index=0, n_context=k))
    helpers.assert_records_equal(part[k:], want[k:], "records after a context prefix, oracle")
    gpu.close(); orc.close()
